#!/usr/bin/env python3
"""integration/patch_server.py -- applies the maintainer's change of INTEGRATION.md section 5 to the reference's examples/server/server.cpp AT BUILD TIME.

    patch_server.py <reference>/examples/server/server.cpp <out.cpp> <reference>/examples/server/public/index.html <index.html.hpp>

Nothing of the reference is stored in this repository: the patched translation unit and the generated asset header land in integration/_build/ (git-ignored, like
oracle/_ref).  Three edits, each required to match exactly once (an upstream change that moves them fails the build loudly instead of silently building a stale loop):
  1. include b200_batch_worker.h (and the loaders header by an include-path name: the copy no longer sits two directories below src/);
  2. worker::loop() (server.cpp:247-254) -> b200::batch_loop, max_batch from B2TTS_SERVER_MAX_BATCH (default 32);
  3. the speech handler frees the task-owned PCM after write_audio_data (server.cpp:709).
The second output is what the reference's cmake/xxd.cmake generates from public/index.html (`xxd -i`)."""
import re
import sys


def once(src: str, pattern: str, repl: str, what: str) -> str:
    out, n = re.subn(pattern, lambda m: repl, src, flags=re.S)
    if n != 1:
        sys.exit(f"patch_server.py: {what}: expected exactly one match, found {n} -- upstream server.cpp changed, update INTEGRATION.md section 5")
    return out


def main() -> None:
    src_path, out_path, html_path, hpp_path = sys.argv[1:5]
    s = open(src_path).read()
    s = once(s, r'#include "\.\./\.\./src/models/loaders\.h"', '#include "models/loaders.h"\n#include "b200_batch_worker.h"', "loaders include")
    loop_upstream = ("    void loop() {\n"
                     "        while (running) {\n"
                     "            struct simple_server_task * task = task_queue->get_next();\n"
                     "            if (task) {\n"
                     "                process_task(task);\n"
                     "            }\n"
                     "        }\n"
                     "    }\n")                                          # the worker loop as upstream has it (server.cpp:247-254), matched literally
    s = once(s, re.escape(loop_upstream),
             '    void loop() {\n'
             '        const char * mb = getenv("B2TTS_SERVER_MAX_BATCH");\n'
             '        b200::batch_loop(running, *task_queue, *response_map, task_timeout, (size_t) (mb ? atoi(mb) : 32), TTS,\n'
             '                         [&](simple_server_task * t) -> tts_generation_runner & { return *runners[t->model]; },\n'
             '                         [&](simple_server_task * t) { process_task(t); });\n'
             '    }\n', "worker::loop")
    s = once(s, r'(        bool success = write_audio_data\(\(float \*\)rtask->response, rtask->length, audio, audio_type, rtask->sample_rate\);\n)',
             '        bool success = write_audio_data((float *)rtask->response, rtask->length, audio, audio_type, rtask->sample_rate);\n        b200::release(rtask);\n',
             "speech handler")
    open(out_path, "w").write(s)
    data = open(html_path, "rb").read()
    open(hpp_path, "w").write("unsigned char index_html[] = {" + ",".join(f"0x{b:02x}" for b in data) + "};\nunsigned int index_html_len = %d;\n" % len(data))


if __name__ == "__main__":
    main()
