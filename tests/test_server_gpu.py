"""GPU: SURVEY 8(f) row 1, the batch-draining server worker (integration/b200_batch_worker.h), run on the B200 runners.

  * worker_demo: five queued prompts are served as forwards of 4 + 1 and every task gets the PCM the same prompt gives one after another through generate();
  * tts-server-b200 = the reference's examples/server/server.cpp with the three build-time edits of INTEGRATION.md section 5: 24 concurrent /v1/audio/speech
    requests over real HTTP are served in batched forwards.
Collected last (conftest._LATE): these are the newest tests of the round and follow every parity test under -x."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, report, rms, synth_gguf

pytestmark = pytest.mark.gpu
BUILD = os.path.join(ROOT, "integration", "_build")


def _need(path):
    assert os.path.exists(path), f"{path} missing: built by `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists, shipped with the snapshot"
    return path


@pytest.fixture(scope="module")
def text_gguf():
    return synth_gguf(text_vocab=True)


def test_batch_draining_worker_serves_the_queue_in_batched_forwards(text_gguf):
    tmp = tempfile.mkdtemp(prefix="b2worker_")
    pf = os.path.join(tmp, "prompts.txt")
    prompts = ["hello world this is a test", "the quick brown fox jumps over the lazy dog", "a second runner starts a fresh noise stream",
               "glue the sheet to the dark blue background", "these days a chicken leg is a rare dish"]
    open(pf, "w").write("\n".join(prompts) + "\n")
    pre = os.path.join(tmp, "o")
    r = subprocess.run([_need(os.path.join(BUILD, "worker_demo")), text_gguf, pf, pre, "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-1500:])
    print(r.stdout.strip().splitlines()[-1])
    assert "batches 4 1;" in r.stdout                          # five queued tasks, max_batch 4: one forward of 4, one of 1
    for i in range(len(prompts)):
        a, b = np.fromfile(f"{pre}.worker.{i}.f32", np.float32), np.fromfile(f"{pre}.single.{i}.f32", np.float32)
        d, rr, mx = report(f"batch worker vs generate, prompt {i}", a, b)
        assert a.shape == b.shape and a.size > 0 and d < 5e-3 * max(rr, 1e-6)   # same bar as generate_batch above


def test_patched_reference_server_batches_concurrent_http_requests(text_gguf):
    """examples/server/server.cpp with the batch-draining worker (three build-time edits) over real HTTP on the GPU: concurrent
    /v1/audio/speech requests are served in batched forwards; every response has the sample count the same prompt has through generate() (durations do not depend on
    the noise stream; the PCM values do, by the order of arrival) and a sane level."""
    from conftest import patched_server
    _need(os.path.join(BUILD, "tts-server-b200"))
    prompts = ["hello world this is a test", "the quick brown fox jumps over the lazy dog", "a second runner starts a fresh noise stream",
               "glue the sheet to the dark blue background", "these days a chicken leg is a rare dish", "the birch canoe slid on the smooth planks"]
    tmp = tempfile.mkdtemp(prefix="b2srv_")
    pf = os.path.join(tmp, "prompts.txt")
    open(pf, "w").write("\n".join(prompts) + "\n")
    pre = os.path.join(tmp, "o")
    r = subprocess.run([_need(os.path.join(BUILD, "batch_demo")), text_gguf, pf, pre], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-1500:])
    want = [np.fromfile(f"{pre}.single.{i}.f32", np.float32) for i in range(len(prompts))]
    with patched_server(text_gguf, max_batch=32) as srv:
        out = srv.speech(prompts * 4, threads=24)               # 24 requests, all in flight at once
        fw = srv.forwards()
    print(f"patched server: {len(out)} requests served in forwards of {fw}")
    assert sum(fw) == len(out) and len(fw) < len(out), fw      # every task through the batch loop, and at least one forward carried several
    for k, (code, pcm, rate) in enumerate(out):
        w = want[k % len(prompts)]
        assert code == 200 and rate == 24000 and pcm.size == w.size, (k, code, rate, pcm.size, w.size)
        assert 0.7 < rms(pcm / 32767.0) / rms(np.clip(w, -1, 1)) < 1.4
