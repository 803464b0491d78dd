"""GPU: the drop-in, run.  integration/_build/{tts-cli,perf_battery} are the reference's UNMODIFIED examples/cli/cli.cpp and examples/perf_battery/perf_battery.cpp linked
against a library named `tts` that holds the reference's own host code (registry, GGUF loading, phonemizer, tokenizer) plus the B200 runners of integration/, registered
ahead of the stock loaders (integration/Makefile).  They call runner_from_file(...)->generate(...) as upstream does -- and the forward runs on the GPU.

  * `tts-cli --prompt <text>` writes the WAV the Python binding's run_batch gives for the token ids the reference's front end produced (dumped by the runner);
  * the same prompt through the reference's own CPU build of the same cli.cpp (oracle/_ref/tts_cli_ref) has the same number of samples: phonemizer -> tokenizer ->
    chunking -> durations agree end to end (PCM values differ free-running at the reference's own build-to-build floor);
  * a multi-sentence prompt is chunked like the reference's (same sample count) and its chunks continue the reference's noise stream (b2tts_kokoro_run_chunks);
  * tts_b200_generate_batch (the one API addition) = the same prompts through generate() one after another;
  * perf_battery runs its 30 Harvard sentences.
(The batch-draining server worker and the patched HTTP server: tests/test_server_gpu.py.)"""
import os
import struct
import subprocess
import tempfile
import wave

import numpy as np
import pytest

from conftest import ROOT, report, rms, synth_gguf

pytestmark = pytest.mark.gpu
BUILD = os.path.join(ROOT, "integration", "_build")
SHORT = "hello world this is a test of the kokoro path"
LONG = ("It's easy to tell the depth of a well. The birch canoe slid on the smooth planks. Glue the sheet to the dark blue background and then these days "
        "a chicken leg is a rare dish.")


def _need(path):
    assert os.path.exists(path), f"{path} missing: built by `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists, shipped with the snapshot"
    return path


def _wav(path):
    with wave.open(path, "rb") as w:
        assert w.getframerate() == 24000 and w.getnchannels() == 1
        return np.frombuffer(w.readframes(w.getnframes()), np.int16 if w.getsampwidth() == 2 else np.int32)


def _cli(exe, gguf, prompt, out, env=None):
    r = subprocess.run([exe, "--model-path", gguf, "--prompt", prompt, "--save-path", out, "--n-threads", "16"], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0 and os.path.exists(out), (r.stdout[-1500:], r.stderr[-1500:])
    return r.stdout


@pytest.fixture(scope="module")
def text_gguf():
    return synth_gguf(text_vocab=True)


@pytest.mark.parametrize("prompt", [SHORT, LONG], ids=["one_chunk", "sentence_chunks"])
def test_unmodified_cli_runs_on_the_gpu_and_matches_run_batch(text_gguf, gpu_ctx, prompt):
    from tts_cpp_b200.binding import lib, runner_from_file
    import ctypes as C
    tmp = tempfile.mkdtemp(prefix="b2cli_")
    tokf, wavf, reff = os.path.join(tmp, "tok.txt"), os.path.join(tmp, "b200.wav"), os.path.join(tmp, "ref.wav")
    _cli(_need(os.path.join(BUILD, "tts-cli")), text_gguf, prompt, wavf, env={"B2TTS_DUMP_TOKENS": tokf})
    got = _wav(wavf)
    chunks = [[int(t) for t in ln.split()] for ln in open(tokf) if ln.strip()]
    assert len(chunks) >= (2 if prompt is LONG else 1) and all(c[0] == 0 and c[-1] == 0 for c in chunks)
    # the same tokens through the Python binding: chunks continue the noise stream (b2tts_kokoro_run_chunks)
    runner = runner_from_file(text_gguf, ctx=gpu_ctx)
    toks = np.concatenate([np.asarray(c, np.uint32) for c in chunks]); n = np.asarray([len(c) for c in chunks], np.int32)
    pcm = (C.POINTER(C.c_float) * len(chunks))(); ns = (C.c_int64 * len(chunks))()
    L = lib()
    L.b2tts_kokoro_run_chunks.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_char_p, C.c_uint64, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int64), C.c_void_p]
    rc = L.b2tts_kokoro_run_chunks(runner.h, len(chunks), toks.ctypes.data_as(C.POINTER(C.c_uint32)), n.ctypes.data_as(C.POINTER(C.c_int32)), None, 0, pcm, ns, None)
    assert rc == 0
    want = np.concatenate([np.ctypeslib.as_array(pcm[b], shape=(ns[b],)).copy() for b in range(len(chunks))])
    runner.close()
    assert got.shape == want.shape, (got.shape, want.shape)
    d = np.abs(got.astype(np.float64) / 32767.0 - np.clip(want, -1.0, 1.0)).max()
    print(f"PARITY drop-in cli ({len(chunks)} chunk(s), {got.size} samples): max |wav - run_chunks pcm| = {d:.3e} (16-bit WAV quantisation 3e-5)")
    assert d < 1e-4                                            # the WAV is the PCM of the same forward, to 16-bit quantisation
    # the reference's own CPU build of the same cli.cpp on the same prompt: same front end, same durations -> same number of samples
    _cli(_need(os.path.join(ROOT, "oracle", "_ref", "tts_cli_ref")), text_gguf, prompt, reff)
    ref = _wav(reff)
    r_got, r_ref = rms(got / 32767.0), rms(ref / 32767.0)
    print(f"PARITY drop-in cli vs the reference's CPU cli: samples {got.size} vs {ref.size}; rms {r_got:.4f} vs {r_ref:.4f}")
    assert got.size == ref.size
    assert 0.7 < r_got / r_ref < 1.4


def test_generate_batch_equals_sequential_generate(text_gguf):
    tmp = tempfile.mkdtemp(prefix="b2batch_")
    pf = os.path.join(tmp, "prompts.txt")
    prompts = ["hello world this is a test", "the quick brown fox jumps over the lazy dog", "a second runner starts a fresh noise stream"]
    open(pf, "w").write("\n".join(prompts) + "\n")
    pre = os.path.join(tmp, "o")
    r = subprocess.run([_need(os.path.join(BUILD, "batch_demo")), text_gguf, pf, pre], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-1500:])
    for i in range(len(prompts)):
        a, b = np.fromfile(f"{pre}.batch.{i}.f32", np.float32), np.fromfile(f"{pre}.single.{i}.f32", np.float32)
        d, rr, mx = report(f"generate_batch vs generate, prompt {i}", a, b)
        # same tokens, same noise offsets; batching changes tile shapes, hence summation order, hence a few fp16 re-roundings of activations (the reference's F16 path
        # re-rounds every conv input): measured 1.1e-3 relative on a B200
        assert a.shape == b.shape and a.size > 0 and d < 5e-3 * max(rr, 1e-6)


def test_unmodified_perf_battery_runs_on_the_gpu(text_gguf):
    r = subprocess.run([_need(os.path.join(BUILD, "perf_battery")), "--model-path", text_gguf], capture_output=True, text=True, timeout=900)
    print(r.stdout[-800:])
    assert r.returncode == 0 and "Mean Stats for arch kokoro" in r.stdout, r.stderr[-1500:]
