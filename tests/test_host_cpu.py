"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/b2tts.h declares, fails loudly
without a GPU (no CPU fallback), the synthetic-GGUF writer is deterministic, and the multi-GPU sharding logic works on a
world_size-2 gloo group."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_the_whole_abi():
    from tts_cpp_b200.abi import declared_symbols
    from tts_cpp_b200.binding import lib
    syms = declared_symbols()
    assert len(syms) >= 25 and "b2tts_kokoro_run_batch" in syms and "b2tts_op_conv_transpose_1d" in syms
    L = lib()
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from tts_cpp_b200.binding import Context, B2TTSError
    with pytest.raises(B2TTSError) as e:
        Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under tts_cpp_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "tts_cpp_b200")):
        if os.path.basename(dp) == "build":
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "kokoro_port" not in src, f


def test_synthetic_gguf_schema_and_determinism(tmp_path):
    from tts_cpp_b200.synth import kokoro_tensors, kokoro_metadata, synthetic_prompts, _f16_ok
    a = kokoro_tensors(seed=0)
    b = kokoro_tensors(seed=0)
    assert len(a) == 748 and sum(t.size for _, t in a) == 81271096          # the reference's Kokoro-82M tensor census
    assert all(n1 == n2 and np.array_equal(t1, t2) for (n1, t1), (n2, t2) in zip(a, b))
    names = dict(a)
    assert names["kokoro.decoder.generator.ups.0.weight"].shape == (512, 256, 20)
    # dtype policy mirrors quantize --convert-non-quantized-to-f16, except ConvTranspose kernels stay F32
    assert _f16_ok("kokoro.albert.layer.0.q") and not _f16_ok("kokoro.albert.embd") and not _f16_ok("kokoro.decoder.generator.ups.0.weight")
    assert not _f16_ok("kokoro.decoder.decoder_blocks.3.pool_weight") and not _f16_ok("kokoro.decoder.generator.resblocks.0.0.gamma1_weight")
    kv = dict(kokoro_metadata(128))
    assert kv["kokoro.decoder.generator.res_blocks.2.2.padding"] == 25 and kv["kokoro.decoder.generator.up_convs.0.stride"] == 10
    p = synthetic_prompts(3)
    assert all(len(u) == 66 and u[0] == 0 and u[-1] == 0 and 1 <= min(u[1:-1]) and max(u) <= 177 for u in p)


def test_plan_shards_is_a_balanced_partition():
    from tts_cpp_b200.sharding import plan_shards
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        n = [int(v) for v in rng.integers(3, 512, size=37)]
        sh = plan_shards(n, world)
        assert sorted(i for s in sh for i in s) == list(range(37))
        sizes = [len(s) for s in sh]
        assert max(sizes) - min(sizes) <= 1
        work = [sum(n[i] for i in s) for s in sh]
        assert max(work) - min(work) <= max(n)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from tts_cpp_b200.sharding import scatter_prompts, gather_pcm
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    prompts = None
    if rank == 0:
        rng = np.random.default_rng(5)
        prompts = [[0] + [int(v) for v in rng.integers(1, 178, size=int(n))] + [0] for n in rng.integers(1, 40, size=7)]
    idx, mine = scatter_prompts(dist, prompts)
    # stand-in for the per-rank forward: "PCM" of utterance = 10 samples per token, value = token id (deterministic, order-revealing)
    pcms = [np.repeat(np.asarray(p, np.float32), 10) for p in mine]
    out = gather_pcm(dist, idx, pcms, 7 if rank == 0 else 0)
    if rank == 0:
        ok = all(np.array_equal(o, np.repeat(np.asarray(p, np.float32), 10)) for o, p in zip(out, prompts))
        q.put(ok)
    dist.destroy_process_group()


def test_scatter_gather_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _worker_dev(rank, world, port, q):
    """the tensor-collective forms bench.py's strong-scaling step uses (broadcast of the packed prompts, all_gather of lengths, exact-length send / grouped irecv),
    here over gloo with CPU tensors"""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from tts_cpp_b200.sharding import scatter_tokens_nccl, gather_pcm_nccl
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    device = torch.device("cpu")
    counts = [4, 3]
    prompts = None
    if rank == 0:
        rng = np.random.default_rng(6)
        prompts = [[0] + [int(v) for v in rng.integers(1, 178, size=int(n))] + [0] for n in rng.integers(1, 40, size=7)]
    mine = scatter_tokens_nccl(dist, torch, device, prompts, counts, src=0)
    assert len(mine) == counts[rank]
    # stand-in for the forward: a padded [utterances][stride] block, utterance b = 10 samples per token with the token id as value
    ns = [10 * len(p) for p in mine]
    stride = max(ns) + 7
    block = torch.full((len(mine), stride), -1.0)
    for b, p in enumerate(mine):
        block[b, :ns[b]] = torch.from_numpy(np.repeat(np.asarray(p, np.float32), 10))
    got = gather_pcm_nccl(dist, torch, device, block, ns, counts, dst=0)
    if rank == 0:
        flat, lens = got
        want = np.concatenate([np.repeat(np.asarray(p, np.float32), 10) for p in prompts])
        q.put(bool(np.array_equal(flat.numpy(), want)) and lens == [10 * len(p) for p in prompts])
    else:
        assert got is None
    dist.destroy_process_group()


def test_device_scatter_gather_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_dev, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_polyphase_identity_of_the_strided_noise_conv():
    """The strided noise conv (K = 12, stride 6, pad 3) runs on the tcgen05 kernel as a stride-1 conv over rows of `stride` input
    frames (tts_cpp_b200/csrc/kokoro.cu, Kokoro::prepare).  This restates that re-indexing in numpy and checks it against the plain
    strided conv, for the model's shape and for shapes that exercise pad % stride == 0 and K not a multiple of the stride."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(3)
    for (K, s, pad, Cin, N, L) in ((12, 6, 3, 22, 8, 121), (12, 6, 6, 5, 3, 66), (7, 3, 2, 4, 2, 40), (10, 5, 0, 3, 2, 35)):
        x = rng.standard_normal((Cin, L)).astype(np.float64)
        w = rng.standard_normal((N, Cin, K)).astype(np.float64)
        want = F.conv1d(torch.from_numpy(x)[None], torch.from_numpy(w), None, stride=s, padding=pad)[0].numpy()
        sh = (s - pad % s) % s
        Kp = (K + sh + s - 1) // s
        padp = pad // s + (1 if sh else 0)
        Lp = -(-L // s) * s                       # per-utterance pitch rounded up to the stride, rows past the end are zeros
        xp = np.zeros((Cin, Lp)); xp[:, :L] = x
        X = xp.reshape(Cin, Lp // s, s)           # X[c][r][j] = x[c][r*s + j]
        wp = np.zeros((N, Kp, s, Cin))
        for m in range(Kp):
            for j in range(s):
                k = m * s + j - sh
                if 0 <= k < K:
                    wp[:, m, j, :] = w[:, :, k]
        Lout = want.shape[1]
        got = np.zeros((N, Lout))
        for t in range(Lout):
            for m in range(Kp):
                r = t - padp + m
                if 0 <= r < Lp // s:
                    got[:, t] += np.einsum("njc,cj->n", wp[:, m], X[:, r, :])
        assert np.allclose(got, want, atol=1e-10), (K, s, pad)


def test_reference_side_binding_type_checks_against_the_reference_headers():
    """integration/kokoro_b200_runner.cpp (the tts_generation_runner subclass + loader a TTS.cpp maintainer adds) must compile against the
    reference's own headers and include/b2tts.h.  Only possible where the reference checkout exists (the build container)."""
    import shutil
    import subprocess
    if not os.path.isdir("/root/reference/src") or shutil.which("g++") is None:
        pytest.skip("reference checkout not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["make", "-C", os.path.join(root, "oracle"), "binding_check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_batch_draining_worker_selftest():
    """integration/b200_batch_worker.h behind a server-shaped queue with the reference's own dummy_runner: compatible tasks share one forward,
    the others keep their place and order, max_batch, timed-out tasks, the one-by-one fallback, task-owned PCM.  The binary is built by build() where the reference
    checkout exists (integration/Makefile) and needs no GPU."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "integration", "_build", "worker_demo")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/worker_demo not built (no reference checkout)")
    r = subprocess.run([exe, "selftest"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "worker selftest OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


def test_batch_draining_worker_two_workers_under_thread_sanitizer():
    """Four producers and two workers (each with its own runner) on ONE queue, max_batch 8, mixed voices / models / non-TTS tasks: every one of 400 tasks is answered
    exactly once with its own audio -- natively and under ThreadSanitizer (worker_demo_tsan: the worker header and the queue are instrumented; no report allowed)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for exe, env in (("worker_demo", {}), ("worker_demo_tsan", {"TSAN_OPTIONS": "halt_on_error=1 exitcode=66"})):
        path = os.path.join(root, "integration", "_build", exe)
        if not os.path.exists(path):
            pytest.skip(f"integration/_build/{exe} not built (no reference checkout)")
        r = subprocess.run([path, "stress"], capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
        assert r.returncode == 0 and "worker stress OK: 400 tasks" in r.stdout and "ThreadSanitizer" not in r.stderr, (exe, r.stdout[-500:], r.stderr[-2000:])


def test_server_patch_fails_loudly_when_upstream_moves(tmp_path):
    """integration/patch_server.py applies three edits to the reference's server.cpp at build time; each must match exactly once.  A server.cpp whose worker loop
    reads differently must stop the build with a message, not produce a server that silently keeps the one-task-at-a-time loop."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = "/root/reference/examples/server/server.cpp"
    if not os.path.exists(src):
        pytest.skip("reference checkout not present")
    html = "/root/reference/examples/server/public/index.html"
    script = os.path.join(root, "integration", "patch_server.py")
    ok = subprocess.run([sys.executable, script, src, str(tmp_path / "ok.cpp"), html, str(tmp_path / "ok.hpp")], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stderr
    out = open(tmp_path / "ok.cpp").read()
    assert out.count("b200::batch_loop(") == 1 and out.count("b200::release(rtask)") == 1 and "task_queue->get_next()" not in out
    moved = open(src).read().replace("struct simple_server_task * task = task_queue->get_next();", "auto * task = task_queue->next();")
    (tmp_path / "moved.cpp").write_text(moved)
    bad = subprocess.run([sys.executable, script, str(tmp_path / "moved.cpp"), str(tmp_path / "bad.cpp"), html, str(tmp_path / "bad.hpp")], capture_output=True, text=True)
    assert bad.returncode != 0 and "worker::loop" in bad.stderr and not os.path.exists(tmp_path / "bad.cpp"), (bad.returncode, bad.stderr)


def test_patched_reference_server_serves_http_with_the_batch_worker():
    """The reference's examples/server/server.cpp with the three edits of INTEGRATION.md section 5 applied at build time (integration/patch_server.py), over real HTTP
    with the reference's `test:dummy` model (one second of a per-character tone per prompt character, src/models/dummy/model.cpp): every concurrent request gets ITS
    OWN audio (the worker hands every task a copy; upstream hands out a pointer into the runner's buffer), and every task went through b200::batch_loop."""
    from conftest import patched_server
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "integration", "_build", "tts-server-b200")):
        pytest.skip("integration/_build/tts-server-b200 not built (no reference checkout)")
    prompts = ["a", "bb", "ccc", "dd", "e", "ff", "ggg", "h", "ab", "ba"]
    with patched_server("test:dummy", max_batch=4) as srv:
        out = srv.speech(prompts, threads=10)
        fw = srv.forwards()
    sr = 44100
    j = np.arange(sr, dtype=np.float32)
    for p, (code, pcm, rate) in zip(prompts, out):
        assert code == 200 and rate == sr and pcm.size == len(p) * sr, (p, code, rate, pcm.size)
        for i, ch in enumerate(p):                             # the dummy model's tone for this character: 16-bit WAV of sin(j pi / sr) * sin(j / wavelength)
            wl = np.float32(np.float32(sr / np.pi / 2) / np.float32(200 + ord(ch)))
            want = np.sin(j * np.float32(np.pi / sr)) * np.sin(j / wl)
            got = pcm[i * sr:(i + 1) * sr].astype(np.float32) / 32768.0
            assert np.abs(got - want).max() < 2e-3, (p, i, float(np.abs(got - want).max()))
    assert sum(fw) == len(prompts) and max(fw) <= 4, fw       # every task through the batch loop, never more than max_batch per forward


def test_delay_pattern_undo_matches_the_reference_indexing():
    """ar_host.{parler,dia}_adjust_output_tokens against a literal restatement of the reference's flat-index loops
    (parler model.cpp:734-760, dia model.cpp:825-847) on random token streams with special ids mixed in."""
    from tts_cpp_b200.ar_host import DIA_DELAY_PATTERN, dia_adjust_output_tokens, parler_adjust_output_tokens
    rng = np.random.default_rng(3)
    H, V = 9, 1024
    for steps in (8, 9, 10, 40):
        t = rng.integers(0, 1040, size=(steps, H))          # ~1.5 % special ids per token
        flat = t.reshape(-1)
        want = []
        for i in range(steps):                               # parler: next_index = i*H + ii*H + ii
            idx = [i * H + ii * H + ii for ii in range(H)]
            if any(j >= flat.size or flat[j] >= V for j in idx):
                continue
            want.append([flat[j] for j in idx])
        got = parler_adjust_output_tokens(t, V)
        assert got.shape == (len(want), H) and (len(want) == 0 or np.array_equal(got, np.asarray(want)))
    for steps in (15, 16, 30, 80):
        t = rng.integers(0, 1030, size=(steps, H))
        flat = t.reshape(-1)
        want = []
        for i in range(steps - 15):                          # dia: next_index = i*H + delay[ii]*H + ii
            idx = [i * H + DIA_DELAY_PATTERN[ii] * H + ii for ii in range(H)]
            if any(flat[j] >= V for j in idx):
                continue
            want.append([flat[j] for j in idx])
        got = dia_adjust_output_tokens(t, V)
        assert got.shape == (len(want), H) and (len(want) == 0 or np.array_equal(got, np.asarray(want)))


def test_orpheus_stream_to_snac_codes():
    """ar_host.orpheus_{n_generated,prepare_output_tokens}: the stopping rule and the 7-token frame -> three SNAC levels mapping (orpheus model.cpp:371-398)."""
    from tts_cpp_b200.ar_host import orpheus_n_generated, orpheus_prepare_output_tokens
    base = 128266
    frames = 5
    stream = []
    for f in range(frames):
        stream += [base + ii * 4096 + (100 * f + ii) for ii in range(7)]
    stream_stop = stream + [128258, 11, 12]
    assert orpheus_n_generated(stream_stop) == len(stream) + 1 and orpheus_n_generated(stream) == len(stream) and orpheus_n_generated(stream, max_generation=9) == 9
    kept = np.asarray(stream_stop[:orpheus_n_generated(stream_stop)])
    c, m, fine = orpheus_prepare_output_tokens(kept)                  # the trailing stop token does not fill a frame and is dropped
    assert c.tolist() == [100 * f for f in range(frames)]
    assert m.tolist() == sum(([100 * f + 1, 100 * f + 4] for f in range(frames)), [])
    assert fine.tolist() == sum(([100 * f + 2, 100 * f + 3, 100 * f + 5, 100 * f + 6] for f in range(frames)), [])
    assert len(m) == 2 * len(c) and len(fine) == 4 * len(c)             # the L/4, L/2, L layout snac_runner::run / b2tts_snac_decode_batch takes


def test_scripts_parse():
    """scripts/ is what the first GPU call of the next round runs: the shell script must parse and the Python ones must compile (no GPU needed for either)."""
    import glob
    import py_compile
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for sh in glob.glob(os.path.join(root, "scripts", "*.sh")):
        assert subprocess.run(["bash", "-n", sh]).returncode == 0, sh
    for py in glob.glob(os.path.join(root, "scripts", "*.py")) + [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py"), os.path.join(root, "profiles", "static_ar_summary.py")]:
        py_compile.compile(py, doraise=True)
