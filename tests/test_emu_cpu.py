"""CPU: the LOGIC of the plain CUDA-core kernels, checked in the GPU-less container.

tests/emu compiles an unmodified .cu file of the library against a functional CPU emulation of the CUDA subset it uses (blocks run one
after another, the threads of a block are fibers, __syncthreads / warp shuffles have their real meaning) and runs the library's own host
code on top.  That proves indexing, reductions, barriers and orchestration against the reference's golden vectors; it proves nothing about
the hardware (coalescing, occupancy, tcgen05 / TMA / mbarrier code, which is not emulated) -- the -m gpu tests remain the parity gate."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import emu_build  # noqa: E402

from tts_cpp_b200.synth import cached_dia_gguf, cached_orpheus_gguf, cached_parler_gguf  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

def _T5_TOL(case):
    """relative RMS bar of a T5 golden case: F32 1e-4 (fp16 GELU table); F16 1.5e-3 (fp16 activation rounding); block-quantised 3e-2 -- ggml_mul_mat re-quantises the
    activations to Q8_0 per 32 columns, so 1e-7 of summation-order noise moves whole quantisation steps (two correct implementations: 1e-7 on one prompt, 1e-2 on the next)"""
    return 3e-2 if case.startswith("q") else 1.5e-3 if case.startswith("f16") else 1e-4



AR_SOURCES = ["orpheus.cu", "parler.cu", "dia.cu", "t5.cu", "pdk.cu", "sampler.cu"]
# The library's defaults are the fast paths (persistent decode kernel for F16 Parler, tensor-core GEMV for F16 matrices, CUDA-graph replay); most tests below are about
# one specific kernel family, so they start from the plain launch-per-op configuration and switch on what they test.
EMU_DEFAULTS = {"B2TTS_AR_PDK": "0", "B2TTS_AR_MMA": "0", "B2TTS_AR_GRAPH": "0"}


# independent emulator runs of one test go to the host's other cores (each is a single-threaded subprocess; subprocess.run releases the GIL): same runs, same checks
from concurrent.futures import ThreadPoolExecutor  # noqa: E402
_POOL = ThreadPoolExecutor(3)


def _bg(fn, *a, **k):
    return _POOL.submit(fn, *a, **k)


def _run_ar(tmp_path, model, gguf_path, prompts, steps, tag, env=None, want_stderr=False):
    """-> (tokens [B][steps][W], logits [B][steps][V]) from the library's generate_greedy under emulation"""
    exe = emu_build.build("ar_emu", AR_SOURCES, ["ar_main.cpp", os.path.join(emu_build.CSRC, "gguf_reader.cpp")])
    pin, pout = str(tmp_path / f"p{tag}.bin"), str(tmp_path / f"o{tag}.bin")
    with open(pin, "wb") as f:
        f.write(struct.pack("ii", len(prompts), steps))
        for p in prompts:
            f.write(struct.pack("i", p.size))
            f.write(np.asarray(p, np.uint32).tobytes())
    r = subprocess.run([exe, model, gguf_path, pin, pout], capture_output=True, text=True, timeout=900, env={**os.environ, **EMU_DEFAULTS, **(env or {})})
    assert r.returncode == 0, r.stderr[-2000:]
    raw = open(pout, "rb").read()
    W, V = struct.unpack("ii", raw[:8])
    n = len(prompts) * steps
    tok = np.frombuffer(raw[8:8 + n * W * 4], np.int32).reshape(len(prompts), steps, W)
    logits = np.frombuffer(raw[8 + n * W * 4:], np.float32).reshape(len(prompts), steps, V)
    return (tok, logits, r.stderr) if want_stderr else (tok, logits)


def _run_orpheus(tmp_path, prompts, steps, tag, env=None):
    tok, logits = _run_ar(tmp_path, "orpheus", cached_orpheus_gguf(seed=0), prompts, steps, tag, env=env)
    return tok[:, :, 0], logits


def test_orpheus_cuda_path_emulated_matches_reference_tokens_and_logits(tmp_path):
    """Orpheus::generate_greedy (orpheus.cu: prefill of a ragged batch, KV cache append, GQA attention, argmax feedback) under emulation
    against the token ids and logits the compiled unmodified reference produced (tests/golden/orpheus_vectors.npz)."""
    g = np.load(os.path.join(GOLD, "orpheus_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].size)
    tok, logits = _run_orpheus(tmp_path, prompts, steps, "b")
    for u in range(2):
        d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
        print(f"PARITY(emulated) orpheus prompt {u}: tokens {tok[u].tolist()}  max |logit diff| {d:.3e}")
        assert np.array_equal(tok[u], g[f"tokens{u}"])
        assert d < 1e-3      # same tolerance as tests/test_orpheus_gpu.py (measured 4.5e-6: fp32 throughout, summation order differs)
    single, _ = _run_orpheus(tmp_path, [prompts[1]], steps, "s")
    assert np.array_equal(single[0], tok[1])         # batching does not change a sequence
    plain, lp = _run_orpheus(tmp_path, prompts, steps, "p", env={"B2TTS_AR_ATT": "plain"})      # one block per query head instead of per kv-head group
    assert np.array_equal(plain, tok) and float(np.abs(lp - logits).max()) < 1e-4
    rev, lr = _run_orpheus(tmp_path, prompts, steps, "r", env={"B2EMU_REVERSE": "1"})             # threads scheduled in descending order: a result that depends on
    assert np.array_equal(rev, tok) and np.array_equal(lr, logits)                                  # the order threads reach a barrier-free region (a race) would differ
    for gn in (1, 2, 4):                              # output rows per warp of the plain GEMV: the per-output summation order does not depend on it
        t2, l2 = _run_orpheus(tmp_path, prompts, steps, f"g{gn}", env={"B2TTS_GEMV_GN": str(gn)})
        assert np.array_equal(t2, tok) and np.array_equal(l2, logits)


@pytest.mark.parametrize("f16", [False, True], ids=["f32", "f16"])
def test_parler_cuda_path_emulated_matches_reference_tokens_and_logits(tmp_path, f16):
    """Parler::generate_greedy (parler.cu: prompt pass, delay-pattern codebook embedding, causal self-attention over the cache, cross-attention
    over the stored text encoding, GELU table, nine heads + per-head argmax) under emulation against tests/golden/parler[_f16]_vectors.npz.
    f16: the GGUF `quantize --quantized-type F16` writes -- the decoder matrices are F16 and every product with one rounds its activations to fp16."""
    g = np.load(os.path.join(GOLD, "parler_f16_vectors.npz" if f16 else "parler_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    tok, logits = _run_ar(tmp_path, "parler", cached_parler_gguf(seed=0, f16=f16), prompts, steps, "b")
    for u in range(2):
        ref = g[f"logits{u}"].reshape(steps, -1)
        d = float(np.abs(logits[u] - ref).max())
        print(f"PARITY(emulated) parler {'f16' if f16 else 'f32'} prompt {u}: tokens {tok[u].tolist()}  max |logit diff| {d:.3e}  (logit std {ref.std():.2f})")
        assert np.array_equal(tok[u], g[f"tokens{u}"])
        # F32: ggml's GELU is an fp16 table -- an activation that lands on the other side of a rounding boundary moves a logit by ~1e-3.
        # F16: every matrix input is rounded to fp16 as well, so summation-order noise flips roundings in all 64 products: ~7e-3 at a logit std of 4
        assert d < (3e-2 if f16 else 1e-2)


@pytest.mark.parametrize("model,kind", [("parler", "f32"), ("parler", "f16_mma"), ("parler", "q5_0"), ("orpheus", "f32"), ("dia", "f16"), ("dia", "q8_0")])
def test_fused_launches_bit_identical(tmp_path, model, kind):
    """The fused launches of the decode step (default) against B2TTS_AR_FUSE=0: q / k / v (gate / up, Dia's cross k / v) as ONE grouped GEMV launch; for Parler the
    k / v rows also go straight into the cache and GELU sits in fc1's epilogue (three launches + store_kv_kernel + gelu_f16lut_kernel before).  The per-output
    arithmetic is the same, so tokens AND logits must be bit-identical for every storage kind of the matrices, with the expected number of launches saved."""
    import re
    gname = {"f32": f"{model}_vectors.npz", "f16": f"{model}_f16_vectors.npz", "f16_mma": f"{model}_f16_vectors.npz", "q5_0": f"{model}_vectors.npz", "q8_0": f"{model}_vectors.npz"}[kind]
    g = np.load(os.path.join(GOLD, gname))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    cached = {"parler": cached_parler_gguf, "orpheus": cached_orpheus_gguf, "dia": cached_dia_gguf}[model]
    gguf = cached(seed=0, quant=kind.upper()) if kind.startswith("q") else (cached(seed=0, f16=True) if kind.startswith("f16") else cached(seed=0))
    env = {"B2TTS_AR_MMA": "1"} if kind == "f16_mma" else {}
    tok_f, log_f, err_f = _run_ar(tmp_path, model, gguf, prompts, steps, "fu", env=env, want_stderr=True)
    tok_u, log_u, err_u = _run_ar(tmp_path, model, gguf, prompts, steps, "un", env=dict(env, B2TTS_AR_FUSE="0"), want_stderr=True)
    assert np.array_equal(tok_f, tok_u) and np.array_equal(log_f.view(np.uint32), log_u.view(np.uint32))
    n_f, n_u = (int(re.search(r"(\d+) launches", e).group(1)) for e in (err_f, err_u))
    print(f"PARITY(emulated) {model} {kind}: fused == unfused bit for bit; launches {n_u} -> {n_f}")
    want = {"parler": 4 * 8 * (steps + 1),                        # 8 layers x (decode passes + the prompt pass) x (q/k/v: 2, KV store: 1, GELU: 1)
            "orpheus": 3 * 2 * steps,                              # 2 layers x passes (step 0 is the prompt pass) x (q/k/v: 2, gate/up: 1)
            "dia": 5 * 2 * steps + 3 * 2 + 2}[model]               # 2 decoder layers x steps x (q/k/v: 2, gate/up: 1, RoPE q + RoPE k + KV store as one: 2), 2 encoder layers x 3 once, cross k/v of 2 decoder layers once
    assert n_u - n_f == want, (n_u, n_f, want)


@pytest.mark.parametrize("model", ["parler", "dia"])
def test_tensor_core_gemv_emulated_f16(tmp_path, model):
    """B2TTS_AR_MMA=1: F16 matrices through gemv_mma_kernel<false> (mma.sync.m16n8k16 with the batch as M, K split over the warps of a block, the k index permuted
    consistently on both operands).  Under emulation the instruction is a functional model of the PTX fragment layout; the tokens must be the F16 reference's."""
    g = np.load(os.path.join(GOLD, f"{model}_f16_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    gguf = cached_parler_gguf(seed=0, f16=True) if model == "parler" else cached_dia_gguf(seed=0, f16=True)
    tok, logits = _run_ar(tmp_path, model, gguf, prompts, steps, "m", env={"B2TTS_AR_MMA": "1"})
    for u in range(2):
        ref = g[f"logits{u}"].reshape(steps, -1)
        rms = float(np.sqrt(((logits[u] - ref) ** 2).mean()))
        print(f"PARITY(emulated, mma gemv) {model} f16 prompt {u}: logit diff rms {rms:.3e} max {float(np.abs(logits[u] - ref).max()):.3e}")
        assert np.array_equal(tok[u], g[f"tokens{u}"])
        assert rms < (0.1 if model == "dia" else 5e-3)


@pytest.mark.parametrize("model", ["orpheus", "parler", "dia"])
def test_cuda_graph_replay_emulated(tmp_path, model):
    """B2TTS_AR_GRAPH=1: one decode step is captured into a CUDA graph and replayed, the step number being device-resident.  Under emulation a capture
    records the launches (closures owning their arguments, like kernel parameters) and cudaGraphLaunch replays them: same tokens as the reference."""
    g = np.load(os.path.join(GOLD, f"{model}_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    gguf = {"orpheus": cached_orpheus_gguf, "parler": cached_parler_gguf, "dia": cached_dia_gguf}[model](seed=0)
    tok, _, err = _run_ar(tmp_path, model, gguf, prompts, steps, "g", env={"B2TTS_AR_GRAPH": "1", "B2EMU_NO_LOGITS": "1"}, want_stderr=True)
    replays = steps - 2 if model == "orpheus" else steps - 1      # the first decode step runs directly (Orpheus' step 0 is the prompt pass itself)
    assert f"{replays} graph replays" in err, err[-300:]
    for u in range(2):
        assert np.array_equal(tok[u].reshape(g[f"tokens{u}"].shape), g[f"tokens{u}"])


@pytest.mark.parametrize("f16", [False, True], ids=["f32", "f16"])
def test_dia_cuda_path_emulated_matches_reference_tokens_and_logits(tmp_path, f16):
    """Dia::generate_greedy (dia.cu: two-sequence encoder pass with its block mask, RoPE'd cross keys for the prompt positions only, GQA decoder
    self-attention, cross-attention, cfg_scale, nine heads + per-head argmax) under emulation against tests/golden/dia[_f16]_vectors.npz.
    f16: the quantize tool's F16 GGUF (every matrix but the output heads F16, activations rounded to fp16 before each such product)."""
    g = np.load(os.path.join(GOLD, "dia_f16_vectors.npz" if f16 else "dia_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    tok, logits = _run_ar(tmp_path, "dia", cached_dia_gguf(seed=0, f16=f16), prompts, steps, "b")
    for u in range(2):
        ref = g[f"logits{u}"].reshape(steps, -1)
        d = float(np.abs(logits[u] - ref).max())
        rms = float(np.sqrt(((logits[u] - ref) ** 2).mean()))
        print(f"PARITY(emulated) dia {'f16' if f16 else 'f32'} prompt {u}: tokens {tok[u].tolist()}  logit diff max {d:.3e} rms {rms:.3e}  (logit std {ref.std():.2f})")
        assert np.array_equal(tok[u], g[f"tokens{u}"])
        # F32: CFG multiplies the fp32 summation-order noise of both passes by 4 at a logit std of ~13.
        # F16: rounding-boundary flips in every product, a softmax without 1/sqrt(d) and the 4x CFG gain: the reference's own F16 and F32 logits differ by
        # 0.03-0.19 RMS; the restated rounding model stays within 0.05 RMS of the F16 reference (and reproduces its tokens, which differ from the F32 ones)
        assert (rms < 0.1 and d < 1.0) if f16 else d < 2e-2


def test_dia_cuda_path_emulated_check_stopping(tmp_path):
    """generate_from_batch's whole loop: the end-of-stream countdown (EOS / PAD injected along the delay pattern from position max_generation - max_delay)
    and the frame count at which check_stopping ends it, against the reference run to completion (tests/golden/dia_stop_vectors.npz: 63 frames)."""
    g = np.load(os.path.join(GOLD, "dia_stop_vectors.npz"))
    ref = g["tokens0"]
    cap = int(g["step_cap"])
    tok, logits = _run_ar(tmp_path, "dia", cached_dia_gguf(seed=0), [g["prompt0"]], cap, "stop")
    n_gen = ref.shape[0]
    assert n_gen < cap
    assert np.array_equal(tok[0, :n_gen], ref)
    assert not tok[0, n_gen:].any() and not logits[0, n_gen:].any()        # rows past the stop are zero
    d = float(np.abs(logits[0, n_gen - 1] - g["logits_last0"].reshape(-1)).max())
    print(f"PARITY(emulated) dia run to check_stopping: {n_gen} frames, last-frame max |logit diff| {d:.3e}")
    assert d < 2e-2


SAMPLER_CFGS = {"greedy": dict(do_sample=0, temperature=1.0, top_k=0, top_p=1.0, rp=1.0),
                "default_top50": dict(do_sample=1, temperature=1.0, top_k=50, top_p=1.0, rp=1.0),      # the reference's default generation_configuration
                "temp_rep": dict(do_sample=1, temperature=0.7, top_k=20, top_p=1.0, rp=1.3),
                "topk_topp": dict(do_sample=1, temperature=1.3, top_k=40, top_p=0.9, rp=1.0),
                "topp_only": dict(do_sample=1, temperature=0.9, top_k=0, top_p=0.8, rp=1.1),
                "full_vocab": dict(do_sample=1, temperature=1.1, top_k=0, top_p=1.0, rp=1.2),
                "top1000_flat": dict(do_sample=1, temperature=4.0, top_k=1000, top_p=1.0, rp=1.0)}      # draws land deep inside a 1 000-entry nucleus


def _sampler_case(shape):
    """(rows, V, steps, logits) of a sampler test: `small` = codebook-sized rows; `wide_ties` = a wide vocabulary on a coarse grid of values (signed zeros
    included), so that the nucleus boundary falls inside a run of exactly equal logits -- the radix select's ordered tie path; `wide` = wide, no ties."""
    rng = np.random.default_rng(33)
    if shape == "small":
        rows, V, steps = 5, 300, 6
        logits = (rng.standard_normal((steps, rows, V)) * 2.5).astype(np.float32)
        logits[:, :, 7] += 6.0                                # a dominant token: repeated picks exercise the repetition counts
        logits[2:, 1, 40] = logits[2:, 1, 41]                  # exact ties inside the nucleus: lower id first
    else:
        rows, V, steps = 8, 20011, 6
        logits = (rng.standard_normal((steps, rows, V)) * 2.0).astype(np.float32)
        if shape.startswith("wide_ties"):
            logits = (np.round(logits * 2.0) / 2.0).astype(np.float32)      # multiples of 0.5: hundreds of equal entries per value, +0.0 and -0.0 both present
            assert np.signbit(logits[logits == 0]).any() and not np.signbit(logits[logits == 0]).all()
        logits[:, :, 11] += 3.0
        if shape == "wide_ties_peaked":                        # a 0.8 nucleus of a few entries (one chunk of picks); "wide_ties" / "wide" with top-p only need thousands: several chunks
            logits = (np.round(logits * 2.0) * 1.5).astype(np.float32)
    return rows, V, steps, logits


@pytest.mark.parametrize("name,shape", [(n, "small") for n in SAMPLER_CFGS if n != "top1000_flat"] + [("top1000_flat", "wide_ties"), ("top1000_flat", "wide")] +
                         [("default_top50", "wide_ties"), ("temp_rep", "wide_ties"), ("topk_topp", "wide_ties"), ("topp_only", "wide_ties_peaked"), ("topp_only", "wide_ties"), ("topp_only", "wide"), ("default_top50", "wide")])
def test_sampler_kernel_emulated_matches_port(tmp_path, name, shape):
    """sample_rows (sampler.cu) under emulation against oracle/sampler_port.py (itself pinned to the reference sampler) over consecutive steps with the
    repetition state carried along: the port is fed the same uniforms the kernel derives from (seed, row, step); tokens and state must be identical."""
    sys.path.insert(0, ROOT)
    from oracle.sampler_port import SamplerPort
    cfg = SAMPLER_CFGS[name]
    exe = emu_build.build("sampler_emu", ["sampler.cu"], ["sampler_main.cpp"])
    seed = 0x1234ABCD5678
    rows, V, steps, logits = _sampler_case(shape)
    pin, pout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(pin, "wb") as f:
        f.write(struct.pack("<iiiifffQi", rows, V, cfg["do_sample"], cfg["top_k"], cfg["top_p"], cfg["temperature"], cfg["rp"], seed, steps))
        f.write(np.full(rows, -1, np.int32).tobytes()); f.write(np.zeros(rows, np.int32).tobytes()); f.write(logits.tobytes())
    r = subprocess.run([exe, pin, pout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    raw = open(pout, "rb").read()
    if shape != "small":
        # the same run with the threads of a block scheduled in descending order (the slots of the unordered collection change, the result must not), and under
        # AddressSanitizer + the alignment sanitizer (shared-memory pick / histogram arrays are statics there: an index past their end is reported)
        r = subprocess.run([exe, pin, pout + ".rev"], capture_output=True, text=True, timeout=600, env=dict(os.environ, B2EMU_REVERSE="1"))
        assert r.returncode == 0, r.stderr[-2000:]
        assert open(pout + ".rev", "rb").read() == raw
        if name in ("default_top50", "topp_only"):
            exe_asan = emu_build.build("sampler_emu_asan", ["sampler.cu"], ["sampler_main.cpp"], asan=True)
            r = subprocess.run([exe_asan, pin, pout + ".asan"], capture_output=True, text=True, timeout=900,
                               env=dict(os.environ, ASAN_OPTIONS="detect_stack_use_after_return=0:detect_leaks=0"))      # (the fiber stacks are a pool kept until exit)
            assert r.returncode == 0, r.stderr[-3000:]
            assert open(pout + ".asan", "rb").read() == raw
    n = steps * rows
    toks = np.frombuffer(raw, np.int32, n, 0).reshape(steps, rows)
    us = np.frombuffer(raw, np.float32, n, 4 * n).reshape(steps, rows)
    last = np.frombuffer(raw, np.int32, rows, 8 * n); counts = np.frombuffer(raw, np.int32, rows, 8 * n + 4 * rows)
    assert us.min() >= 0.0 and us.max() < 1.0 and len(np.unique(us)) == n
    from oracle.sampler_port import uniform_from_counter
    assert all(us[s_, r_] == uniform_from_counter(seed, r_, s_) for s_ in range(steps) for r_ in range(rows))      # the Python mirror of the counter hash
    port = SamplerPort(rows, V, cfg["temperature"], cfg["top_k"], cfg["top_p"], cfg["rp"])
    for s in range(steps):
        if cfg["do_sample"]:
            want = port.draw(logits[s], us[s])
        else:
            want = np.array([int(np.argmax(port._eff(logits[s][i], i))) for i in range(rows)])
        assert np.array_equal(toks[s], want), f"step {s}: {toks[s]} vs {want}"
    if cfg["do_sample"] and cfg["rp"] != 1.0:
        assert np.array_equal(last, port.last) and np.array_equal(counts, port.counts)


def test_parler_sampling_loop_emulated_matches_port(tmp_path):
    """Parler::generate with the reference's default sampler settings (top_k 50, temperature 1) under emulation, against oracle/parler_port.py stepping with
    oracle/sampler_port.py on the same uniforms: the whole loop -- logits, nucleus, draw, delay-pattern feedback of SAMPLED tokens -- must give the same ids."""
    sys.path.insert(0, ROOT)
    from oracle.parler_port import ParlerPort
    from oracle.sampler_port import SamplerPort, uniform_from_counter
    import torch
    g = np.load(os.path.join(GOLD, "parler_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps, seed, top_k, temp = 5, 77, 50, 1.0
    tok, _ = _run_ar(tmp_path, "parler", cached_parler_gguf(seed=0), prompts, steps, "smp", env={"B2EMU_SAMPLE": f"{top_k} 1.0 {temp} 1.0 {seed}"})
    port = ParlerPort(cached_parler_gguf(seed=0))
    H = port.n_out
    for u, prompt in enumerate(prompts):
        samp = SamplerPort(H, port.vocab, temp, top_k, 1.0, 1.0)
        port.reset()
        t = torch.from_numpy(np.asarray(prompt).astype(np.int64))
        port.step(port.w["embed_prompts"][t] + port.w["positional_embed"][torch.arange(t.numel())])
        last, want = None, []
        for s in range(steps):
            ids = [int(last[i]) if s > i else port.bos for i in range(H)]
            x = None
            for i in range(H):
                e = port.w[f"embed_tokens.{i}.weight"][ids[i]]
                x = e if x is None else e + x
            lg = port.step((x + port.w["positional_embed"][port.pos])[None, :])[:, 0, :].numpy()
            us = np.array([uniform_from_counter(seed, u * H + i, s) for i in range(H)], np.float32)      # row = sequence * heads + head
            last = samp.draw(lg, us)
            want.append(last.astype(np.int32))
        assert np.array_equal(tok[u], np.stack(want)), f"prompt {u}: {tok[u].tolist()} vs {np.stack(want).tolist()}"
    assert not np.array_equal(tok[0], g["tokens0"])          # and it is not the greedy stream


@pytest.mark.parametrize("case", ["all_eos", "max_generation"])
def test_parler_stop_rule_emulated(tmp_path, case):
    """Parler::generate with n_generated (the reference's stop rule on the device: eos_seen per head, check_stopping per sequence) under emulation against the
    reference run to completion (tests/golden/parler_stop_vectors.npz): same frame count, same tokens, zero rows past the stop."""
    g = np.load(os.path.join(GOLD, "parler_stop_vectors.npz"))
    ref = g[f"{case}.tokens"]
    prompt, boost = g[f"{case}.prompt"], float(g[f"{case}.boost"])
    cap = ref.shape[0] + 8                                     # a few steps past the stop: the early exit (stop flags read back every 4 steps here) skips most of them
    exe = emu_build.build("ar_emu", AR_SOURCES, ["ar_main.cpp", os.path.join(emu_build.CSRC, "gguf_reader.cpp")])
    pin, pout = str(tmp_path / "p.bin"), str(tmp_path / "o.bin")
    with open(pin, "wb") as f:
        f.write(struct.pack("ii", 1, cap)); f.write(struct.pack("i", prompt.size)); f.write(prompt.astype(np.uint32).tobytes())
    r = subprocess.run([exe, "parler", cached_parler_gguf(seed=0, eos_boost=boost), pin, pout], capture_output=True, text=True, timeout=900,
                       env={**os.environ, **EMU_DEFAULTS, "B2EMU_STOP": "1", "B2EMU_NO_LOGITS": "1", "B2TTS_AR_EXIT_EVERY": "4"})
    assert r.returncode == 0, r.stderr[-2000:]
    launches = int(r.stderr.split("emulated ")[1].split(" launches")[0])
    # the stop flags are read back every 4 steps here, so the batch stops stepping at most 4 steps after every sequence has ended: 16 launches of prepare, 121 of
    # the prompt pass, 126 per audio step of the 8-layer test model -- without the early exit all `cap` = frames + 8 steps would run
    assert launches <= 16 + 121 + (ref.shape[0] + 5) * 126, launches
    raw = open(pout, "rb").read()
    W, V = struct.unpack("ii", raw[:8])
    tok = np.frombuffer(raw, np.int32, cap * W, 8).reshape(cap, W)
    n_gen = int(np.frombuffer(raw[-4:], np.int32)[0])
    assert n_gen == ref.shape[0]
    assert np.array_equal(tok[:n_gen], ref) and not tok[n_gen:].any()


@pytest.mark.parametrize("mma", [False, True], ids=["plain", "split_mma"])
def test_orpheus_wide_emulated(tmp_path, mma):
    """Orpheus with head size 128 (hidden 768: every matrix has K % 256 == 0).  split_mma: B2TTS_AR_MMA=1 sends the F32 matrices through gemv_mma_kernel<true>
    -- W and x as fp16 (hi, lo) pairs, x.W ~ xh.Wh + (xl.Wh + xh.Wl), three tensor-core products instead of an fp32 FMA chain -- which has to stay
    fp32-faithful: same greedy tokens as the reference and logits within 1e-4 (the plain fp32 path sits at ~6e-6)."""
    g = np.load(os.path.join(GOLD, "orpheus_wide_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].size)
    tok, logits = _run_ar(tmp_path, "orpheus", cached_orpheus_gguf(seed=0, head_dim=128), prompts, steps, "w", env={"B2TTS_AR_MMA": "1"} if mma else None)
    for u in range(2):
        d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
        print(f"PARITY(emulated) orpheus wide {'split mma' if mma else 'plain'} prompt {u}: max |logit diff| {d:.3e}")
        assert np.array_equal(tok[u, :, 0], g[f"tokens{u}"])
        assert d < 1e-4


def test_tensor_core_gemv_multi_tile_emulated(tmp_path):
    """A batch of 18 sequences through the split-operand tensor-core GEMV: the decode steps use two m16 tiles per block (18 rows), the 171-row prompt pass four
    (chunks of 64) -- the weight fragment of a k-step is reused for every tile, so a large batch still streams each matrix once.  Every sequence must produce the
    reference's tokens for its prompt, whatever its position in the batch."""
    g = np.load(os.path.join(GOLD, "orpheus_wide_vectors.npz"))
    prompts = [g[f"prompt{u % 2}"] for u in range(18)]
    steps = 2
    tok, logits = _run_ar(tmp_path, "orpheus", cached_orpheus_gguf(seed=0, head_dim=128), prompts, steps, "mt", env={"B2TTS_AR_MMA": "1"})
    for u in range(18):
        assert np.array_equal(tok[u, :, 0], g[f"tokens{u % 2}"][:steps]), u
        assert float(np.abs(logits[u] - g[f"logits{u % 2}"][:steps]).max()) < 1e-4


def test_parler_tensor_core_prompt_pass_row_chunks_emulated(tmp_path):
    """Parler F16 through the tensor-core GEMV with a prompt pass of more than 64 rows: the grouped q / k / v launch then runs in row chunks of 64 and the k / v rows of
    the later chunks must still land in THEIR cache slots (GemvOut::row_dst advanced per chunk).  Every sequence of the batch must give the reference's tokens."""
    g = np.load(os.path.join(GOLD, "parler_f16_vectors.npz"))
    n = 2 * (64 // int(g["prompt0"].size + g["prompt1"].size) + 1)
    prompts = [g[f"prompt{u % 2}"] for u in range(n)]
    assert sum(p.size for p in prompts) > 64
    steps = 2
    tok, _ = _run_ar(tmp_path, "parler", cached_parler_gguf(seed=0, f16=True), prompts, steps, "rc", env={"B2TTS_AR_MMA": "1"})
    for u in range(n):
        assert np.array_equal(tok[u], g[f"tokens{u % 2}"][:steps]), u


@pytest.mark.parametrize("quant", ["Q8_0", "Q5_0", "Q4_0"])
def test_parler_quantised_emulated_teacher_forced(tmp_path, quant):
    """parler.cu on block-quantised GGUFs (gemv_rows_q_kernel: activations quantised to Q8_0 per 32 columns in shared memory, dp4a over the ggml blocks as they are
    in the file) under emulation, teacher-forced on the reference's tokens (see test_parler_port_quantised_teacher_forced for why): logits within 0.1 RMS of the
    reference at every step and the same token wherever the reference's top-2 gap exceeds 0.5."""
    g = np.load(os.path.join(GOLD, f"parler_{quant.lower()}_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    tf = str(tmp_path / "teacher.bin")
    np.stack([g["tokens0"], g["tokens1"]]).astype(np.int32).tofile(tf)
    tok, logits = _run_ar(tmp_path, "parler", cached_parler_gguf(seed=0, quant=quant), prompts, steps, "q", env={"B2EMU_TEACHER": tf})
    for u in range(2):
        ref_t, ref_l = g[f"tokens{u}"], g[f"logits{u}"].reshape(steps, -1)
        rms = np.sqrt(((logits[u] - ref_l) ** 2).mean(axis=1))
        top2 = np.sort(g[f"logits{u}"], axis=2)[:, :, -2:]
        clear = (top2[:, :, 1] - top2[:, :, 0]) > 0.5
        print(f"PARITY(emulated) parler {quant} prompt {u}: per-step logit rms {np.round(rms, 4).tolist()}, tokens equal {int((tok[u] == ref_t).sum())}/{ref_t.size}")
        assert float(rms.max()) < 0.1
        assert np.array_equal(tok[u][clear], ref_t[clear])


def test_dia_quantised_emulated_teacher_forced(tmp_path):
    """dia.cu on the Q8_0 GGUF of the quantize tool (every matrix but the output heads as Q8_0 blocks), teacher-forced on the reference's tokens.  Dia amplifies the
    re-quantisation noise far more than Parler (softmax without 1/sqrt(d), 4x CFG gain, logit std ~13): the restated port itself sits up to 2.6 RMS from the
    reference on one prompt and is exact on the other.  Bar (smoke level): per-step logit RMS below 4 and at least 80 % of the teacher-forced tokens equal."""
    g = np.load(os.path.join(GOLD, "dia_q8_0_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    tf = str(tmp_path / "teacher.bin")
    np.stack([g["tokens0"], g["tokens1"]]).astype(np.int32).tofile(tf)
    tok, logits = _run_ar(tmp_path, "dia", cached_dia_gguf(seed=0, quant="Q8_0"), prompts, steps, "q", env={"B2EMU_TEACHER": tf})
    for u in range(2):
        ref_t, ref_l = g[f"tokens{u}"], g[f"logits{u}"].reshape(steps, -1)
        rms = np.sqrt(((logits[u] - ref_l) ** 2).mean(axis=1))
        agree = float((tok[u] == ref_t).mean())
        print(f"PARITY(emulated) dia Q8_0 prompt {u}: per-step logit rms {np.round(rms, 3).tolist()}, tokens equal {agree:.2f}")
        assert float(rms.max()) < 4.0 and agree >= 0.8


def test_parler_replacement_text_encoding_emulated(tmp_path):
    """Parler::set_text_encoding (the cross K / V of every layer recomputed on the device from a 7-row encoding instead of the stored 12 rows -- what
    update_conditional_prompt does after its T5 pass) under emulation against the reference run with the same replacement."""
    g = np.load(os.path.join(GOLD, "parler_encoding_vectors.npz"))
    ef = str(tmp_path / "enc.f32")
    g["encoding"].astype(np.float32).tofile(ef)
    steps = int(g["tokens0"].shape[0])
    tok, logits = _run_ar(tmp_path, "parler", cached_parler_gguf(seed=0), [g["prompt0"]], steps, "e", env={"B2EMU_ENCODING": f"{ef} {g['encoding'].shape[0]}"})
    assert np.array_equal(tok[0], g["tokens0"])
    assert float(np.abs(logits[0] - g["logits0"].reshape(steps, -1)).max()) < 1e-2


ASAN_CASES = {
    "orpheus_plain": ("orpheus", lambda: cached_orpheus_gguf(seed=0), "orpheus_vectors", {}),
    "orpheus_split_mma_graph": ("orpheus", lambda: cached_orpheus_gguf(seed=0, head_dim=128), "orpheus_wide_vectors", {"B2TTS_AR_MMA": "1", "B2TTS_AR_GRAPH": "1", "B2EMU_NO_LOGITS": "1"}),
    "parler_f16_mma_sampling_stop": ("parler", lambda: cached_parler_gguf(seed=0, f16=True), "parler_f16_vectors", {"B2TTS_AR_MMA": "1", "B2EMU_SAMPLE": "20 0.9 0.8 1.2 5", "B2EMU_STOP": "1"}),
    "parler_q5_0": ("parler", lambda: cached_parler_gguf(seed=0, quant="Q5_0"), "parler_q5_0_vectors", {}),
    "dia_q8_0_plain_attention": ("dia", lambda: cached_dia_gguf(seed=0, quant="Q8_0"), "dia_q8_0_vectors", {"B2TTS_AR_ATT": "plain"}),
    # the persistent decode kernel: ring stages, activation / attention scratch, page pool, page table, replicated hand-off buffers -- one exact-size allocation each
    "parler_f16_persistent_kernel": ("parler", lambda: cached_parler_gguf(seed=0, f16=True), "parler_f16_vectors", {"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "3", "B2TTS_AR_EXIT_EVERY": "2"}),
    "orpheus_f16_persistent_kernel": ("orpheus", lambda: cached_orpheus_gguf(seed=0, head_dim=128, f16=True), "orpheus_wide_vectors", {"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "3", "B2TTS_AR_EXIT_EVERY": "1"}),
}


@pytest.mark.parametrize("case", list(ASAN_CASES))
def test_address_sanitizer_emulated(tmp_path, case):
    """The emulated decode paths under AddressSanitizer + the alignment sanitizer: "device" memory is the host heap and the workspace arena hands out one exact-size
    allocation per buffer, so a kernel that reads or writes past the end of a weight matrix, a KV cache or a workspace buffer aborts here (on a GPU it would fault, or
    silently corrupt a neighbour), and so does a float4 / uint4 / half2 access that is not naturally aligned (a "misaligned address" fault on a GPU, silent on x86)."""
    model, gguf, gold, env = ASAN_CASES[case]
    exe = emu_build.build("ar_emu_asan", AR_SOURCES, ["ar_main.cpp", os.path.join(emu_build.CSRC, "gguf_reader.cpp")], asan=True)
    g = np.load(os.path.join(GOLD, gold + ".npz"))
    pin, pout = str(tmp_path / "p.bin"), str(tmp_path / "o.bin")
    steps = 3
    prompts = [g["prompt0"]] if "split_mma" in case else [g["prompt0"], g["prompt1"]]         # the wide model is the slow one under the sanitizer
    with open(pin, "wb") as f:
        f.write(struct.pack("ii", len(prompts), steps))
        for p in prompts:
            f.write(struct.pack("i", p.size)); f.write(p.astype(np.uint32).tobytes())
    r = subprocess.run([exe, model, gguf(), pin, pout], capture_output=True, text=True, timeout=900,
                       env={**os.environ, **EMU_DEFAULTS, "ASAN_OPTIONS": "detect_stack_use_after_return=0:detect_leaks=0", **env})
    assert r.returncode == 0, r.stderr[-3000:]
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr


# ---- the persistent decode kernel (tts_cpp_b200/csrc/pdk.cuh): one cooperative launch runs up to 32 decode steps; under emulation every block of the grid is alive at
# once (b2emu::launch_coop), mbarriers / bulk copies / the grid barrier have functional models.  F16 Parler GGUFs take this path by default (B2TTS_AR_PDK=0: per-op path).
@pytest.mark.parametrize("env", [{"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "1"}, {"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "3", "B2TTS_AR_EXIT_EVERY": "2"}, {"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "5", "B2EMU_REVERSE": "1"},
                                 {"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "7", "B2TTS_AR_EXIT_EVERY": "1", "B2TTS_KV": "f32"}], ids=["grid1", "grid3_chunks_of_2", "grid5_reversed", "grid7_kv_f32_chunks_of_1"])
def test_persistent_decode_kernel_emulated_matches_reference(tmp_path, env):
    """the reference's F16 tokens and logits (tests/golden/parler_f16_vectors.npz) from the persistent kernel with its paged fp16 (or fp32) KV cache, for several grid sizes
    (different unit -> CTA distributions; grid 1 = no concurrency), several launches per generation (step counter / ring / page state carried across launches) and both
    fiber schedules; the run must really have gone through the persistent kernel (launch count)."""
    g = np.load(os.path.join(GOLD, "parler_f16_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    tok, logits, err = _run_ar(tmp_path, "parler", cached_parler_gguf(seed=0, f16=True), prompts, steps, "pk", env=env, want_stderr=True)
    _, _, err0 = _run_ar(tmp_path, "parler", cached_parler_gguf(seed=0, f16=True), prompts, steps, "op", env={"B2TTS_AR_PDK": "0"}, want_stderr=True)
    n_pk, n_op = (int(e.split("emulated ")[1].split(" launches")[0]) for e in (err, err0))
    assert n_pk < n_op - (steps - 1) * 50, (n_pk, n_op)          # the per-op path launches ~126 kernels per step of the 8-layer test model
    for u in range(2):
        ref = g[f"logits{u}"].reshape(steps, -1)
        d = float(np.abs(logits[u] - ref).max())
        print(f"PARITY(emulated, persistent kernel {env}) parler f16 prompt {u}: max |logit diff| {d:.3e}")
        assert np.array_equal(tok[u], g[f"tokens{u}"])
        assert d < 3e-2


def test_persistent_decode_kernel_emulated_stop_rule_and_teacher(tmp_path):
    """the delay pattern's stop bookkeeping (eos_seen, check_stopping) and teacher forcing inside the persistent kernel against the per-op path on an EOS-boosted F16 GGUF:
    same tokens, same frame count, zero rows past the stop; teacher-forced: same tokens, logits within the F16 floor."""
    g = np.load(os.path.join(GOLD, "parler_stop_vectors.npz"))
    prompt, boost = g["all_eos.prompt"], float(g["all_eos.boost"])
    cap = int(g["all_eos.tokens"].shape[0]) + 6
    exe = emu_build.build("ar_emu", AR_SOURCES, ["ar_main.cpp", os.path.join(emu_build.CSRC, "gguf_reader.cpp")])
    gguf = cached_parler_gguf(seed=0, eos_boost=boost, f16=True)
    # the teacher-forced run of the second half is independent of the two stop-rule runs: all three go to separate cores
    g16 = np.load(os.path.join(GOLD, "parler_f16_vectors.npz"))
    tf = str(tmp_path / "teacher.bin")
    np.stack([g16["tokens0"], g16["tokens1"]]).astype(np.int32).tofile(tf)
    f_tf = _bg(_run_ar, tmp_path, "parler", cached_parler_gguf(seed=0, f16=True), [g16["prompt0"], g16["prompt1"]], int(g16["tokens0"].shape[0]), "tf",
               env={"B2EMU_TEACHER": tf, "B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "4"})
    outs, runs = {}, {}
    for tag, env in (("pk", {"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "3"}), ("op", {"B2TTS_AR_PDK": "0"})):
        pin, pout = str(tmp_path / f"p{tag}.bin"), str(tmp_path / f"o{tag}.bin")
        with open(pin, "wb") as f:
            f.write(struct.pack("ii", 1, cap)); f.write(struct.pack("i", prompt.size)); f.write(prompt.astype(np.uint32).tobytes())
        runs[tag] = (pout, _bg(subprocess.run, [exe, "parler", gguf, pin, pout], capture_output=True, text=True, timeout=900,
                               env={**os.environ, **EMU_DEFAULTS, "B2EMU_STOP": "1", "B2EMU_NO_LOGITS": "1", "B2TTS_AR_EXIT_EVERY": "4", **env}))
    for tag, (pout, fut) in runs.items():
        r = fut.result()
        assert r.returncode == 0, r.stderr[-2000:]
        raw = open(pout, "rb").read()
        W, V = struct.unpack("ii", raw[:8])
        outs[tag] = (np.frombuffer(raw, np.int32, cap * W, 8).reshape(cap, W), int(np.frombuffer(raw[-4:], np.int32)[0]))
    assert outs["pk"][1] == outs["op"][1] and 0 < outs["pk"][1] < cap, (outs["pk"][1], outs["op"][1])
    assert np.array_equal(outs["pk"][0], outs["op"][0]) and not outs["pk"][0][outs["pk"][1]:].any()
    # teacher-forced on the reference's F16 tokens
    g = np.load(os.path.join(GOLD, "parler_f16_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].shape[0])
    tok, logits = f_tf.result()
    for u in range(2):
        assert np.array_equal(tok[u], g[f"tokens{u}"])
        assert float(np.abs(logits[u] - g[f"logits{u}"].reshape(steps, -1)).max()) < 3e-2


@pytest.mark.parametrize("kind", ["q8_0", "f16"])
def test_orpheus_quantised_and_f16_matrices_emulated(tmp_path, kind):
    """Orpheus with Q8_0 (BASELINE config 5's dtype; our own writer -- the reference's quantize tool refuses Orpheus and its runtime is F32-only) or F16 matrices through
    the block-quantised / F16 GEMV kernels: no reference output exists for these files, so the yardstick is the reference's F32 run of the same weights
    (tests/golden/orpheus_vectors.npz): logits within the storage format's noise and the same greedy token wherever the F32 top-2 gap is clear."""
    g = np.load(os.path.join(GOLD, "orpheus_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].size)
    gguf = cached_orpheus_gguf(seed=0, quant="Q8_0") if kind == "q8_0" else cached_orpheus_gguf(seed=0, f16=True)
    tok, logits = _run_ar(tmp_path, "orpheus", gguf, prompts, steps, kind)
    for u in range(2):
        ref = g[f"logits{u}"].reshape(steps, -1)
        rel = float(np.sqrt(((logits[u] - ref) ** 2).mean()) / ref.std())
        top2 = np.sort(ref, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 8.0 * np.abs(logits[u] - ref).max(axis=1)
        print(f"PARITY(emulated) orpheus {kind} prompt {u}: logit rms / std {rel:.3e}; tokens equal {int((tok[u, :, 0] == g[f'tokens{u}']).sum())}/{steps}, clear decisions {int(clear.sum())}")
        assert rel < (0.05 if kind == "q8_0" else 0.01)
        assert np.array_equal(tok[u, :, 0][clear], g[f"tokens{u}"][clear])


@pytest.mark.parametrize("env", [{"B2TTS_PDK_GRID": "3"}, {"B2TTS_PDK_GRID": "32", "B2TTS_KV": "f32", "B2TTS_PDK_AK": "768", "B2EMU_REVERSE": "1"}],
                         ids=["grid3_f16kv", "grid32_f32kv_two_k_chunks_reverse"])
def test_persistent_decode_kernel_emulated_orpheus(tmp_path, env):
    """Orpheus (llama-3 style) through the persistent kernel under emulation: RMSNorm folded into the staging, paired units (NeoX RoPE halves of q / k, gate + up for SwiGLU)
    with their joint epilogues, GQA attention over the pages, argmax partials combined by the next step's rows phase, three launches per generation
    (B2TTS_AR_EXIT_EVERY=16), the down projection in one or two k-chunks.  Yardstick: the reference's F32 run of the same (fp16-representable) weights
    (tests/golden/orpheus_wide_long_vectors.npz): same tokens up to the first near-tie, logits within the F16 floor."""
    g = np.load(os.path.join(GOLD, "orpheus_wide_long_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = 34 if "B2TTS_PDK_AK" not in env else 20             # (34: past the first KV-page boundary of both prompts and two launch boundaries)
    gguf = cached_orpheus_gguf(seed=0, head_dim=128, f16=True)
    tok, logits, err = _run_ar(tmp_path, "orpheus", gguf, prompts, steps, "pk", env={"B2TTS_AR_PDK": "1", "B2TTS_AR_EXIT_EVERY": "16", **env}, want_stderr=True)
    _, _, err0 = _run_ar(tmp_path, "orpheus", gguf, prompts, 4, "op", env={"B2TTS_AR_PDK": "0"}, want_stderr=True)
    n_pk, n_op = (int(e.split("emulated ")[1].split(" launches")[0]) for e in (err, err0))
    assert n_pk < n_op + 8, (n_pk, n_op)                         # prompt pass + 3 cooperative launches vs prompt pass + 3 steps of ~25 launches
    for u in range(2):
        ref_t, ref_l = g[f"tokens{u}"][:steps], g[f"logits{u}"][:steps]
        neq = np.nonzero(tok[u, :, 0] != ref_t)[0]
        upto = int(neq[0]) if neq.size else steps - 1
        d = np.abs(logits[u][:upto + 1] - ref_l[:upto + 1]).max(axis=1)
        print(f"PARITY(emulated, persistent kernel {env}) orpheus wide f16 prompt {u}: tokens equal for {upto + (0 if neq.size else 1)}/{steps} steps, max |logit diff| {float(d.max()):.3e}")
        assert float(d.max()) < 3e-2
        if neq.size:                                            # a differing token must be a near-tie of the reference
            top2 = np.sort(ref_l[upto])[-2:]
            assert float(top2[1] - top2[0]) <= 2.0 * float(d[upto]), (upto, float(top2[1] - top2[0]), float(d[upto]))
        assert upto >= 18


def test_persistent_decode_kernel_emulated_multi_tile_units(tmp_path):
    """k extents of several weight tiles (ffn 2 560: the down projection's units are 1 024 + 1 024 + 512 columns -- the two-tiles-at-a-time loop and its odd tail; K = 768
    elsewhere): no reference output exists for this shape, so the persistent kernel is held against the launch-per-op path on the same F16 GGUF: same tokens (or a near-tie
    of the per-op logits), logits within twice the F16 floor the two paths show against the reference elsewhere."""
    g = np.load(os.path.join(GOLD, "orpheus_wide_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = 8
    gguf = cached_orpheus_gguf(seed=0, head_dim=128, ffn=2560, f16=True)
    tok, logits = _run_ar(tmp_path, "orpheus", gguf, prompts, steps, "pk", env={"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "5", "B2TTS_AR_EXIT_EVERY": "3"})
    tok0, logits0 = _run_ar(tmp_path, "orpheus", gguf, prompts, steps, "op", env={"B2TTS_AR_PDK": "0"})
    for u in range(2):
        neq = np.nonzero(tok[u, :, 0] != tok0[u, :, 0])[0]
        upto = int(neq[0]) if neq.size else steps - 1
        d = np.abs(logits[u][:upto + 1] - logits0[u][:upto + 1]).max(axis=1)
        print(f"PARITY(emulated, persistent kernel vs per-op, ffn 2560) prompt {u}: tokens equal for {upto + (0 if neq.size else 1)}/{steps} steps, max |logit diff| {float(d.max()):.3e}")
        assert float(d.max()) < 3e-2
        if neq.size:
            top2 = np.sort(logits0[u][upto])[-2:]
            assert float(top2[1] - top2[0]) <= 2.0 * float(d[upto])
        assert upto >= 4


@pytest.mark.parametrize("env", [{"B2TTS_PDK_GRID": "3", "B2TTS_PDK_TSPLIT": "1"}, {"B2TTS_PDK_GRID": "6", "B2EMU_REVERSE": "1", "B2TTS_PDK_TSPLIT": "3"}], ids=["grid3", "grid6_reverse_split_attention"])
def test_persistent_decode_kernel_emulated_dia(tmp_path, env):
    """Dia (encoder per op, then the whole CFG decoder loop inside the persistent kernel: delay pattern + end-of-stream injection in the rows phase, RoPE'd self and cross
    queries, GQA self-attention over the pages, cross-attention over each row's own encoding, SwiGLU, cfg_scale + argmax) against the reference's F16 run
    (tests/golden/dia_wide_f16_vectors.npz): teacher-forced logits around the first page boundary, then the free-running loop with check_stopping's frame count."""
    g = np.load(os.path.join(GOLD, "dia_wide_f16_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps, steps_op = 34, 10
    gguf = cached_dia_gguf(seed=0, f16=True, head_dim=64)
    tf = str(tmp_path / "teacher.bin")
    np.stack([g["tokens0"][:steps], g["tokens1"][:steps]]).astype(np.int32).tofile(tf)
    exe = emu_build.build("ar_emu", AR_SOURCES, ["ar_main.cpp", os.path.join(emu_build.CSRC, "gguf_reader.cpp")])      # built before the concurrent runs below
    f_pk = _bg(_run_ar, tmp_path, "dia", gguf, prompts, steps, "pk", env={"B2TTS_AR_PDK": "1", "B2TTS_AR_EXIT_EVERY": "16", "B2EMU_TEACHER": tf, **env}, want_stderr=True)
    tf2 = str(tmp_path / "teacher_op.bin")
    np.stack([g["tokens0"][:steps_op], g["tokens1"][:steps_op]]).astype(np.int32).tofile(tf2)
    f_op = _bg(_run_ar, tmp_path, "dia", gguf, prompts, steps_op, "op", env={"B2TTS_AR_PDK": "0", "B2EMU_TEACHER": tf2}, want_stderr=True)
    f_free = None
    cap = int(g["step_cap"])
    pin, pout = str(tmp_path / "ps.bin"), str(tmp_path / "os.bin")
    if "B2EMU_REVERSE" not in env:      # free-running with the stop rule (checked at the end), started now
        with open(pin, "wb") as f:
            f.write(struct.pack("ii", 1, cap)); f.write(struct.pack("i", prompts[0].size)); f.write(prompts[0].astype(np.uint32).tobytes())
        f_free = _bg(subprocess.run, [exe, "dia", gguf, pin, pout], capture_output=True, text=True, timeout=900,
                     env={**os.environ, **EMU_DEFAULTS, "B2TTS_AR_PDK": "1", "B2EMU_STOP": "1", "B2EMU_NO_LOGITS": "1", "B2TTS_AR_EXIT_EVERY": "16", **env})
    tok, logits, err = f_pk.result()
    tok0, logits0, err0 = f_op.result()
    n_pk, n_op = (int(e.split("emulated ")[1].split(" launches")[0]) for e in (err, err0))
    assert n_pk < n_op - (steps_op - 1) * 30, (n_pk, n_op)       # encoder pass + 3 cooperative launches (34 steps) vs ~45 launches per decoder step of the 2-layer test model (10 steps)
    keep = [int(s) for s in g["logit_steps"] if s < steps]
    for u in range(2):
        # Dia's F16 noise floor against the reference is large (logit std 13, differences up to ~0.7 on either path: tests/test_dia_gpu.py); the persistent kernel must sit
        # on the per-op path (same arithmetic, other summation order) and both within that floor of the reference, with the reference's token wherever its top-2 gap is clear
        ref = g[f"logits{u}"][:len(keep)].reshape(len(keep), -1)
        d = np.abs(logits[u][keep] - ref).max(axis=1)
        dp = float(np.abs(logits[u][:steps_op] - logits0[u]).max())
        clear = g[f"gap{u}"][:steps] > 4.0 * float(d.max())
        print(f"PARITY(emulated, persistent kernel {env}) dia wide f16 prompt {u}: max |logit diff| vs the reference over frames {keep[0]}-{keep[-1]} {float(d.max()):.3e}, vs the per-op path {dp:.3e}; "
              f"tokens equal {int((tok[u] == g[f'tokens{u}'][:steps]).sum())}/{steps * 9}, clear decisions {int(clear.sum())}")
        assert float(d.max()) < 1.5 and dp < 0.2                 # (fp32 pages, Dia's default: with fp16 pages the cache's rounding alone moves these logits by up to 2.5)
        assert np.array_equal(tok[u][clear], g[f"tokens{u}"][:steps][clear])
    if f_free is None: return
    # free-running with the stop rule: the loop ends after the reference's number of frames
    r = f_free.result()
    assert r.returncode == 0, r.stderr[-2000:]
    raw = open(pout, "rb").read()
    W, V = struct.unpack("ii", raw[:8])
    toks = np.frombuffer(raw, np.int32, cap * W, 8).reshape(cap, W)
    n_gen = int(np.frombuffer(raw[-4:], np.int32)[0])
    ref = g["tokens0"]
    assert n_gen == ref.shape[0] and not toks[n_gen:].any()
    neq = np.argwhere(toks[:n_gen] != ref)
    first = int(neq[0][0]) if neq.size else n_gen
    print(f"free-running: {n_gen} frames, tokens equal up to frame {first}")
    dmax0 = float(np.abs(logits[0][keep] - g["logits0"][:len(keep)].reshape(len(keep), -1)).max())
    assert all(g["gap0"][s, h] < 2.0 * dmax0 for s, h in neq if s == first)      # the first difference, if any, is within the F16 floor measured above on the reference's own tokens
    assert first >= 20


def test_gguf_reader_rejects_hostile_files(tmp_path):
    """csrc/gguf_reader.cpp on corrupt / hostile input: counts and string lengths beyond the file, a truncated file, a tensor directory pointing past the
    end, an impossible shape -- every one must come back as a load error from the reader (exit code 1 of the emulation driver), never a crash or an exception."""
    exe = emu_build.build("ar_emu", AR_SOURCES, ["ar_main.cpp", os.path.join(emu_build.CSRC, "gguf_reader.cpp")])
    good = open(cached_orpheus_gguf(seed=0), "rb").read()
    hdr = struct.pack("<II", 0x46554747, 3)
    cases = {
        "huge_counts": hdr + struct.pack("<QQ", 1 << 60, 1 << 60) + b"\0" * 64,
        "huge_tensor_count": hdr + struct.pack("<QQ", 1 << 40, 0) + b"\0" * 64,
        "huge_string": hdr + struct.pack("<QQ", 0, 1) + struct.pack("<Q", (1 << 64) - 8) + b"x" * 64,
        "huge_array": hdr + struct.pack("<QQ", 0, 1) + struct.pack("<Q", 1) + b"k" + struct.pack("<IIQ", 9, 4, 1 << 62) + b"\0" * 64,
        "truncated_directory": good[:len(good) // 200],
        "truncated_data": good[:len(good) // 10],
        "too_short": hdr,
    }
    # a tensor whose shape overflows: patch the first dimension of the first "orpheus." tensor of the good file
    at = good.find(b"orpheus.")
    name_len = struct.unpack("<Q", good[at - 8:at])[0]
    dims_at = at + name_len + 4
    cases["impossible_shape"] = good[:dims_at] + struct.pack("<q", -7) + good[dims_at + 8:]
    pin, pout = str(tmp_path / "p.bin"), str(tmp_path / "o.bin")
    with open(pin, "wb") as f:
        f.write(struct.pack("iii", 1, 2, 1)); f.write(struct.pack("I", 5))
    for tag, blob in cases.items():
        path = str(tmp_path / f"{tag}.gguf")
        open(path, "wb").write(blob)
        r = subprocess.run([exe, "orpheus", path, pin, pout], capture_output=True, text=True, timeout=120, env={**os.environ, **EMU_DEFAULTS})
        assert r.returncode == 1 and "load:" in r.stderr, (tag, r.returncode, r.stderr[-300:])


def test_persistent_decode_kernel_emulated_orpheus_q8_0(tmp_path):
    """Orpheus with Q8_0 matrices (BASELINE config 5's dtype) through the persistent kernel: activations quantised per 32-block while they are staged (RMSNorm'd fp32 rows
    and the fp16 hand-offs), int8 MMA per block, fp32 scale products -- ggml_vec_dot_q8_0_q8_0's arithmetic.  No reference output exists for Q8_0 Orpheus (its runtime is
    F32-only): the yardsticks are the reference's F32 run of the same weights (format noise, as for the launch-per-op Q8_0 path) and the launch-per-op Q8_0 path itself."""
    g = np.load(os.path.join(GOLD, "orpheus_wide_vectors.npz"))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = int(g["tokens0"].size)
    gguf = cached_orpheus_gguf(seed=0, head_dim=128, quant="Q8_0")
    tok, logits, err = _run_ar(tmp_path, "orpheus", gguf, prompts, steps, "pk", env={"B2TTS_AR_PDK": "1", "B2TTS_PDK_GRID": "4", "B2TTS_AR_EXIT_EVERY": "2"}, want_stderr=True)
    tok0, logits0, err0 = _run_ar(tmp_path, "orpheus", gguf, prompts, steps, "op", env={"B2TTS_AR_PDK": "0"}, want_stderr=True)
    n_pk, n_op = (int(e.split("emulated ")[1].split(" launches")[0]) for e in (err, err0))
    assert n_pk < n_op - (steps - 2) * 15, (n_pk, n_op)
    for u in range(2):
        ref = g[f"logits{u}"].reshape(steps, -1)
        rel = float(np.sqrt(((logits[u] - ref) ** 2).mean()) / ref.std()), float(np.sqrt(((logits0[u] - ref) ** 2).mean()) / ref.std())
        top2 = np.sort(ref, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 8.0 * np.abs(logits[u] - ref).max(axis=1)
        print(f"PARITY(emulated, persistent kernel) orpheus wide Q8_0 prompt {u}: logit rms / std vs the F32 reference {rel[0]:.3e} (launch-per-op Q8_0: {rel[1]:.3e}); "
              f"tokens equal {int((tok[u, :, 0] == g[f'tokens{u}']).sum())}/{steps}, clear decisions {int(clear.sum())}")
        assert rel[0] < 0.05 and rel[0] < 1.5 * rel[1] + 1e-3
        assert np.array_equal(tok[u, :, 0][clear], g[f"tokens{u}"][clear])


def test_vad_kernels_emulated_match_reference(tmp_path):
    """vad.cu (vad_energy_kernel + vad_decide_kernel through vad_trim_rows) under emulation against the compiled unmodified examples/cli/vad.cpp
    (tests/golden/vad_vectors.npz): trimmed lengths and frame energies bit for bit on every case of vad_cases.py, in both thread orders and under AddressSanitizer
    (exact-size PCM / energy buffers: a read past an utterance's last whole frame or a write past the energy array is reported)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, ROOT)
    import vad_cases
    g = np.load(os.path.join(ROOT, "tests", "golden", "vad_vectors.npz"))
    exe = emu_build.build("vad_emu", ["vad.cu"], ["vad_main.cpp"])
    exe_asan = emu_build.build("vad_emu_asan", ["vad.cu"], ["vad_main.cpp"], asan=True)
    for name, kw, utts in vad_cases.cases():
        pin, pout = str(tmp_path / f"{name}.in"), str(tmp_path / f"{name}.out")
        open(pin, "wb").write(vad_cases.pack_input(kw, utts))
        raws = []
        for e, env in ((exe, {}), (exe, {"B2EMU_REVERSE": "1"}), (exe_asan, {"ASAN_OPTIONS": "detect_stack_use_after_return=0:detect_leaks=0"})):
            r = subprocess.run([e, pin, pout], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
            assert r.returncode == 0, (name, r.stderr[-3000:])
            raws.append(open(pout, "rb").read())
        assert raws[0] == raws[1] == raws[2]
        n_out, en = vad_cases.unpack_output(raws[0], kw, utts)
        assert np.array_equal(n_out, g[name + ".n_out"]), (name, n_out, g[name + ".n_out"])
        for b in range(len(utts)):
            assert np.array_equal(en[b], g[f"{name}.energies.{b}"]), (name, b)


@pytest.mark.parametrize("case", ["f32", "f16", "no_down_proj", "wide", "f16_wide", "q8_0", "q5_0", "q4_0"])
def test_t5_encoder_cuda_path_emulated_matches_reference(tmp_path, case):
    """T5::encode (t5.cu: embedding rows, RMS norm eps 1e-6, the storage-aware GEMVs, bidirectional attention with the relative-position bias table, gated GELU, down
    projection + bias) under emulation against the compiled unmodified t5_runner::run (tests/golden/t5_vectors.npz), the prompts of a case as ONE ragged batch:
    at the port's own distance from the reference (fp16 GELU table; fp16 activation rounding for F16 matrices).  The "wide" 88-token prompt hits every
    relative-position bucket, incl. the log-spaced ones with the reference's integer division."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, ROOT)
    import make_golden_t5 as M
    from tts_cpp_b200.synth import cached_t5_gguf
    g = np.load(os.path.join(ROOT, "tests", "golden", "t5_vectors.npz"))
    kw, prompts = M.CASES[case]
    exe = emu_build.build("ar_emu", AR_SOURCES, ["ar_main.cpp", os.path.join(emu_build.CSRC, "gguf_reader.cpp")])
    pin, pout = str(tmp_path / "p.bin"), str(tmp_path / "o.bin")
    with open(pin, "wb") as f:
        f.write(struct.pack("<ii", len(prompts), 0))
        for p in prompts:
            f.write(struct.pack("<i", len(p))); f.write(np.asarray(p, np.uint32).tobytes())
    r = subprocess.run([exe, "t5", cached_t5_gguf(**kw), pin, pout], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    raw = open(pout, "rb").read()
    O = struct.unpack("<i", raw[:4])[0]
    enc = np.frombuffer(raw[4:], np.float32).reshape(-1, O)
    assert enc.shape[0] == sum(len(p) for p in prompts)
    at = 0
    for i, p in enumerate(prompts):
        ref = g[f"{case}.encoding.{i}"]
        got = enc[at:at + len(p)]; at += len(p)
        d = float(np.sqrt(((got - ref) ** 2).mean())); rr = float(np.sqrt((ref ** 2).mean()))
        print(f"PARITY t5 emulated {case}.{i}: rms {d:.3g} of {rr:.3g}, max {float(np.abs(got - ref).max()):.3g}")
        assert d < _T5_TOL(case) * rr, (case, i, d, rr)
    if case == "f32":      # a wrong prompt is refused, not read out of bounds: token id >= vocabulary, an empty prompt
        for bad in ([5, 96, 1], []):
            with open(pin, "wb") as f:
                f.write(struct.pack("<ii", 1, 0)); f.write(struct.pack("<i", len(bad))); f.write(np.asarray(bad, np.uint32).tobytes())
            r = subprocess.run([exe, "t5", cached_t5_gguf(**kw), pin, pout], capture_output=True, text=True, timeout=900)
            assert r.returncode == 1 and "t5: prompt 0" in r.stderr, r.stderr[-500:]
