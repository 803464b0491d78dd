"""Generate tests/golden/*.npz from the compiled UNMODIFIED reference (oracle/_ref, built by oracle/Makefile).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The vectors pin oracle/kokoro_port.py (tests/test_oracle_port.py, CPU) and, through it, the CUDA path.

  kokoro_stage_vectors.npz : one 8-token utterance on the deterministic synthetic GGUF (seed 0, f16 policy, ctx 128):
      tokens, lens, hidden (duration_hidden_states), f0, n, dec (generator input), har_spec (STFT of the harmonic source),
      pcm -- all dumped from the reference's own GGML graph nodes by oracle/ref_kokoro_driver.cpp
  op_vectors.npz           : known-answer vectors of the patched ggml ops from oracle/ref_ops_driver.cpp
  dac_vectors.npz          : two 24-frame utterances of codebook indices and the PCM the reference's dac_runner produces for them
      on the deterministic synthetic DAC GGUF (seed 0, all F32), from oracle/ref_dac_driver.cpp
  snac_vectors.npz         : two utterances (16 fine frames: 4 + 8 + 16 indices) decoded IN ONE PROCESS by the reference's snac_runner
      (its std::normal_distribution noise stream carries over from the first to the second), from oracle/ref_snac_driver.cpp
  orpheus_vectors.npz      : two prompts (7 and 12 token ids) and, for 6 greedy decode steps each, the tokens the reference's decode loop +
      sampler produced and the logits of every step, on the small synthetic Orpheus GGUF (2 layers, 6/2 heads x 64, vocab 2048, F32),
      from oracle/ref_orpheus_driver.cpp
  parler_vectors.npz       : two prompts (5 and 9 ids) and, for 5 greedy audio steps each, the 9 codebook tokens per step and their logits
      from the reference's Parler decode loop (delay pattern included) on the small synthetic Parler GGUF, from oracle/ref_parler_driver.cpp
  parler_f16_vectors.npz   : as parler_vectors.npz for the GGUF `quantize --quantized-type F16` would write (decoder matrices and codebook tables F16: the
      reference then rounds the activations to fp16 before every such product)
  parler_q{8,5,4}_0_vectors.npz : as parler_vectors.npz for the GGUFs `quantize --quantized-type Q8_0 / Q5_0 / Q4_0` would write (decoder matrices and codebook
      tables as ggml blocks; the reference re-quantises the activations to Q8_0 per 32 columns before every such product)
  parler_encoding_vectors.npz : the Parler loop after the stored conditional-prompt encoding was replaced (update_conditional_prompt's prep_cross_key_values call)
  parler_stop_vectors.npz  : the Parler loop run to completion under the reference's stop rule (eos_seen feeding + check_stopping) on two EOS-boosted synthetic
      GGUFs: one ends at max_generation, one because every head produced EOS; from oracle/ref_parler_driver.cpp --stop
  dia_f16_vectors.npz      : as dia_vectors.npz for the F16 GGUF of the quantize tool (all matrices and embeddings but the output heads F16)
  orpheus_wide_vectors.npz : as orpheus_vectors.npz for a GGUF with head size 128 (hidden 768)
  dia_wide_f16_vectors.npz : Dia F16 with decoder width 256, two prompts, all 64 frames of the reference's loop (tokens, top-2 gaps, logits of 16 frames)
  orpheus_wide_long_vectors.npz : the same GGUF, prompts of 7 and 40 ids, 72 greedy steps (crosses KV-page and persistent-kernel launch boundaries)
  sampler_vectors.npz      : the reference sampler (src/sampler.cpp) on fixed logits under four configurations: nucleus, probabilities, max_head_probs and a
      histogram of 20 000 draws each, from oracle/ref_sampler_driver.cpp
  dia_q8_0_vectors.npz     : as dia_vectors.npz for the Q8_0 GGUF of the quantize tool
  dia_stop_vectors.npz     : one byte-token prompt run until the reference's check_stopping ends the loop (64 frames of 9 tokens, the logits of the
      last frame), from oracle/ref_dia_driver.cpp with a step cap of 80
  dia_vectors.npz          : two byte-token prompts (10 and 18 tokens, the second after the first in the same process) and, for 5 greedy steps
      each, the 9 codebook tokens and the CFG-combined logits from the reference's Dia encoder + decode loop, from oracle/ref_dia_driver.cpp
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.dirname(os.path.abspath(__file__))


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout


def kokoro_vectors():
    from tts_cpp_b200.synth import cached_gguf, synthetic_prompts
    gguf = cached_gguf("f16", 128, 0)
    toks = synthetic_prompts(1, n_phonemes=6, seed0=4242)[0]
    tmp = tempfile.mkdtemp()
    tokf = os.path.join(tmp, "tok.txt")
    open(tokf, "w").write(" ".join(map(str, toks)) + "\n")
    pre = os.path.join(tmp, "g")
    listing = run([os.path.join(REF, "kokoro_ref"), gguf, tokf, pre, "--threads", "4", "--list-nodes"])
    lens = np.fromfile(pre + ".u0.lens.f32", np.float32)
    T = int(lens.sum())
    nodes = [(int(m.group(1)), int(m.group(2)), [int(v) for v in m.group(3).split(",")])
             for m in re.finditer(r"NODE gen (\d+) op=(\d+) name=\S* ne=\[([\d,]+)\]", listing)]
    OP_DIV, OP_CONCAT, OP_LRELU = 7, 20, 60
    first_lrelu = min(i for i, op, ne in nodes if op == OP_LRELU and ne[:2] == [512, 2 * T])
    dec_idx = max(i for i, op, ne in nodes if op == OP_DIV and ne[:2] == [2 * T, 512] and i < first_lrelu)
    har_idx = min(i for i, op, ne in nodes if op == OP_CONCAT and ne[:2] == [22, 120 * T + 1])
    run([os.path.join(REF, "kokoro_ref"), gguf, tokf, pre, "--threads", "4", "--dump-gen", f"#{dec_idx},#{har_idx},f0_out,n_out"])
    rd = lambda n: np.fromfile(f"{pre}.u0.{n}.f32", np.float32)
    np.savez_compressed(os.path.join(OUT, "kokoro_stage_vectors.npz"),
                        tokens=np.array(toks, np.int32), lens=lens, hidden=rd("hidden").reshape(len(toks), 640),
                        f0=rd("gen.f0_out"), n=rd("gen.n_out"), dec=rd(f"gen.n{dec_idx}").reshape(512, 2 * T),
                        har_spec=rd(f"gen.n{har_idx}").reshape(120 * T + 1, 22), pcm=rd("pcm"),
                        meta=np.array([dec_idx, har_idx, T], np.int64))
    print("kokoro vectors: T =", T, "dec node", dec_idx, "har node", har_idx)


def op_vectors():
    tmp = tempfile.mkdtemp()
    ops = os.path.join(REF, "ops_ref")
    rng = np.random.default_rng(2024)
    out = {}

    def call(op, name, *args):
        o = os.path.join(tmp, name + ".f32")
        run([ops, op, o] + [str(a) for a in args])
        return np.fromfile(o, np.float32)

    def put(name, arr):
        p = os.path.join(tmp, name + ".in.f32")
        np.asarray(arr, np.float32).tofile(p)
        return p

    out["uniform_first_4096"] = call("uniform", "uni", 4096)
    out["wss_20_5_37"] = call("wss", "wss", 20, 5, 37)
    x = np.tanh(0.3 * np.sin(np.arange(615) * 0.031) + 0.01 * rng.standard_normal(615)).astype(np.float32)
    out["stft_in"] = x
    out["stft_out"] = call("stft", "stft", put("stft", x), 615, 20, 5, 1, 1).reshape(2, 124, 11)      # [mag|phase][frames][bins]
    mp = np.stack([np.exp(0.5 * rng.standard_normal((124, 11))), np.sin(rng.standard_normal((124, 11)))]).astype(np.float32)
    out["istft_in"] = mp
    out["istft_out"] = call("istft", "istft", put("istft", mp), 11, 124, 20, 5)
    cs = (rng.standard_normal((9, 40)) * 3).astype(np.float32)
    out["cumsum_in"] = cs
    out["cumsum_out"] = call("cumsum", "cumsum", put("cumsum", cs), 40, 9).reshape(9, 40)
    out["mod_out"] = call("mod", "mod", put("mod", cs), cs.size, 1.0).reshape(9, 40)
    out["round_out"] = call("round", "round", put("round", cs), cs.size).reshape(9, 40)
    out["upscale_linear_out"] = call("upscale_linear", "ul", put("ul", np.cumsum(np.abs(cs), axis=1) * 100), 40, 9, 300).reshape(9, 12000)
    out["upscale_linear_in"] = (np.cumsum(np.abs(cs), axis=1) * 100).astype(np.float32)
    al = rng.uniform(0.3, 2.0, 12).astype(np.float32)
    sx = (rng.standard_normal((12, 50)) * 2).astype(np.float32)
    out["snake_alpha"], out["snake_in"] = al, sx
    out["snake_out"] = call("snake", "snake", put("sa", al), 12, put("sx", sx), 50).reshape(12, 50)
    g = (rng.standard_normal(512) * 3).astype(np.float32)
    out["gelu_in"] = g
    out["gelu_out"] = call("gelu", "gelu", put("gelu", g), 512)
    # ConvTranspose1d: the generator's two shapes (scaled down) and the depthwise "pool"
    for tag, (K, cout, cin, L, s, p, op_, grp) in {"ct_up0": (20, 8, 16, 9, 10, 5, 0, 1), "ct_up1": (12, 4, 8, 7, 6, 3, 0, 1), "ct_pool": (3, 6, 6, 5, 2, 1, 1, 6)}.items():
        W = rng.standard_normal((cin, cout // grp, K)).astype(np.float32)
        xx = rng.standard_normal((cin, L)).astype(np.float32)
        lout = (L - 1) * s - 2 * p + (K - 1) + op_ + 1
        out[tag + "_w"], out[tag + "_x"] = W, xx
        out[tag + "_cfg"] = np.array([K, cout, cin, L, s, p, op_, grp], np.int32)
        out[tag + "_y"] = call("convt1d", tag, put(tag + "w", W), K, cout // grp, cin, put(tag + "x", xx), L, s, p, op_, grp, 0).reshape(cout, lout)
    # F16 conv1d (im2col in fp16): dilated k7
    W = (rng.standard_normal((16, 32, 7)) / 15).astype(np.float16).astype(np.float32)
    xx = rng.standard_normal((32, 60)).astype(np.float32)
    out["conv_w"], out["conv_x"] = W, xx
    out["conv_y"] = call("conv1d", "conv", put("cw", W), 7, 32, 16, put("cx", xx), 60, 1, 9, 3, 1).reshape(16, 60)
    np.savez_compressed(os.path.join(OUT, "op_vectors.npz"), **out)
    print("op vectors:", sorted(out))


def dac_vectors():
    from tts_cpp_b200.synth import cached_dac_gguf, synthetic_codes
    gguf = cached_dac_gguf(seed=0, max_frames=64)
    codes = synthetic_codes(2, 24)
    tmp = tempfile.mkdtemp()
    cf = os.path.join(tmp, "codes.txt")
    open(cf, "w").write("\n".join(" ".join(map(str, c.reshape(-1))) for c in codes) + "\n")
    pre = os.path.join(tmp, "d")
    run([os.path.join(REF, "dac_ref"), gguf, cf, pre, "--threads", "4", "--quiet"])
    pcm = np.stack([np.fromfile(f"{pre}.u{u}.pcm.f32", np.float32) for u in range(2)])
    np.savez_compressed(os.path.join(OUT, "dac_vectors.npz"), codes=np.stack(codes).astype(np.int32), pcm=pcm)
    print("dac vectors:", pcm.shape, "rms", float(np.sqrt((pcm ** 2).mean())))


def snac_vectors():
    from tts_cpp_b200.synth import cached_snac_gguf, synthetic_snac_codes
    gguf = cached_snac_gguf(seed=0, max_frames=64)
    codes = synthetic_snac_codes(2, 16)
    tmp = tempfile.mkdtemp()
    cf = os.path.join(tmp, "codes.txt")
    open(cf, "w").write("\n".join(" ".join(map(str, np.concatenate(c))) for c in codes) + "\n")
    pre = os.path.join(tmp, "s")
    run([os.path.join(REF, "snac_ref"), gguf, cf, pre, "--threads", "4", "--quiet"])
    pcm = np.stack([np.fromfile(f"{pre}.u{u}.pcm.f32", np.float32) for u in range(2)])
    np.savez_compressed(os.path.join(OUT, "snac_vectors.npz"), codes=np.stack([np.concatenate(c) for c in codes]).astype(np.int32), pcm=pcm)
    print("snac vectors:", pcm.shape, "rms", float(np.sqrt((pcm ** 2).mean())))


def orpheus_vectors(wide: bool = False, long: bool = False):
    """wide: head size 128 (hidden 768, a multiple of 256): the shape the tensor-core GEMV of the CUDA path accepts for every matrix.
    long (orpheus_wide_long_vectors.npz): the wide GGUF, prompts of 7 and 40 ids, 72 greedy steps -- the run crosses KV-page (32 positions) and persistent-kernel launch
    (32 steps) boundaries; the yardstick for the F16 file of the same (fp16-representable) weights through the persistent decode kernel."""
    from tts_cpp_b200.synth import cached_orpheus_gguf
    gguf = cached_orpheus_gguf(seed=0, head_dim=128) if wide else cached_orpheus_gguf(seed=0)
    rng = np.random.default_rng(5)
    prompts = [rng.integers(2, 2000, size=n) for n in ((7, 40) if long else (7, 12))]
    tmp = tempfile.mkdtemp()
    pf = os.path.join(tmp, "prompts.txt")
    open(pf, "w").write("\n".join(" ".join(map(str, q)) for q in prompts) + "\n")
    pre = os.path.join(tmp, "o")
    steps = 72 if long else 6
    run([os.path.join(REF, "orpheus_ref"), gguf, pf, pre, "--steps", str(steps), "--threads", "4", "--quiet"])
    out = {}
    for u, q in enumerate(prompts):
        out[f"prompt{u}"] = np.asarray(q, np.int32)
        out[f"tokens{u}"] = np.fromfile(f"{pre}.u{u}.tokens.i32", np.int32)
        out[f"logits{u}"] = np.fromfile(f"{pre}.u{u}.logits.f32", np.float32).reshape(steps, -1)
    np.savez_compressed(os.path.join(OUT, "orpheus_wide_long_vectors.npz" if long else "orpheus_wide_vectors.npz" if wide else "orpheus_vectors.npz"), **out)
    print("orpheus wide vectors:" if wide else "orpheus vectors:", {k: v.shape for k, v in out.items()})


def parler_vectors(f16: bool = False, quant: str | None = None, mini: bool = False):
    """mini: BASELINE config 3's model size (Parler-TTS-Mini-shaped F16 decoder: 24 layers x 1024, 16 heads x 64, ffn 4096), two prompts of 24 / 13 ids, 24 greedy
    frames -- the reference's tokens and logits at the size the benchmark runs (parler_mini_vectors.npz)."""
    from tts_cpp_b200.synth import PARLER_MINI_SHAPE, cached_parler_gguf
    gguf = cached_parler_gguf(seed=0, f16=f16, quant=quant, **(PARLER_MINI_SHAPE if mini else {}))
    rng = np.random.default_rng(7)
    prompts = [rng.integers(1, 500, size=n) for n in ((24, 13) if mini else (5, 9))]
    tmp = tempfile.mkdtemp()
    pf = os.path.join(tmp, "prompts.txt")
    open(pf, "w").write("\n".join(" ".join(map(str, q)) for q in prompts) + "\n")
    pre = os.path.join(tmp, "p")
    steps = 24 if mini else 5
    run([os.path.join(REF, "parler_ref"), gguf, pf, pre, "--steps", str(steps), "--threads", "4", "--quiet"])
    out = {}
    for u, q in enumerate(prompts):
        out[f"prompt{u}"] = np.asarray(q, np.int32)
        out[f"tokens{u}"] = np.fromfile(f"{pre}.u{u}.tokens.i32", np.int32).reshape(steps, 9)
        out[f"logits{u}"] = np.fromfile(f"{pre}.u{u}.logits.f32", np.float32).reshape(steps, 9, -1)
    tag = "_mini" if mini else (f"_{quant.lower()}" if quant else ("_f16" if f16 else ""))
    np.savez_compressed(os.path.join(OUT, f"parler{tag}_vectors.npz"), **out)
    print(f"parler{tag} vectors:", {k: v.shape for k, v in out.items()})


def parler_encoding_vectors():
    """update_conditional_prompt's second half: the Parler loop after the stored text encoding (12 rows) was replaced by another one (7 rows) through
    prep_cross_key_values(n_threads, response) -- ref_parler_driver --encoding."""
    from tts_cpp_b200.synth import cached_parler_gguf
    rng = np.random.default_rng(41)
    enc = rng.standard_normal((7, 256)).astype(np.float32)
    q = rng.integers(1, 500, size=6)
    tmp = tempfile.mkdtemp()
    ef, pf, pre = os.path.join(tmp, "enc.f32"), os.path.join(tmp, "prompts.txt"), os.path.join(tmp, "e")
    enc.tofile(ef)
    open(pf, "w").write(" ".join(map(str, q)) + "\n")
    steps = 5
    run([os.path.join(REF, "parler_ref"), cached_parler_gguf(seed=0), pf, pre, "--steps", str(steps), "--threads", "4", "--quiet", "--encoding", ef, "7"])
    np.savez_compressed(os.path.join(OUT, "parler_encoding_vectors.npz"), encoding=enc, prompt0=np.asarray(q, np.int32),
                        tokens0=np.fromfile(f"{pre}.u0.tokens.i32", np.int32).reshape(steps, 9), logits0=np.fromfile(f"{pre}.u0.logits.f32", np.float32).reshape(steps, 9, -1))
    print("parler encoding vectors written")


def parler_stop_vectors():
    """The reference's Parler loop run to completion with its stop rule (ref_parler_driver --stop) on GGUFs whose EOS logit row is boosted so that greedy
    decoding emits EOS.  x6, 7-token prompt: heads are fed EOS once they have produced one and the loop ends when every head has (13 frames).  x3, 55-token
    prompt: the loop ends at max_generation (64 positions) with some heads pinned to EOS (kept short so that the emulated CUDA path can replay it quickly)."""
    from tts_cpp_b200.synth import cached_parler_gguf
    rng = np.random.default_rng(7)
    tmp = tempfile.mkdtemp()
    out = {"step_cap": np.int32(40)}
    for name, boost, n_prompt in (("all_eos", 6.0, 7), ("max_generation", 3.0, 55)):
        q = rng.integers(1, 500, size=n_prompt)
        pf = os.path.join(tmp, f"{name}.txt")
        open(pf, "w").write(" ".join(map(str, q)) + "\n")
        pre = os.path.join(tmp, name)
        run([os.path.join(REF, "parler_ref"), cached_parler_gguf(seed=0, eos_boost=boost), pf, pre, "--steps", "40", "--threads", "4", "--quiet", "--stop"])
        out[f"{name}.prompt"] = np.asarray(q, np.int32)
        out[f"{name}.boost"] = np.float32(boost)
        out[f"{name}.tokens"] = np.fromfile(f"{pre}.u0.tokens.i32", np.int32).reshape(-1, 9)
    np.savez_compressed(os.path.join(OUT, "parler_stop_vectors.npz"), **out)
    print("parler stop vectors:", {k: v.shape for k, v in out.items()})


def dia_vectors(f16: bool = False, quant: str | None = None):
    from tts_cpp_b200.synth import cached_dia_gguf
    gguf = cached_dia_gguf(seed=0, f16=f16, quant=quant)
    rng = np.random.default_rng(11)
    prompts = [np.concatenate([[1], rng.integers(32, 127, size=n)]) for n in (9, 17)]      # [S1] + printable bytes
    tmp = tempfile.mkdtemp()
    pf = os.path.join(tmp, "prompts.txt")
    open(pf, "w").write("\n".join(" ".join(map(str, q)) for q in prompts) + "\n")
    pre = os.path.join(tmp, "d")
    steps = 5
    run([os.path.join(REF, "dia_ref"), gguf, pf, pre, "--steps", str(steps), "--threads", "4", "--quiet"])
    out = {}
    for u, q in enumerate(prompts):
        out[f"prompt{u}"] = np.asarray(q, np.int32)
        out[f"tokens{u}"] = np.fromfile(f"{pre}.u{u}.tokens.i32", np.int32).reshape(steps, 9)
        out[f"logits{u}"] = np.fromfile(f"{pre}.u{u}.logits.f32", np.float32).reshape(steps, 9, -1)
    tag = f"_{quant.lower()}" if quant else ("_f16" if f16 else "")
    np.savez_compressed(os.path.join(OUT, f"dia{tag}_vectors.npz"), **out)
    print(f"dia{tag} vectors:", {k: v.shape for k, v in out.items()})


def dia_wide_vectors():
    """dia_wide_f16_vectors.npz: the F16 GGUF with decoder width 256 (head size 64: the shape the persistent decode kernel accepts), two prompts, the whole of
    generate_from_batch's loop (64 frames: check_stopping's countdown from position 49): tokens of every frame, the reference's top-2 logit gap per (frame, head), and
    the logits of frames 0-7 and 28-35 (around the first KV-page boundary)."""
    from tts_cpp_b200.synth import cached_dia_gguf
    gguf = cached_dia_gguf(seed=0, f16=True, head_dim=64)
    rng = np.random.default_rng(13)
    prompts = [np.concatenate([[1], rng.integers(32, 127, size=n)]) for n in (9, 17)]
    tmp = tempfile.mkdtemp()
    pf = os.path.join(tmp, "prompts.txt")
    open(pf, "w").write("\n".join(" ".join(map(str, q)) for q in prompts) + "\n")
    pre = os.path.join(tmp, "d")
    run([os.path.join(REF, "dia_ref"), gguf, pf, pre, "--steps", "80", "--threads", "4", "--quiet"])
    keep = list(range(0, 8)) + list(range(28, 36))
    out = {"logit_steps": np.asarray(keep, np.int32), "step_cap": np.int32(80)}
    for u, q in enumerate(prompts):
        toks = np.fromfile(f"{pre}.u{u}.tokens.i32", np.int32).reshape(-1, 9)
        lg = np.fromfile(f"{pre}.u{u}.logits.f32", np.float32).reshape(toks.shape[0], 9, -1)
        top2 = np.sort(lg, axis=-1)[:, :, -2:]
        out[f"prompt{u}"] = np.asarray(q, np.int32); out[f"tokens{u}"] = toks; out[f"gap{u}"] = (top2[:, :, 1] - top2[:, :, 0]).astype(np.float32); out[f"logits{u}"] = lg[keep]
    np.savez_compressed(os.path.join(OUT, "dia_wide_f16_vectors.npz"), **out)
    print("dia wide f16 vectors:", {k: getattr(v, "shape", v) for k, v in out.items()})


def dia_stop_vectors():
    """The whole of generate_from_batch's loop: with dia.decoder.max_generation_size = 64 and max_delay = 15 check_stopping starts the end-of-stream
    countdown at position 49 (EOS / PAD injected head by head along the delay pattern) and ends the loop after 64 frames."""
    from tts_cpp_b200.synth import cached_dia_gguf
    gguf = cached_dia_gguf(seed=0)
    rng = np.random.default_rng(12)
    q = np.concatenate([[1], rng.integers(32, 127, size=13)])
    tmp = tempfile.mkdtemp()
    pf = os.path.join(tmp, "prompts.txt")
    open(pf, "w").write(" ".join(map(str, q)) + "\n")
    pre = os.path.join(tmp, "d")
    run([os.path.join(REF, "dia_ref"), gguf, pf, pre, "--steps", "80", "--threads", "4", "--quiet"])
    toks = np.fromfile(f"{pre}.u0.tokens.i32", np.int32).reshape(-1, 9)
    lg = np.fromfile(f"{pre}.u0.logits.f32", np.float32).reshape(toks.shape[0], 9, -1)
    np.savez_compressed(os.path.join(OUT, "dia_stop_vectors.npz"), prompt0=np.asarray(q, np.int32), tokens0=toks, logits_last0=lg[-1], step_cap=np.int32(80))
    print("dia stop vectors:", toks.shape)


def sampler_vectors():
    """The reference sampler's nucleus / probabilities / max_head_probs and a histogram of 20 000 of its draws for four configurations on fixed logits."""
    import struct
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_port import SAMPLER_CFGS
    rng = np.random.default_rng(21)
    H, V, n_draws = 3, 300, 20000
    logits = (rng.standard_normal((H, V)) * 2.5).astype(np.float32)
    last = np.array([int(np.argmax(logits[0])), 5, -1], np.int32); counts = np.array([2, 1, 0], np.uint32)     # head 0: the penalty hits its best token
    out = {"logits": logits, "last": last, "counts": counts, "n_draws": np.int32(n_draws)}
    tmp = tempfile.mkdtemp()
    for name, cfg in SAMPLER_CFGS.items():
        pin, pout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        with open(pin, "wb") as f:
            f.write(struct.pack("<IIfIff", H, V, cfg["temperature"], cfg["top_k"], cfg["top_p"], cfg["rp"]))
            f.write(last.tobytes()); f.write(counts.tobytes()); f.write(struct.pack("<I", n_draws)); f.write(logits.tobytes())
        run([os.path.join(REF, "sampler_ref"), pin, pout])
        raw = open(pout, "rb").read()
        off = 0
        for i in range(H):
            n = struct.unpack_from("<I", raw, off)[0]; off += 4
            out[f"{name}.picks{i}"] = np.frombuffer(raw, np.uint32, n, off).copy(); off += 4 * n
            out[f"{name}.probs{i}"] = np.frombuffer(raw, np.float32, n, off).copy(); off += 4 * n
            out[f"{name}.mh{i}"] = np.float32(struct.unpack_from("<f", raw, off)[0]); off += 4
        out[f"{name}.hist"] = np.frombuffer(raw, np.uint32, H * V, off).reshape(H, V).copy()
    np.savez_compressed(os.path.join(OUT, "sampler_vectors.npz"), **out)
    print("sampler vectors:", len(out), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["kokoro", "ops", "dac", "snac", "orpheus", "parler", "parler_f16", "parler_q8_0", "parler_q5_0", "parler_q4_0", "parler_stop", "parler_encoding", "dia", "dia_f16", "dia_q8_0", "dia_stop", "sampler", "orpheus_wide"]
    if "dia_stop" in which: dia_stop_vectors()
    if "parler_stop" in which: parler_stop_vectors()
    if "parler_encoding" in which: parler_encoding_vectors()
    if "sampler" in which: sampler_vectors()
    if "orpheus_wide" in which: orpheus_vectors(wide=True)
    if "dia_wide" in which: dia_wide_vectors()
    if "orpheus_wide_long" in which: orpheus_vectors(wide=True, long=True)
    if "parler_f16" in which: parler_vectors(f16=True)
    if "parler_mini" in which: parler_vectors(f16=True, mini=True)      # (not in the default list: ~2 minutes of CPU and a 1.5 GB GGUF)
    for q in ("Q8_0", "Q5_0", "Q4_0"):
        if f"parler_{q.lower()}" in which: parler_vectors(quant=q)
    if "dia_f16" in which: dia_vectors(f16=True)
    if "dia_q8_0" in which: dia_vectors(quant="Q8_0")
    if "dia" in which: dia_vectors()
    if "parler" in which: parler_vectors()
    if "orpheus" in which: orpheus_vectors()
    if "snac" in which: snac_vectors()
    if "kokoro" in which: kokoro_vectors()
    if "ops" in which: op_vectors()
    if "dac" in which: dac_vectors()
