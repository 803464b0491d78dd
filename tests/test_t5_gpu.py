"""GPU: the T5 conditional-prompt encoder (t5.cu) through the C-ABI (b2tts_t5_load_gguf / b2tts_t5_encode) against the compiled unmodified
reference (t5_runner::run; tests/golden/t5_vectors.npz made by tests/golden/make_golden_t5.py), and chained into Parler: T5 encoding -> b2tts_parler_set_text_encoding.
Collected last (conftest._LATE): new in this round's last session; validated under tests/emu before its first hardware run."""
import os
import sys

import numpy as np
import pytest

from conftest import CACHE, ROOT, report

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

def _T5_TOL(case):
    """relative RMS bar of a T5 golden case: F32 1e-4 (fp16 GELU table); F16 1.5e-3 (fp16 activation rounding); block-quantised 3e-2 -- ggml_mul_mat re-quantises the
    activations to Q8_0 per 32 columns, so 1e-7 of summation-order noise moves whole quantisation steps (two correct implementations: 1e-7 on one prompt, 1e-2 on the next)"""
    return 3e-2 if case.startswith("q") else 1.5e-3 if case.startswith("f16") else 1e-4



@pytest.mark.parametrize("case", ["f32", "f16", "no_down_proj", "wide", "f16_wide", "q8_0", "q5_0", "q4_0"])
def test_t5_encode_matches_reference(gpu_ctx, case):
    """Every golden case as one ragged batch and prompt by prompt.  "f16" (40 rows batched) and "f16_wide" (148 rows, 88 and 60 alone) exceed the 32 rows above which
    F16 matrices go through the tensor-core GEMM (t5.cu T5_GEMM_MIN_ROWS): the same numerics class as the reference's F16 mul_mat (fp16-rounded activations, exact
    products, fp32 accumulation), another summation order."""
    import make_golden_t5 as M
    from tts_cpp_b200.binding import lib, t5_runner_from_file
    from tts_cpp_b200.synth import cached_t5_gguf
    g = np.load(os.path.join(ROOT, "tests", "golden", "t5_vectors.npz"))
    kw, prompts = M.CASES[case]
    t5 = t5_runner_from_file(cached_t5_gguf(cache_dir=CACHE, **kw), ctx=gpu_ctx)
    try:
        batch = t5.run(prompts)                                # the case's prompts as one ragged batch
        for i, p in enumerate(prompts):
            ref = g[f"{case}.encoding.{i}"]
            d, r, mx = report(f"t5 {case}.{i} (batched)", batch[i], ref)
            assert batch[i].shape == ref.shape and d < _T5_TOL(case) * r, (case, i, d, r)
            one = t5.run([p])[0]                               # and alone: a prompt's encoding does not depend on its batch
            if case.startswith(("f16", "q")):                  # ... up to the summation order where the row count picks another kernel (F16: GEMV or GEMM; quantised: the per-row-count tiles)
                d1, r1, _ = report(f"t5 {case}.{i} (alone)", one, ref)
                assert d1 < _T5_TOL(case) * r1, (case, i, d1, r1)
            else:
                assert np.array_equal(one, batch[i]), (case, i)
        if case == "f16_wide":
            assert lib().b2tts_t5_last_used_gemm(t5.h) == 1    # 60 rows alone: the GEMM path ran
        with pytest.raises(RuntimeError):
            t5.run([[5, t5.vocab_size, 1]])
        with pytest.raises(RuntimeError):
            t5.run([list(range(2, 2 + t5.context_length)) + [1]])
    finally:
        t5.close()


def test_t5_gemm_path_agrees_with_gemv_path_at_the_flan_t5_large_width(gpu_ctx):
    """No reference run at this size in the goldens (hidden 1024, ffn 2816: the tile shapes the real text encoder uses), so a size-independent property: the same
    200-token prompt through the tensor-core GEMM and -- B2TTS_T5_GEMM=0, read at every call -- through the GEMV family, two implementations of the same F16 numerics
    that share no matrix kernel, agree at the F16 rounding floor."""
    from tts_cpp_b200.binding import lib, t5_runner_from_file
    from tts_cpp_b200.synth import cached_t5_gguf
    t5 = t5_runner_from_file(cached_t5_gguf(cache_dir=CACHE, f16=True, layers=2, heads=16, ffn=2816, vocab=512, out_size=1024, context_length=256), ctx=gpu_ctx)
    prompt = [(13 * i) % 500 + 2 for i in range(199)] + [1]
    try:
        a = t5.run([prompt])[0]
        assert lib().b2tts_t5_last_used_gemm(t5.h) == 1
        os.environ["B2TTS_T5_GEMM"] = "0"
        b = t5.run([prompt])[0]
        assert lib().b2tts_t5_last_used_gemm(t5.h) == 0
    finally:
        os.environ.pop("B2TTS_T5_GEMM", None)
        t5.close()
    d, r, mx = report("t5 GEMM vs GEMV path, 200 x 1024", a, b)
    assert np.isfinite(a).all() and d < 1.5e-3 * r, (d, r)


def test_t5_encoding_feeds_parler_cross_attention(gpu_ctx):
    """update_conditional_prompt as a whole on the device: T5 encode -> b2tts_parler_set_text_encoding -> decode.  The tokens equal those of the same decode with
    the same encoding handed over from the host (the path tests/test_parler_gpu.py pins to the reference), and differ from the stored-encoding run."""
    from tts_cpp_b200.binding import parler_runner_from_file, t5_runner_from_file
    from tts_cpp_b200.synth import cached_parler_gguf, cached_t5_gguf
    t5 = t5_runner_from_file(cached_t5_gguf(cache_dir=CACHE), ctx=gpu_ctx)          # output_size 256 = the toy Parler's hidden (32 heads x 8)
    parler = parler_runner_from_file(cached_parler_gguf(cache_dir=CACHE), ctx=gpu_ctx)
    assert t5.output_size == parler.hidden_size
    prompts = [[3, 17, 250, 9], [44, 2, 300]]
    base = parler.generate_greedy(prompts, 6)
    enc = t5.run([[5, 17, 3, 90, 60, 61, 62, 1]])[0]
    parler.set_text_encoding(enc)
    a = parler.generate_greedy(prompts, 6)
    parler.set_text_encoding(enc.copy())
    b = parler.generate_greedy(prompts, 6)
    assert np.array_equal(a, b) and not np.array_equal(a, base)
    t5.close()
