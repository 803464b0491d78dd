#!/usr/bin/env python3
"""tests/golden/make_golden_t5.py -- runs the compiled UNMODIFIED reference T5 encoder (oracle/_ref/t5_ref = src/models/parler/t5/model.cpp behind
oracle/ref_t5_driver.cpp) on synthetic T5 GGUFs (tts_cpp_b200/synth.py write_t5_gguf, deterministic in its arguments) and stores the encodings:
tests/golden/t5_vectors.npz.  Only runs where oracle/_ref exists."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tts_cpp_b200.synth import cached_t5_gguf  # noqa: E402

# name -> (GGUF arguments, prompts as token-id lists; the reference appends EOS (1) itself in generate(), the driver enters below that)
CASES = {
    "f32": (dict(), [[5, 17, 3, 90, 1], list(range(7, 41)) + [1], [9, 1], [42] * 20 + [1]]),
    "f16": (dict(f16=True), [[5, 17, 3, 90, 1], list(range(7, 41)) + [1]]),
    "no_down_proj": (dict(down_proj=False, layers=2), [[11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 1]]),
    "wide": (dict(heads=4, ffn=512, layers=2, out_size=128, context_length=128), [list(range(3, 90)) + [1]]),      # 88 tokens: every log-spaced bucket is hit
    # block-quantised layer matrices: activations re-quantised to Q8_0 per 32 columns by ggml_mul_mat (vec_dot_type), integer dot products
    "q8_0": (dict(quant="Q8_0"), [[5, 17, 3, 90, 1], list(range(7, 41)) + [1]]),
    "q4_0": (dict(quant="Q4_0"), [list(range(7, 41)) + [1]]),
    "q5_0": (dict(quant="Q5_0"), [list(range(7, 41)) + [1]]),
    # F16 matrices whose shapes the tensor-core GEMM takes (K % 64 == 0, N % 128 == 0) with enough rows to be routed there (> 32 per encode, t5.cu)
    "f16_wide": (dict(f16=True, heads=4, ffn=512, layers=2, out_size=128, context_length=128), [list(range(3, 90)) + [1], [(7 * i) % 90 + 2 for i in range(59)] + [1]]),
}


def main():
    out = {}
    for name, (kw, prompts) in CASES.items():
        g = cached_t5_gguf(**kw)
        with tempfile.TemporaryDirectory() as tmp:
            fo = os.path.join(tmp, "o.bin")
            subprocess.run([os.path.join(ROOT, "oracle", "_ref", "t5_ref"), g, fo, "4"] + [",".join(str(t) for t in p) for p in prompts], check=True,
                           stdout=subprocess.DEVNULL)
            raw = open(fo, "rb").read()
        at = 0
        for i, p in enumerate(prompts):
            n, hs = np.frombuffer(raw[at:at + 8], np.uint32); at += 8
            e = np.frombuffer(raw[at:at + 4 * int(n) * int(hs)], np.float32).reshape(int(n), int(hs)).copy(); at += 4 * int(n) * int(hs)
            assert n == len(p)
            out[f"{name}.tokens.{i}"] = np.asarray(p, np.int32)
            out[f"{name}.encoding.{i}"] = e
            print(name, i, e.shape, float(e.std()))
    np.savez_compressed(os.path.join(HERE, "t5_vectors.npz"), **out)


if __name__ == "__main__":
    main()
