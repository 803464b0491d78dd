#!/usr/bin/env python3
"""scripts/t5_timing.py -- device time of the T5 conditional-prompt encoder pass (b2tts_t5_encode) on a flan-t5-large-shaped synthetic GGUF (24 layers, hidden 1024,
16 heads, ffn 2816, down projection to 1024 -- the text encoder parler-tts-mini-v1 conditions on), F16 layer matrices.  One JSON line per (batch, tokens) point:
ms from CUDA events around the forward (b2tts_t5_last_ms, after two warm-up passes, best of 5), the weight bytes a pass streams once, and their quotient; every point also with the GEMV family only (B2TTS_T5_GEMM=0)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from tts_cpp_b200.binding import Context, lib, t5_runner_from_file  # noqa: E402
from tts_cpp_b200.synth import cached_t5_gguf  # noqa: E402
import ctypes as C  # noqa: E402


def main():
    t0 = time.time()
    g = cached_t5_gguf(f16=True, layers=24, heads=16, ffn=2816, vocab=2048, out_size=1024, context_length=512)
    t1 = time.time()
    ctx = Context(0)
    t5 = t5_runner_from_file(g, ctx=ctx)
    t2 = time.time()
    lib().b2tts_t5_weight_bytes.restype = C.c_size_t
    wbytes = int(lib().b2tts_t5_weight_bytes(t5.h))
    rng = np.random.default_rng(0)
    for B, n in ((1, 16), (1, 32), (1, 64), (8, 32), (1, 256)):
        prompts = [list(rng.integers(2, 2048, n - 1)) + [1] for _ in range(B)]
        res = {}
        for path, env in (("default", None), ("gemv_only", "0")):      # B2TTS_T5_GEMM is read at every encode: > 32 rows take the tensor-core GEMM unless it is "0"
            if env is None:
                os.environ.pop("B2TTS_T5_GEMM", None)
            else:
                os.environ["B2TTS_T5_GEMM"] = env
            ms = []
            for it in range(6):
                t5.run(prompts)
                if it >= 2:
                    ms.append(t5.last_ms())
            res[path] = (min(ms), float(np.median(ms)), int(lib().b2tts_t5_last_used_gemm(t5.h)))
        os.environ.pop("B2TTS_T5_GEMM", None)
        best = res["default"][0]
        print(json.dumps({"workload": "t5_encode", "model": "flan-t5-large-shaped synthetic, F16 layer matrices", "batch": B, "tokens_per_prompt": n, "ms": round(best, 4),
                          "ms_median": round(res["default"][1], 4), "tensor_core_gemm": bool(res["default"][2]), "ms_gemv_only": round(res["gemv_only"][0], 4),
                          "weight_bytes": wbytes, "weights_GBps": round(wbytes / best / 1e6, 1), "frac_of_6566_GBps": round(wbytes / best / 1e6 / 6566, 4),
                          "gguf_write_s": round(t1 - t0, 1), "load_s": round(t2 - t1, 1)}), flush=True)
    t5.close()


if __name__ == "__main__":
    main()
