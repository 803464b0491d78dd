"""scripts/pdk_gpu_check.py -- first hardware contact of the persistent decode kernel: the F16 Parler goldens (tokens + logits of the compiled reference) through the default
path (persistent kernel, paged fp16 KV), teacher-forced, and the stop rule against the launch-per-op path run in a child process."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tts_cpp_b200.binding import parler_runner_from_file
    from tts_cpp_b200.synth import cached_parler_gguf
    print("switches", {k: os.environ.get(k) for k in ("B2TTS_AR_PDK", "B2TTS_KV", "B2TTS_PDK_GRID", "B2TTS_AR_EXIT_EVERY")}, flush=True)
    g = np.load(os.path.join(ROOT, "tests", "golden", "parler_f16_vectors.npz"))
    par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=True))
    prompts = [g["prompt0"], g["prompt1"]]
    steps = g["tokens0"].shape[0]
    t0 = time.time()
    toks, logits = par.generate_greedy(prompts, steps, want_logits=True)
    print(f"generate_greedy {time.time() - t0:.3f} s wall, device {par.last_ms():.3f} ms, pdk stats {par.pdk_stats()}", flush=True)
    ok = True
    for u in range(2):
        d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
        same = bool(np.array_equal(toks[u], g[f"tokens{u}"]))
        top2 = np.sort(g[f"logits{u}"], axis=-1)[..., -2:]
        print(f"PARITY parler f16 prompt {u}: tokens {'EQUAL' if same else 'DIFFER'}  max |logit diff| {d:.3e}  min top-2 gap {float((top2[..., 1] - top2[..., 0]).min()):.3e}", flush=True)
        ok &= same and d < 3e-2
    t2 = par.generate_greedy(prompts, steps)                 # without logits
    ok &= bool(np.array_equal(t2, toks))
    single = par.generate_greedy([prompts[1]], steps)
    ok &= bool(np.array_equal(single[0], toks[1]))
    print("no-logits run and single-sequence run equal:", ok, flush=True)
    tf_t, tf_l = par.generate_teacher_forced(prompts, np.stack([g["tokens0"], g["tokens1"]]))
    for u in range(2):
        ok &= bool(np.array_equal(tf_t[u], g[f"tokens{u}"]))
    print("teacher-forced tokens equal:", ok, " pdk stats", par.pdk_stats(), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
