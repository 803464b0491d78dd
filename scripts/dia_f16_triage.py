"""scripts/dia_f16_triage.py -- why does the Dia F16 greedy run differ from the reference on a B200 when the CPU emulation of the same .cu passes?

Prints, for the F16 Dia GGUF of tests/golden/dia_f16_vectors.npz: the free-running tokens against the reference's step by step, the logit differences, and --
teacher-forced on the reference's tokens -- the per-step logit RMS / max difference next to the reference's own top-2 gap for every (step, head), i.e. whether a
differing token sits on a near-tie (rounding noise multiplied by the CFG gain of 4) or on a clear decision (a bug).  Run once per switch setting."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tts_cpp_b200.binding import dia_runner_from_file
    from tts_cpp_b200.synth import cached_dia_gguf
    print("switches", {k: os.environ.get(k) for k in ("B2TTS_AR_FUSE", "B2TTS_AR_GRAPH", "B2TTS_AR_MMA", "B2TTS_AR_ATT", "B2TTS_GEMV_GN")})
    for f16 in (False, True):
        g = np.load(os.path.join(ROOT, "tests", "golden", "dia_f16_vectors.npz" if f16 else "dia_vectors.npz"))
        dia = dia_runner_from_file(cached_dia_gguf(seed=0, f16=f16))
        prompts = [g["prompt0"], g["prompt1"]]
        steps = g["tokens0"].shape[0]
        toks, ngen, logits = dia.generate_greedy(prompts, steps, want_logits=True)
        tf_t, tf_l = dia.generate_teacher_forced(prompts, np.stack([g["tokens0"], g["tokens1"]]))
        for u in range(2):
            rt, rl = g[f"tokens{u}"], g[f"logits{u}"]
            print(f"== {'f16' if f16 else 'f32'} prompt {u}: free-running tokens equal: {bool(np.array_equal(toks[u], rt))}; logit std {rl.std():.2f}")
            for s in range(steps):
                fl = np.asarray(logits[u][s]).reshape(rl[s].shape)
                tl = np.asarray(tf_l[u][s]).reshape(rl[s].shape)
                top2 = np.sort(rl[s], axis=-1)[:, -2:]
                gap = top2[:, 1] - top2[:, 0]
                bad_free = np.nonzero(np.asarray(toks[u][s]) != rt[s])[0].tolist()
                bad_tf = np.nonzero(np.asarray(tf_t[u][s]) != rt[s])[0].tolist()
                print(f"   step {s}: free max|d| {np.abs(fl - rl[s]).max():.3e}  teacher-forced rms {np.sqrt(((tl - rl[s]) ** 2).mean()):.3e} max {np.abs(tl - rl[s]).max():.3e}  "
                      f"min top-2 gap {gap.min():.3e}  heads differing free {bad_free} (gaps {[round(float(gap[h]), 4) for h in bad_free]}) teacher-forced {bad_tf} (gaps {[round(float(gap[h]), 4) for h in bad_tf]})")
        dia.close()


if __name__ == "__main__":
    main()
