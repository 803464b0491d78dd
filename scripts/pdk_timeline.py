"""scripts/pdk_timeline.py -- where one decode step of the persistent kernel spends its time: run Parler-Mini F16 (BASELINE config 3 shape, batch 16) for N steps with
B2TTS_PDK_PROF=<step> and summarise the %globaltimer timeline (every op x every CTA: op begin, activations staged, barrier entered, barrier left).
    python scripts/pdk_timeline.py [steps=480] [prof_step=450] [parler|orpheus|dia] [batch=16] > profiles/r2x_pdk_timeline.txt"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 480
    pstep = int(sys.argv[2]) if len(sys.argv) > 2 else 450
    pf = "/tmp/pdk_prof.bin"
    os.environ["B2TTS_PDK_PROF"] = str(pstep); os.environ["B2TTS_PDK_PROF_FILE"] = pf
    model = sys.argv[3] if len(sys.argv) > 3 else "parler"
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    rng = np.random.default_rng(5)
    if model == "orpheus":                                 # BASELINE config 5's decoder shape, F16 matrices, random weights handed over tensor by tensor
        from tts_cpp_b200.binding import Context
        from tts_cpp_b200.synth import build_orpheus_direct
        par = build_orpheus_direct(Context(0), dtype=os.environ.get("B2TTS_TIMELINE_DTYPE", "f16"))
        prompts = [rng.integers(1, 100000, size=40).astype(np.uint32) for _ in range(batch)]
    elif model == "dia":                                   # BASELINE config 4's model shape, F16, `batch` utterances (each a CFG row pair)
        from tts_cpp_b200.binding import Context
        from tts_cpp_b200.synth import build_dia_direct
        par = build_dia_direct(Context(0), dtype="f16")
        prompts = [np.concatenate([[1], rng.integers(32, 127, size=62), [2], rng.integers(32, 127, size=64)]).astype(np.uint32) for _ in range(batch)]
    else:
        from tts_cpp_b200.binding import parler_runner_from_file
        from tts_cpp_b200.synth import PARLER_MINI_SHAPE, cached_parler_gguf
        par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=True, **PARLER_MINI_SHAPE))
        prompts = [rng.integers(1, 500, size=24).astype(np.uint32) for _ in range(batch)]
    par.generate_greedy(prompts, 40)                       # warm-up
    par.generate_greedy(prompts, steps)
    print(f"# { {'orpheus': 'Orpheus-3B-shaped', 'dia': 'Dia-1.6B-shaped'}.get(model, 'Parler-Mini') } F16, batch {batch}, {steps} steps: {par.last_ms() / steps:.4f} ms per step (incl. the prompt pass); timeline of step {pstep}; switches "
          f"{ {k: os.environ.get(k) for k in ('B2TTS_KV', 'B2TTS_PDK_GRID')} }")
    raw = open(pf, "rb").read()
    n_ops, G, W, ps = struct.unpack("iiii", raw[:16])
    kinds = np.frombuffer(raw, np.int32, n_ops * 4, 16).reshape(n_ops, 4)
    t = np.frombuffer(raw, np.uint64, n_ops * G * W, 16 + n_ops * 16).reshape(n_ops, G, W).astype(np.int64)
    names = {0: "rows+embed", 1: "gemv", 2: "attention", 3: "argmax", 4: "attn-combine", 5: "quantise"}
    t0 = t[:, :, 0].min()
    print(f"# step wall (first op begin -> last barrier left): {(t[:, :, 3].max() - t0) / 1e3:.1f} us over {n_ops} ops on {G} CTAs")
    print("# per op: wall = max_cta(barrier left) - min_cta(op begin); work = median over CTAs of (barrier entered - op begin); stage = median(staged - begin) [gemv]; "
          "bar = min over CTAs of (barrier left - barrier entered) = the barrier's own latency for the LAST arriver")
    agg = {}
    for oi in range(n_ops):
        k, layer, K, nu = kinds[oi]
        wall = (t[oi, :, 3].max() - t[oi, :, 0].min()) / 1e3
        work = np.median(t[oi, :, 2] - t[oi, :, 0]) / 1e3
        workmax = (t[oi, :, 2] - t[oi, :, 0]).max() / 1e3
        stage = np.median(t[oi, :, 1] - t[oi, :, 0]) / 1e3 if k == 1 else 0.0
        bar = (t[oi, :, 3] - t[oi, :, 2]).min() / 1e3
        wwait = np.median(t[oi, :, 4]) / 1e3 if W > 4 else 0.0
        nready = np.median(t[oi, :, 5] - t[oi, :, 0]) / 1e3 if (W > 4 and k == 1 and t[oi, :, 5].max() > 0) else 0.0
        key = (names[k], int(K), int(nu)) if k == 1 else (names[k], 0, 0)
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += wall; a[2] += work; a[3] += workmax; a[4] += stage; a[5] += bar; a[6] += wwait; a[7] += nready
        if oi < 12:
            print(f"  op {oi:3d} {names[k]:10s} layer {layer:2d} K {K:5d} units {nu:5d}: wall {wall:7.2f} us  work med {work:7.2f} max {workmax:7.2f}  stage {stage:6.2f}  barrier {bar:5.2f}  tile-wait {wwait:5.2f}  norm-tile-ready {nready:5.2f}")
    print("# totals per op class (us per step): count, wall, median work, max work, stage, barrier")
    tot = 0.0
    for key, a in sorted(agg.items()):
        print(f"  {str(key):34s} x{a[0]:3d}  wall {a[1]:8.1f}  work {a[2]:8.1f}  workmax {a[3]:8.1f}  stage {a[4]:7.1f}  barrier {a[5]:7.1f}  tile-wait {a[6]:7.1f}  norm-tile-ready {a[7]:7.1f}")
        tot += a[1]
    print(f"# sum of walls {tot:.1f} us")


if __name__ == "__main__":
    main()
