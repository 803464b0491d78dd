"""scripts/rowb_first_contact.py -- the shortest possible first run of the row-B decode paths on a B200 (used with the last seconds of round 1's GPU budget, and the
first thing to run in round 2): Parler F32 (default fused launches) and Orpheus against the reference's tokens / logits (tests/golden), then the sampler's default
configuration against the pinned port.  Every result is printed and appended to gpurun_out/rowb_first_contact.log as soon as it exists."""
import os
import sys
import time

import numpy as np

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "rowb_first_contact.log"), "a")


def log(*a):
    line = f"[{time.time() - T0:6.2f}s] " + " ".join(str(x) for x in a)
    print(line, flush=True)
    LOG.write(line + "\n"); LOG.flush(); os.fsync(LOG.fileno())


def main():
    from tts_cpp_b200 import binding, synth
    gold = os.path.join(ROOT, "tests", "golden")
    log("switches", {k: os.environ.get(k) for k in ("B2TTS_AR_FUSE", "B2TTS_AR_GRAPH", "B2TTS_AR_MMA", "B2TTS_AR_ATT")})
    for model in ("parler", "orpheus", "dia"):
        g = np.load(os.path.join(gold, f"{model}_vectors.npz"))
        runner = getattr(binding, f"{model}_runner_from_file")(getattr(synth, f"cached_{model}_gguf")(seed=0))
        log(model, "loaded")
        prompts = [g["prompt0"], g["prompt1"]]
        steps = int(g["tokens0"].shape[0])
        out = runner.generate_greedy(prompts, steps, want_logits=True)
        toks, logits = out[0], out[-1]
        for u in range(2):
            ref_t, ref_l = g[f"tokens{u}"], g[f"logits{u}"]
            same = bool(np.array_equal(np.asarray(toks[u]).reshape(ref_t.shape), ref_t))
            d = float(np.abs(np.asarray(logits[u]).reshape(ref_l.shape) - ref_l).max())
            log(f"PARITY {model} prompt {u}: tokens {'EQUAL' if same else 'DIFFER'}  max |logit diff| {d:.3e}  decode {runner.last_ms():.2f} ms")
    import ctypes as C
    from oracle.sampler_port import SamplerPort, uniform_from_counter
    ctx = binding.Context(0)
    rng = np.random.default_rng(33)
    rows, V, seed = 18, 1088, 0x1234ABCD5678
    logits = (rng.standard_normal((rows, V)) * 2.5).astype(np.float32)
    toks = np.empty(rows, np.int32)
    binding._chk(binding.lib().b2tts_op_sample(ctx.h, logits.ctypes.data_as(C.POINTER(C.c_float)), rows, V, 1, 50, C.c_float(1.0), C.c_float(1.0), C.c_float(1.0), None, None,
                                                 C.c_uint64(seed), 0, toks.ctypes.data_as(C.POINTER(C.c_int32))))
    us = np.array([uniform_from_counter(seed, r, 0) for r in range(rows)], np.float32)
    want = SamplerPort(rows, V, 1.0, 50, 1.0, 1.0).draw(logits, us)
    log("PARITY sampler top-50:", "EQUAL" if np.array_equal(toks, want) else f"DIFFER {toks.tolist()} vs {want.tolist()}")


if __name__ == "__main__":
    try:
        main()
    except BaseException as e:      # noqa: BLE001 -- the log must say what stopped the run
        log("STOPPED:", type(e).__name__, e)
        raise
