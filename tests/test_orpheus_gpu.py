"""GPU: the Orpheus decode loop (tts_cpp_b200/csrc/orpheus.cu) against the token ids and logits the compiled UNMODIFIED reference produced
(tests/golden/orpheus_vectors.npz: two prompts, 6 greedy steps each, small synthetic Orpheus GGUF; orpheus_wide_vectors.npz: hidden 768).
All variants have passed on a B200 (GPUTEST_r01, gpurun_out/r2a): plain tests.  Variants that need a different B2TTS_* switch than the default run in a child
process (the library reads its switches once per process)."""
import pytest

from conftest import run_snippet

pytestmark = pytest.mark.gpu

BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from conftest import tie_report
from tts_cpp_b200.binding import orpheus_runner_from_file
from tts_cpp_b200.synth import cached_orpheus_gguf
wide = sys.argv[2] == "wide"
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "orpheus_wide_vectors.npz" if wide else "orpheus_vectors.npz"))
orph = orpheus_runner_from_file(cached_orpheus_gguf(seed=0, head_dim=128) if wide else cached_orpheus_gguf(seed=0))
prompts = [g["prompt0"], g["prompt1"]]
steps = g["tokens0"].size
toks, logits = orph.generate_greedy(prompts, steps, want_logits=True)         # one ragged batch of both prompts
ok = True
for u in range(2):
    d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
    print(f"PARITY orpheus{' wide' if wide else ''} (B2TTS_AR_MMA={os.environ.get('B2TTS_AR_MMA')}) prompt {u}: tokens {toks[u].tolist()} vs {g[f'tokens{u}'].tolist()}  max |logit diff| {d:.3e}")
    tie_report(f"orpheus prompt {u}", g[f"logits{u}"][:, None, :], g[f"tokens{u}"][:, None], logits[u][:, None, :], toks[u][:, None])
    ok &= bool(np.array_equal(toks[u], g[f"tokens{u}"])) and d < (1e-4 if wide else 1e-3)          # bit-exact token ids at temperature 0
single = orph.generate_greedy([prompts[1]], steps)
ok &= bool(np.array_equal(single[0], toks[1]))                                  # batching does not change a sequence; this run replays the CUDA graph (no logits requested)
orph.close()
sys.exit(0 if ok else 1)
'''


def test_orpheus_greedy_tokens_and_logits_match_reference():
    assert run_snippet(BODY, ["small"]) == 0


@pytest.mark.parametrize("mma", [None, "1"], ids=["plain", "split_mma"])
def test_orpheus_wide_tokens_and_logits_match_reference(mma):
    """hidden 768 (every matrix eligible for the tensor-core GEMV); split_mma (B2TTS_AR_MMA=1): the fp32-faithful three-product path over fp16 (hi, lo) pairs."""
    assert run_snippet(BODY, ["wide"], env=None if mma is None else {"B2TTS_AR_MMA": mma}) == 0


ALT_BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import orpheus_runner_from_file
from tts_cpp_b200.synth import cached_orpheus_gguf
kind = sys.argv[2]
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "orpheus_vectors.npz"))
orph = orpheus_runner_from_file(cached_orpheus_gguf(seed=0, quant="Q8_0") if kind == "q8_0" else cached_orpheus_gguf(seed=0, f16=True))
prompts = [g["prompt0"], g["prompt1"]]
steps = g["tokens0"].size
toks, logits = orph.generate_greedy(prompts, steps, want_logits=True)
ok = True
for u in range(2):
    ref = g[f"logits{u}"]
    rel = float(np.sqrt(((logits[u] - ref) ** 2).mean()) / ref.std())
    top2 = np.sort(ref, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 8.0 * np.abs(logits[u] - ref).max(axis=1)
    print(f"PARITY orpheus {kind} prompt {u}: logit rms / std {rel:.3e} vs the reference's F32 run; tokens equal {int((toks[u] == g[f'tokens{u}']).sum())}/{steps}; clear decisions {int(clear.sum())}")
    ok &= rel < (0.05 if kind == "q8_0" else 0.01) and bool(np.array_equal(toks[u][clear], g[f"tokens{u}"][clear]))
replay = orph.generate_greedy(prompts, steps)                                # graph replay, no logits
ok &= bool(np.array_equal(replay, toks))
orph.close()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("kind", ["q8_0", "f16"])
def test_orpheus_q8_0_and_f16_matrices_track_the_f32_reference(kind):
    """BASELINE config 5 runs Orpheus as q8_0; the reference cannot (its quantize tool refuses Orpheus, its runtime is F32-only), so the yardstick is the reference's F32
    run of the same weights: logits within the storage format's noise (Q8_0: 5 % of the logit std, measured 1.7 %; F16: 1 %) and the same token wherever the F32
    top-2 gap exceeds 8x the step's largest logit difference."""
    assert run_snippet(ALT_BODY, [kind]) == 0


STOP_BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import orpheus_runner_from_file
from tts_cpp_b200.synth import cached_orpheus_gguf
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "orpheus_vectors.npz"))
orph = orpheus_runner_from_file(cached_orpheus_gguf(seed=0))
prompts = [g["prompt0"], g["prompt1"]]
ref0, ref1 = g["tokens0"], g["tokens1"]
stop = int(ref0[2])                                                           # the reference's third greedy token of sequence 0 becomes the stopping token
orph.set_stopping_token(stop)
toks, ngen = orph.generate_until_stop(prompts, 40)
want0 = 3
want1 = next((i + 1 for i, t in enumerate(toks[1]) if t == stop), 40)
print("n_generated", ngen.tolist(), "expected", [want0, want1])
ok = int(ngen[0]) == want0 and bool(np.array_equal(toks[0, :3], ref0[:3])) and not toks[0, 3:].any()
ok &= int(ngen[1]) == want1 and bool(np.array_equal(toks[1, :min(want1, ref1.size)], ref1[:min(want1, ref1.size)]))
orph.close()
sys.exit(0 if ok else 1)
'''


def test_orpheus_stop_rule():
    """generate_from_batch's stop condition (reference src/models/orpheus/model.cpp:389-398) on the device: the loop ends at the stopping token per sequence, the batch
    stops stepping once every sequence has ended, tokens up to the stop are the reference's."""
    assert run_snippet(STOP_BODY, []) == 0


PDK_BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import orpheus_runner_from_file
from tts_cpp_b200.synth import cached_orpheus_gguf
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "orpheus_wide_long_vectors.npz"))
q8 = len(sys.argv) > 2 and sys.argv[2] == "q8_0"
orph = orpheus_runner_from_file(cached_orpheus_gguf(seed=0, head_dim=128, quant="Q8_0") if q8 else cached_orpheus_gguf(seed=0, head_dim=128, f16=True))
pdk_on = os.environ.get("B2TTS_AR_PDK") != "0"
prompts = [g["prompt0"], g["prompt1"]]
steps = int(g["tokens0"].size)
toks_b, logits_b = orph.generate_greedy(prompts, steps, want_logits=True)     # one ragged batch: 7- and 40-token prompts
launches, psteps = orph.pdk_stats()
print("persistent-kernel launches / steps", launches, psteps)
ok = (psteps == steps - 1 and launches == -(-(steps - 1) // 32)) if pdk_on else psteps == 0
for u in range(2):
    ref_t, ref_l = g[f"tokens{u}"], g[f"logits{u}"]
    # free-running greedy decoding against the reference's F32 run of the same weights: the first differing token must be a near-tie of the reference (top-2 gap within
    # 2x the step's largest logit difference); the run is then re-anchored on the reference's tokens (prompt + its tokens so far) and the comparison continues
    done, prompt, got_t, got_l, anchors, worst = 0, prompts[u], toks_b[u], logits_b[u], 0, 0.0
    while done < steps:
        n = steps - done
        d = np.abs(got_l[:n] - ref_l[done:]).max(axis=1)
        neq = np.nonzero(got_t[:n] != ref_t[done:])[0]
        upto = int(neq[0]) if neq.size else n - 1
        worst = max(worst, float(d[:upto + 1].max()))
        if not neq.size: break
        s = done + upto
        top2 = np.sort(ref_l[s])[-2:]
        gap = float(top2[1] - top2[0])
        print(f"TIE-MARGIN orpheus {'q8_0' if q8 else 'f16'} pdk={int(pdk_on)} prompt {u} step {s}: token {int(got_t[upto])} vs {int(ref_t[s])}, reference top-2 gap {gap:.3e}, max |logit diff| {float(d[upto]):.3e}")
        if gap > 2.0 * float(d[upto]): ok = False; print("  CLEAR DECISION DIFFERS"); break
        anchors += 1
        done = s + 1
        if done >= steps - 1: break
        prompt = np.concatenate([prompts[u], ref_t[:done]]).astype(np.uint32)
        t2, l2 = orph.generate_greedy([prompt], steps - done, want_logits=True)
        got_t, got_l = t2[0], l2[0]
    print(f"PARITY orpheus wide {'Q8_0' if q8 else 'F16'} pdk={int(pdk_on)} prompt {u}: {steps} steps, {anchors} near-tie re-anchorings, max |logit diff| vs the reference's F32 run {worst:.3e}")
    ok &= worst < (0.8 if q8 else 5e-2) and anchors <= (24 if q8 else 4)      # Q8_0: format noise ~2 % of the logit std (4) on both paths, near-ties are frequent
    single = orph.generate_greedy([prompts[u]], steps)                          # batching does not change a sequence
    ok &= bool(np.array_equal(single[0], toks_b[u]))
# 18 sequences: two groups (16 + 2) through the persistent kernel, each sequence as in the batch of two
big = orph.generate_greedy([prompts[i % 2] for i in range(18)], 12)
ok &= all(bool(np.array_equal(big[i], toks_b[i % 2][:12])) for i in range(18))
print("18-sequence call equals the 2-sequence batch per sequence:", ok)
# the stop rule inside the persistent kernel: sequence 0's third token becomes the stopping token
stop = int(toks_b[0][2])
orph.set_stopping_token(stop)
toks, ngen = orph.generate_until_stop(prompts, steps)
want = [next((i + 1 for i, t in enumerate(toks_b[u]) if t == stop), steps) for u in range(2)]
print("n_generated", ngen.tolist(), "expected", want)
for u in range(2):
    ok &= int(ngen[u]) == want[u] and bool(np.array_equal(toks[u, :want[u]], toks_b[u][:want[u]])) and not toks[u, want[u]:].any()
orph.close()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("variant", ["pdk_f16kv", "pdk_f32kv", "per_op"])
def test_orpheus_f16_persistent_kernel_tracks_the_reference(variant):
    """The F16 file of the wide GGUF (its weights are fp16-representable, so the reference's F32 run of the same values is the yardstick:
    tests/golden/orpheus_wide_long_vectors.npz, 72 greedy steps, prompts of 7 and 40 ids) through the persistent decode kernel (pdk.cuh: RMSNorm in the staging, NeoX RoPE +
    cache append and SwiGLU in the epilogues, GQA attention over fp16 / fp32 pages, argmax partials), across KV-page and launch boundaries; the launch-per-op path under
    the same rule; the stop rule; batch invariance."""
    env = {"pdk_f16kv": None, "pdk_f32kv": {"B2TTS_KV": "f32"}, "per_op": {"B2TTS_AR_PDK": "0"}}[variant]
    assert run_snippet(PDK_BODY, [], env=env) == 0


@pytest.mark.parametrize("variant", ["pdk", "per_op"])
def test_orpheus_q8_0_persistent_kernel_tracks_the_reference(variant):
    """BASELINE config 5's dtype through the persistent decode kernel (int8 MMA over Q8_0-quantised activations: ggml_vec_dot_q8_0_q8_0's arithmetic) and through the
    launch-per-op dp4a path, 72 greedy steps each against the reference's F32 run of the same weights: logits within the Q8_0 format noise, every differing token a near-tie
    of the reference (the run is re-anchored on the reference's tokens after each), stop rule, batch invariance."""
    assert run_snippet(PDK_BODY, ["q8_0"], env=None if variant == "pdk" else {"B2TTS_AR_PDK": "0"}) == 0
