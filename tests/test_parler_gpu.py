"""GPU: the Parler decode loop (tts_cpp_b200/csrc/parler.cu, pdk.cuh) against the token ids and logits the compiled UNMODIFIED reference produced
(tests/golden/parler*_vectors.npz: two prompts, 5 greedy frames of 9 codebooks each, small synthetic Parler GGUFs in F32 / F16 / Q8_0 / Q5_0 / Q4_0).

Paths: F16 GGUFs take the PERSISTENT DECODE KERNEL (one cooperative launch per 32 steps, paged fp16 KV cache) by default; F32 and block-quantised GGUFs, sampling and
batches above 16 take the launch-per-op path (CUDA-graph replay, tensor-core GEMV for F16 matrices).  Every variant below has passed on a B200 (gpurun_out/r2a, r2b):
plain tests.  A variant that needs a different B2TTS_* switch than the default runs in a child process (the library reads its switches once per process)."""
import pytest

from conftest import run_snippet

pytestmark = pytest.mark.gpu

BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from conftest import tie_report
from tts_cpp_b200.binding import parler_runner_from_file
from tts_cpp_b200.synth import cached_parler_gguf
f16 = sys.argv[2] == "f16"
want_pdk = sys.argv[3] == "pdk"
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "parler_f16_vectors.npz" if f16 else "parler_vectors.npz"))
par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=f16))
prompts = [g["prompt0"], g["prompt1"]]
steps = g["tokens0"].shape[0]
toks, logits = par.generate_greedy(prompts, steps, want_logits=True)           # one ragged batch of both prompts
ok = True
for u in range(2):
    d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
    print(f"PARITY parler {'f16' if f16 else 'f32'} prompt {u}: tokens {'EQUAL' if np.array_equal(toks[u], g[f'tokens{u}']) else 'DIFFER'}  max |logit diff| {d:.3e}")
    tie_report(f"parler prompt {u}", g[f"logits{u}"], g[f"tokens{u}"], logits[u], toks[u])
    ok &= bool(np.array_equal(toks[u], g[f"tokens{u}"])) and d < (3e-2 if f16 else 1e-2)   # bit-exact ids at temperature 0; logits: ggml's fp16 GELU table (+ fp16 activation rounding for F16 weights)
single = par.generate_greedy([prompts[1]], steps)                              # (no logits: the launch-per-op path replays its CUDA graph here)
ok &= bool(np.array_equal(single[0], toks[1]))                                   # batching does not change a sequence
launches, psteps = par.pdk_stats()
print("persistent-kernel launches / steps:", launches, psteps)
ok &= (psteps == 2 * steps) if want_pdk else (psteps == 0)                      # the path this variant is about is the one that ran
big = par.generate_greedy([prompts[i % 2] for i in range(18)], steps)            # 18 sequences: on the persistent path two groups (16 + 2); every sequence as in the batch of two
ok &= all(bool(np.array_equal(big[i], toks[i % 2])) for i in range(18))
par.close()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("variant", ["f32", "f16_persistent_kernel", "f16_persistent_kernel_kv_f32", "f16_persistent_kernel_small_grid_chunks_of_2", "f16_per_op_tensor_core", "f16_per_op_plain"])
def test_parler_greedy_tokens_and_logits_match_reference(variant):
    """f16: the GGUF `quantize --quantized-type F16` writes (decoder matrices F16, activations rounded to fp16 before each such product).  Two runs of that model that
    differ only in summation order already differ by 1.5e-3 RMS / 6e-3 max in the logits (rounding boundaries), so the bar there is identical token ids + 3e-2."""
    env = {"f32": None, "f16_persistent_kernel": None, "f16_persistent_kernel_kv_f32": {"B2TTS_KV": "f32"},
           "f16_persistent_kernel_small_grid_chunks_of_2": {"B2TTS_PDK_GRID": "37", "B2TTS_AR_EXIT_EVERY": "2"},
           "f16_per_op_tensor_core": {"B2TTS_AR_PDK": "0"}, "f16_per_op_plain": {"B2TTS_AR_PDK": "0", "B2TTS_AR_MMA": "0", "B2TTS_AR_GRAPH": "0"}}[variant]
    assert run_snippet(BODY, ["f32" if variant == "f32" else "f16", "pdk" if "persistent" in variant else "ops"], env=env) == 0


STOP_BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import parler_runner_from_file
from tts_cpp_b200.synth import cached_parler_gguf
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "parler_stop_vectors.npz"))
ok = True
if sys.argv[2] == "reference":
    for case in ("all_eos", "max_generation"):
        par = parler_runner_from_file(cached_parler_gguf(seed=0, eos_boost=float(g[f"{case}.boost"])))
        ref = g[f"{case}.tokens"]
        toks, ngen = par.generate([g[f"{case}.prompt"]], int(g["step_cap"]))            # greedy, with the reference's stop rule
        good = int(ngen[0]) == ref.shape[0] and bool(np.array_equal(toks[0, :ref.shape[0]], ref)) and not toks[0, ref.shape[0]:].any()
        print(f"PARITY parler stop rule {case}: frames {int(ngen[0])} vs {ref.shape[0]} ->", good)
        ok &= good
        par.close()
else:                                                                           # the same bookkeeping inside the persistent kernel (F16 GGUF) against the per-op path's dump
    par = parler_runner_from_file(cached_parler_gguf(seed=0, eos_boost=float(g["all_eos.boost"]), f16=True))
    toks, ngen = par.generate([g["all_eos.prompt"], g["max_generation.prompt"][:9]], int(g["step_cap"]))
    print("frames", ngen.tolist(), "pdk", par.pdk_stats())
    if sys.argv[2] == "dump":
        np.savez(sys.argv[3], toks=toks, ngen=ngen)
    else:
        want = np.load(sys.argv[3])
        ok &= par.pdk_stats()[1] > 0 and bool(np.array_equal(toks, want["toks"])) and bool(np.array_equal(ngen, want["ngen"])) and bool((ngen < int(g["step_cap"])).any())
    par.close()
sys.exit(0 if ok else 1)
'''


def test_parler_stop_rule_matches_reference():
    """eos_seen feeding + check_stopping on the device against the reference run to completion (tests/golden/parler_stop_vectors.npz; F32 GGUF: launch-per-op path)."""
    assert run_snippet(STOP_BODY, ["reference"]) == 0


def test_parler_stop_rule_inside_persistent_kernel(tmp_path):
    """the same stop bookkeeping inside the persistent kernel (EOS-boosted F16 GGUF, two sequences, early exit between launches) = the launch-per-op path's result."""
    f = str(tmp_path / "ops.npz")
    assert run_snippet(STOP_BODY, ["dump", f], env={"B2TTS_AR_PDK": "0"}) == 0
    assert run_snippet(STOP_BODY, ["check", f]) == 0


QUANT_BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import parler_runner_from_file
from tts_cpp_b200.synth import cached_parler_gguf
quant = sys.argv[2]
g = np.load(os.path.join(sys.argv[1], "tests", "golden", f"parler_{quant.lower()}_vectors.npz"))
par = parler_runner_from_file(cached_parler_gguf(seed=0, quant=quant))
prompts = [g["prompt0"], g["prompt1"]]
teacher = np.stack([g["tokens0"], g["tokens1"]])
toks, logits = par.generate_teacher_forced(prompts, teacher)
ok = True
for u in range(2):
    ref_t, ref_l = g[f"tokens{u}"], g[f"logits{u}"]
    rms = np.sqrt(((logits[u] - ref_l) ** 2).mean(axis=(1, 2)))
    top2 = np.sort(ref_l, axis=2)[:, :, -2:]
    clear = (top2[:, :, 1] - top2[:, :, 0]) > 0.5
    print(f"PARITY parler {quant} prompt {u}: per-step logit rms {np.round(rms, 4).tolist()}, tokens equal {int((toks[u] == ref_t).sum())}/{ref_t.size}")
    ok &= float(rms.max()) < 0.1 and bool(np.array_equal(toks[u][clear], ref_t[clear]))
par.close()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("quant", ["Q8_0", "Q5_0", "Q4_0"])
def test_parler_quantised_teacher_forced(quant):
    """Block-quantised decoder matrices (gemv_rows_q_kernel), teacher-forced on the reference's tokens: logits within 0.1 RMS at every step, the same token wherever
    the reference's top-2 gap exceeds 0.5 (two correct implementations differ by ~0.04 RMS here: activation re-quantisation amplifies summation-order noise)."""
    assert run_snippet(QUANT_BODY, [quant]) == 0
