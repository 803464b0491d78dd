"""GPU: the Parler decode loop (tts_cpp_b200/csrc/parler.cu) against the token ids and logits the compiled UNMODIFIED reference produced
(tests/golden/parler_vectors.npz: two prompts, 5 greedy frames of 9 codebooks each, small synthetic Parler GGUF).

Written after round 1's GPU budget was spent: never run on a B200 (its logic is checked under the CPU emulation, tests/test_emu_cpu.py), hence
xfail(strict=False) and a CHILD PROCESS, so that a fault in an unvalidated kernel cannot poison the CUDA context of the tests that follow.
Round 2 removes both once it has passed on hardware.

Update (end of round 1): the default greedy path ran on a B200 through scripts/rowb_first_contact.py and reproduced the reference's tokens (profiles/
r1i_rowb_first_contact.log); its test below is a plain test now, the variants that have not run yet keep xfail(strict=False) (UNRUN)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
# the F32 model on the default path (fused launches) has run on a B200 (profiles/r1i_rowb_first_contact.log: reference tokens, logits 1.5e-3); the other variants have not
UNRUN = pytest.mark.xfail(strict=False, reason="this variant of the Parler decode path has not run on a B200 yet (round 1 GPU budget exhausted)")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import parler_runner_from_file
from tts_cpp_b200.synth import cached_parler_gguf
f16 = sys.argv[2] == "f16"
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "parler_f16_vectors.npz" if f16 else "parler_vectors.npz"))
par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=f16))
prompts = [g["prompt0"], g["prompt1"]]
steps = g["tokens0"].shape[0]
toks, logits = par.generate_greedy(prompts, steps, want_logits=True)           # one ragged batch of both prompts
ok = True
for u in range(2):
    d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
    print(f"PARITY parler prompt {u}: tokens {toks[u].tolist()}  max |logit diff| {d:.3e}")
    ok &= bool(np.array_equal(toks[u], g[f"tokens{u}"])) and d < (3e-2 if f16 else 1e-2)   # bit-exact ids at temperature 0; logits: ggml's fp16 GELU table (+ fp16 activation rounding for F16 weights)
single = par.generate_greedy([prompts[1]], steps)
ok &= bool(np.array_equal(single[0], toks[1]))                                   # batching does not change a sequence
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("dtype", ["f32", pytest.param("f16", marks=UNRUN)])
def test_parler_greedy_tokens_and_logits_match_reference(dtype):
    """f16: the GGUF `quantize --quantized-type F16` writes (decoder matrices F16, activations rounded to fp16 before each such product).  Two runs
    of that model that differ only in summation order already differ by 1.5e-3 RMS / 6e-3 max in the logits (rounding boundaries), so the bar there
    is identical token ids + 3e-2."""
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, dtype], capture_output=True, text=True, timeout=150)
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0


@UNRUN
def test_parler_tensor_core_gemv_f16_matches_reference_tokens():
    """B2TTS_AR_MMA=1: the F16 matrices through gemv_mma_kernel<false> (mma.sync with the batch as M) -- same token ids as the F16 reference."""
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, "f16"], capture_output=True, text=True, timeout=150, env=dict(os.environ, B2TTS_AR_MMA="1"))
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0


STOP_CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import parler_runner_from_file
from tts_cpp_b200.synth import cached_parler_gguf
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "parler_stop_vectors.npz"))
ok = True
for case in ("all_eos", "max_generation"):
    par = parler_runner_from_file(cached_parler_gguf(seed=0, eos_boost=float(g[f"{case}.boost"])))
    ref = g[f"{case}.tokens"]
    toks, ngen = par.generate([g[f"{case}.prompt"]], int(g["step_cap"]))            # greedy, with the reference's stop rule
    good = int(ngen[0]) == ref.shape[0] and bool(np.array_equal(toks[0, :ref.shape[0]], ref)) and not toks[0, ref.shape[0]:].any()
    print(f"PARITY parler stop rule {case}: frames {int(ngen[0])} vs {ref.shape[0]} ->", good)
    ok &= good
    par.close()
sys.exit(0 if ok else 1)
'''


@UNRUN
def test_parler_stop_rule_matches_reference():
    """eos_seen feeding + check_stopping on the device against the reference run to completion (tests/golden/parler_stop_vectors.npz)."""
    r = subprocess.run([sys.executable, "-c", STOP_CHILD, ROOT], capture_output=True, text=True, timeout=150)
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0


QUANT_CHILD = r'''
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
sys.exit(0 if ok else 1)
'''


@UNRUN
@pytest.mark.parametrize("quant", ["Q8_0", "Q5_0", "Q4_0"])
def test_parler_quantised_teacher_forced(quant):
    """Block-quantised decoder matrices (gemv_rows_q_kernel), teacher-forced on the reference's tokens: logits within 0.1 RMS at every step, the same token wherever
    the reference's top-2 gap exceeds 0.5 (two correct implementations differ by ~0.04 RMS here: activation re-quantisation amplifies summation-order noise)."""
    r = subprocess.run([sys.executable, "-c", QUANT_CHILD, ROOT, quant], capture_output=True, text=True, timeout=150)
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0
