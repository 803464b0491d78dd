"""GPU: the Dia decode loop (tts_cpp_b200/csrc/dia.cu) against the token ids and logits the compiled UNMODIFIED reference produced
(tests/golden/dia_vectors.npz: two byte-token prompts, 5 greedy frames of 9 codebooks each with the CFG-combined logits, small synthetic Dia GGUF).

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
# the F32 model on the default path (fused launches) has run on a B200 (profiles/r1i_rowb_first_contact.log: reference tokens, logits 3.5e-3); the other variants have not
UNRUN = pytest.mark.xfail(strict=False, reason="this variant of the Dia decode path has not run on a B200 yet (round 1 GPU budget exhausted)")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import dia_runner_from_file
from tts_cpp_b200.synth import cached_dia_gguf
f16 = sys.argv[2] == "f16"
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "dia_f16_vectors.npz" if f16 else "dia_vectors.npz"))
par = dia_runner_from_file(cached_dia_gguf(seed=0, f16=f16))
prompts = [g["prompt0"], g["prompt1"]]
steps = g["tokens0"].shape[0]
toks, ngen, logits = par.generate_greedy(prompts, steps, want_logits=True)           # one ragged batch of both prompts
ok = True
for u in range(2):
    d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
    print(f"PARITY dia prompt {u}: tokens {toks[u].tolist()}  max |logit diff| {d:.3e}")
    ok &= bool(np.array_equal(toks[u], g[f"tokens{u}"])) and d < (1.0 if f16 else 2e-2)   # bit-exact ids at temperature 0; logits: CFG multiplies summation noise by 4 at std ~13 (f16: + rounding-boundary flips)
single, _ = par.generate_greedy([prompts[1]], steps)
ok &= bool(np.array_equal(single[0], toks[1])) and bool((ngen == steps).all())                                   # batching does not change a sequence
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("dtype", ["f32", pytest.param("f16", marks=UNRUN)])
def test_dia_greedy_tokens_and_logits_match_reference(dtype):
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, dtype], capture_output=True, text=True, timeout=150)
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0


QUANT_CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import dia_runner_from_file
from tts_cpp_b200.synth import cached_dia_gguf
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "dia_q8_0_vectors.npz"))
dia = dia_runner_from_file(cached_dia_gguf(seed=0, quant="Q8_0"))
toks, logits = dia.generate_teacher_forced([g["prompt0"], g["prompt1"]], np.stack([g["tokens0"], g["tokens1"]]))
ok = True
for u in range(2):
    rms = np.sqrt(((logits[u] - g[f"logits{u}"]) ** 2).mean(axis=(1, 2)))
    agree = float((toks[u] == g[f"tokens{u}"]).mean())
    print(f"PARITY dia Q8_0 prompt {u}: per-step logit rms {np.round(rms, 3).tolist()}, tokens equal {agree:.2f}")
    ok &= float(rms.max()) < 4.0 and agree >= 0.8
sys.exit(0 if ok else 1)
'''


@UNRUN
def test_dia_quantised_teacher_forced():
    """Q8_0 matrices (gemv_rows_q_kernel), teacher-forced on the reference's tokens; smoke-level bar (Dia amplifies re-quantisation noise: see tests/test_emu_cpu.py)."""
    r = subprocess.run([sys.executable, "-c", QUANT_CHILD, ROOT], capture_output=True, text=True, timeout=150)
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0
