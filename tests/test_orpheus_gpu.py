"""GPU: the Orpheus decode loop (tts_cpp_b200/csrc/orpheus.cu) against the token ids and logits the compiled UNMODIFIED reference produced
(tests/golden/orpheus_vectors.npz: two prompts, 6 greedy steps each, small synthetic Orpheus GGUF).

This path was written after round 1's GPU budget was spent and has never run on a B200, hence xfail(strict=False): the test reports
XPASS / XFAIL without gating the suite, and it runs the check in a CHILD PROCESS so that a fault in the unvalidated kernels cannot poison
the CUDA context of the tests that follow.  Round 2 removes both once validated.

Update (end of round 1): the default greedy path ran on a B200 through scripts/rowb_first_contact.py and reproduced the reference's tokens (profiles/
r1i_rowb_first_contact.log); its test below is a plain test now, the variants that have not run yet keep xfail(strict=False) (UNRUN)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
# the default (plain fp32, fused launches) path has run on a B200 (profiles/r1i_rowb_first_contact.log: reference tokens, logits 5e-6); the variants below it have not
UNRUN = pytest.mark.xfail(strict=False, reason="this variant of the Orpheus decode path has not run on a B200 yet (round 1 GPU budget exhausted)")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import orpheus_runner_from_file
from tts_cpp_b200.synth import cached_orpheus_gguf
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "orpheus_vectors.npz"))
orph = orpheus_runner_from_file(cached_orpheus_gguf(seed=0))
prompts = [g["prompt0"], g["prompt1"]]
steps = g["tokens0"].size
toks, logits = orph.generate_greedy(prompts, steps, want_logits=True)         # one ragged batch of both prompts
ok = True
for u in range(2):
    d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
    print(f"PARITY orpheus prompt {u}: tokens {toks[u].tolist()} vs {g[f'tokens{u}'].tolist()}  max |logit diff| {d:.3e}")
    ok &= bool(np.array_equal(toks[u], g[f"tokens{u}"])) and d < 1e-3          # bit-exact token ids at temperature 0
single = orph.generate_greedy([prompts[1]], steps)
ok &= bool(np.array_equal(single[0], toks[1]))                                  # batching does not change a sequence
sys.exit(0 if ok else 1)
'''


def test_orpheus_greedy_tokens_and_logits_match_reference():
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], capture_output=True, text=True, timeout=150)
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0


WIDE_CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import orpheus_runner_from_file
from tts_cpp_b200.synth import cached_orpheus_gguf
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "orpheus_wide_vectors.npz"))
orph = orpheus_runner_from_file(cached_orpheus_gguf(seed=0, head_dim=128))
prompts = [g["prompt0"], g["prompt1"]]
toks, logits = orph.generate_greedy(prompts, g["tokens0"].size, want_logits=True)
ok = True
for u in range(2):
    d = float(np.abs(logits[u] - g[f"logits{u}"]).max())
    print(f"PARITY orpheus wide ({os.environ.get('B2TTS_AR_MMA', '0')}) prompt {u}: max |logit diff| {d:.3e}")
    ok &= bool(np.array_equal(toks[u], g[f"tokens{u}"])) and d < 1e-4
sys.exit(0 if ok else 1)
'''


@UNRUN
@pytest.mark.parametrize("mma", ["0", "1"], ids=["plain", "split_mma"])
def test_orpheus_wide_tokens_and_logits_match_reference(mma):
    """hidden 768 (every matrix eligible for the tensor-core GEMV); split_mma: the fp32-faithful three-product path over fp16 (hi, lo) pairs."""
    r = subprocess.run([sys.executable, "-c", WIDE_CHILD, ROOT], capture_output=True, text=True, timeout=150, env=dict(os.environ, B2TTS_AR_MMA=mma))
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0
