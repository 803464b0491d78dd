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
