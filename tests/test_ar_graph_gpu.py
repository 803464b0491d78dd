"""GPU: B2TTS_AR_GRAPH=1 -- one decode step of each autoregressive model captured into a CUDA graph and replayed (device-resident step counter); the token
ids must be the reference's (tests/golden/{orpheus,parler,dia}_vectors.npz).  Never run on a B200 yet (logic checked under tests/emu, stream capture
included): xfail(strict=False) in a child process, like the other tests of these paths."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="AR decode paths not yet run on a B200 (round 1 GPU budget exhausted)")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
model = sys.argv[2]
from tts_cpp_b200 import binding, synth
g = np.load(os.path.join(sys.argv[1], "tests", "golden", f"{model}_vectors.npz"))
runner = getattr(binding, f"{model}_runner_from_file")(getattr(synth, f"cached_{model}_gguf")(seed=0))
prompts = [g["prompt0"], g["prompt1"]]
out = runner.generate_greedy(prompts, g["tokens0"].shape[0])              # no logits requested -> the captured graph is replayed
toks = out[0] if isinstance(out, tuple) else out
ok = all(np.array_equal(np.asarray(toks[u]).reshape(g[f"tokens{u}"].shape), g[f"tokens{u}"]) for u in range(2))
print(f"PARITY {model} graph replay:", ok)
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("model", ["orpheus", "parler", "dia"])
def test_cuda_graph_replay_matches_reference_tokens(model):
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, model], capture_output=True, text=True, timeout=150, env=dict(os.environ, B2TTS_AR_GRAPH="1"))
    print(r.stdout[-2000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0
