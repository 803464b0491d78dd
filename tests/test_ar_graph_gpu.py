"""GPU: CUDA-graph replay of a decode step (the default of the launch-per-op paths since it passed on a B200) against direct launches (B2TTS_AR_GRAPH=0, child
process): both must give the reference's token ids (tests/golden/{orpheus,parler,dia}_vectors.npz, F32 models: these stay on the launch-per-op path)."""
import pytest

from conftest import run_snippet

pytestmark = pytest.mark.gpu

BODY = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
model = sys.argv[2]
from tts_cpp_b200 import binding, synth
g = np.load(os.path.join(sys.argv[1], "tests", "golden", f"{model}_vectors.npz"))
runner = getattr(binding, f"{model}_runner_from_file")(getattr(synth, f"cached_{model}_gguf")(seed=0))
prompts = [g["prompt0"], g["prompt1"]]
out = runner.generate_greedy(prompts, g["tokens0"].shape[0])              # no logits requested -> the captured graph is replayed (unless B2TTS_AR_GRAPH=0)
toks = out[0] if isinstance(out, tuple) else out
ok = all(np.array_equal(np.asarray(toks[u]).reshape(g[f"tokens{u}"].shape), g[f"tokens{u}"]) for u in range(2))
print(f"PARITY {model} B2TTS_AR_GRAPH={os.environ.get('B2TTS_AR_GRAPH')}:", ok)
runner.close()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("model", ["orpheus", "parler", "dia"])
@pytest.mark.parametrize("graph", [None, "0"], ids=["graph_replay", "direct_launches"])
def test_decode_step_graph_replay_and_direct_launches_match_reference_tokens(model, graph):
    assert run_snippet(BODY, [model], env=None if graph is None else {"B2TTS_AR_GRAPH": graph}) == 0
