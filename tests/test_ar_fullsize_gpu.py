"""GPU: the Parler decode loop at BASELINE config 3's FULL model size (Parler-TTS-Mini-shaped F16 decoder: 24 layers x 1024, 16 heads, ffn 4096, nine
1088-wide heads; synthetic weights) through size-independent properties -- the reference cannot be run at this size inside a test, so nothing is compared with it:

  * the fused launches (grouped q/k/v GEMV writing k / v into the cache, GELU in fc1's epilogue) give the tokens of B2TTS_AR_FUSE=0 -- same per-output arithmetic;
  * CUDA-graph replay (B2TTS_AR_GRAPH=1) gives the tokens of the direct launches -- same kernels, same arguments, device-resident step counter;
  * a sequence's tokens do not depend on what else is in the batch (ragged prompts), and a shorter run is a prefix of a longer one (the cache ranges are causal).

The small-size tests (tests/test_parler_gpu.py) hold the comparison with the reference.  Like them: written without a GPU (logic checked under tests/emu), so
xfail(strict=False) in child processes until it has passed on a B200."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="AR decode paths not yet run on a B200 (round 1 GPU budget exhausted)")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from tts_cpp_b200.binding import parler_runner_from_file
from tts_cpp_b200.synth import PARLER_MINI_SHAPE, cached_parler_gguf
par = parler_runner_from_file(cached_parler_gguf(seed=0, f16=True, **PARLER_MINI_SHAPE))
rng = np.random.default_rng(11)
prompts = [rng.integers(1, 500, size=n).astype(np.uint32) for n in (24, 9, 17, 31)]
steps = int(sys.argv[3])
which = sys.argv[4]
toks = par.generate_greedy(prompts if which == "batch" else [prompts[2]], steps)
np.save(sys.argv[2], np.stack([np.asarray(t) for t in toks]))
print("decode ms", par.last_ms())
'''


def _run(tmp_path, tag, steps, which="batch", env=None):
    out = str(tmp_path / f"{tag}.npy")
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, out, str(steps), which], capture_output=True, text=True, timeout=420, env=dict(os.environ, **(env or {})))
    print(tag, r.stdout[-300:], r.stderr[-1500:])
    assert r.returncode == 0
    return np.load(out)


def test_parler_mini_size_fused_graph_batch_prefix_properties(tmp_path):
    steps = 40
    base = _run(tmp_path, "fused", steps)                                  # [4][steps][9]
    assert base.shape == (4, steps, 9) and base.min() >= 0 and base.max() < 1088
    assert len(np.unique(base)) > 50                                       # not a degenerate constant stream
    unfused = _run(tmp_path, "unfused", steps, env={"B2TTS_AR_FUSE": "0"})
    assert np.array_equal(base, unfused)
    graph = _run(tmp_path, "graph", steps, env={"B2TTS_AR_GRAPH": "1"})
    assert np.array_equal(base, graph)
    alone = _run(tmp_path, "alone", steps, which="single")
    assert np.array_equal(alone[0], base[2])                               # batching does not change a sequence
    short = _run(tmp_path, "short", 12)
    assert np.array_equal(short, base[:, :12])                             # a shorter run is a prefix of a longer one
