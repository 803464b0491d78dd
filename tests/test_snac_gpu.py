"""GPU: the SNAC codec decoder (tts_cpp_b200/csrc/dac.cu, struct Snac) against the PCM of the compiled UNMODIFIED reference
(tests/golden/snac_vectors.npz: two utterances decoded in one process, so the second continues the reference's noise stream).

Measured on a B200 at the end of round 1: 3.6e-6 / 3.7e-6 RMS against the reference's PCM (signal RMS 0.12)."""
import os

import numpy as np
import pytest

from conftest import report

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _split(c):
    c = np.asarray(c, np.uint32)
    L = c.size * 4 // 7
    return [c[:L // 4], c[L // 4:L // 4 + L // 2], c[L // 4 + L // 2:]]


def test_snac_matches_reference_pcm(gpu_ctx):
    from tts_cpp_b200.binding import snac_runner_from_file
    from tts_cpp_b200.synth import cached_snac_gguf
    g = np.load(os.path.join(GOLD, "snac_vectors.npz"))
    snac = snac_runner_from_file(cached_snac_gguf(seed=0, max_frames=64), ctx=gpu_ctx)
    outs = snac.run_batch([_split(g["codes"][u]) for u in range(g["codes"].shape[0])])     # one batch == the reference's two sequential runs
    for u, got in enumerate(outs):
        d, r, mx = report(f"snac vs reference, utterance {u}", got, g["pcm"][u])
        assert got.shape == g["pcm"][u].shape and d < 1e-4 and mx < 1e-3
    snac.reset_noise()
    again = snac.run_batch([_split(g["codes"][0])])[0]
    assert np.abs(again - outs[0]).max() < 1e-6                                                # reset_noise rewinds the stream
