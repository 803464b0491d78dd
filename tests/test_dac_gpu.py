"""GPU: the DAC codec decoder (tts_cpp_b200/csrc/dac.cu) through the C-ABI against (1) the PCM the compiled UNMODIFIED reference
produced for the same codes (tests/golden/dac_vectors.npz, from oracle/_ref/dac_ref) and (2) the CPU restatement oracle/dac_port.py
on other lengths, including a ragged batch."""
import os

import numpy as np
import pytest

from conftest import report, rms

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dac_gguf():
    from tts_cpp_b200.synth import cached_dac_gguf
    return cached_dac_gguf(seed=0, max_frames=64)


@pytest.fixture(scope="module")
def dac(gpu_ctx, dac_gguf):
    from tts_cpp_b200.binding import dac_runner_from_file
    return dac_runner_from_file(dac_gguf, ctx=gpu_ctx)


def test_dac_matches_reference_pcm(dac):
    g = np.load(os.path.join(GOLD, "dac_vectors.npz"))
    outs = dac.run_batch([g["codes"][u] for u in range(g["codes"].shape[0])])
    for u, got in enumerate(outs):
        want = g["pcm"][u]
        d, r, mx = report(f"dac vs reference, utterance {u}", got, want)
        assert got.shape == want.shape and d < 1e-4 and mx < 1e-3          # north-star tolerance: PCM within 1e-4 RMS


def test_dac_ragged_batch_vs_port(dac, dac_gguf):
    from oracle.dac_port import DacPort
    from tts_cpp_b200.synth import synthetic_codes
    port = DacPort(dac_gguf)
    codes = [synthetic_codes(1, n, seed0=900 + n)[0] for n in (5, 31, 17)]
    outs = dac.run_batch(codes)
    for c, got in zip(codes, outs):
        want = port.decode(c)
        d, r, mx = report(f"dac vs port, {c.shape[0]} frames", got, want)
        assert got.shape == want.shape and d < 1e-4 and mx < 1e-3
    single = dac.run(codes[1])
    assert rms(single - outs[1]) < 1e-6                                     # batching changes tile shapes (summation order), nothing else


def test_dac_rejects_bad_codes(dac):
    from tts_cpp_b200.binding import B2TTSError
    bad = np.full((4, dac.n_heads), dac.codebook_size, np.uint32)
    with pytest.raises(B2TTSError):
        dac.run(bad)
