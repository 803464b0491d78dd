"""GPU: the trailing-silence trim on the device (vad.cu) through the C-ABI (b2tts_op_vad_trim) against the compiled unmodified reference
(examples/cli/vad.cpp; tests/golden/vad_vectors.npz made by tests/golden/make_golden_vad.py): integers, so bit-exact -- and so are the frame energies.
Collected last (conftest._LATE): new in this round's last session."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_vad_trim_matches_reference_bit_for_bit(gpu_ctx):
    import vad_cases
    g = np.load(os.path.join(ROOT, "tests", "golden", "vad_vectors.npz"))
    for name, kw, utts in vad_cases.cases():
        n_out, en = gpu_ctx.vad_trim(utts, **kw)
        print(f"PARITY vad {name}: n_in {[u.size for u in utts]} -> n_out {n_out.tolist()} (reference {g[name + '.n_out'].tolist()})")
        assert np.array_equal(n_out, g[name + ".n_out"]), name
        for b in range(len(utts)):
            assert np.array_equal(en[b], g[f"{name}.energies.{b}"]), (name, b)


def test_vad_trim_full_size_batch_property(gpu_ctx):
    """BASELINE config 3's codec output size (16 utterances x 10 s at 44.1 kHz) -- no reference run at this size in the goldens, so a size-independent property:
    the batch result equals the utterance-by-utterance result, appended digital silence beyond frame_threshold is cut back to the trailing allowance, and the kept
    length never exceeds the input's."""
    import vad_cases
    rng_len = [441000 - 137 * b for b in range(16)]
    utts = [np.concatenate([vad_cases.hash_noise(n - 44100, 50 + b) * np.float32(0.4), vad_cases.hash_noise(44100, 90 + b) * np.float32(0.0004)]) for b, n in enumerate(rng_len)]
    n_batch, _ = gpu_ctx.vad_trim(utts)
    for b, u in enumerate(utts):
        n_one, _ = gpu_ctx.vad_trim([u])
        assert int(n_one[0]) == int(n_batch[b])
        spf = 441
        voiced_frames = -(-(u.size - 44100) // spf)                      # the frame holding the last loud sample stays
        assert voiced_frames * spf + 5 * spf - spf <= int(n_batch[b]) <= voiced_frames * spf + 5 * spf + spf, (b, int(n_batch[b]), voiced_frames)
        assert int(n_batch[b]) <= u.size


def test_vad_trim_rejects_what_the_reference_divides_by_zero_on(gpu_ctx):
    with pytest.raises(RuntimeError):
        gpu_ctx.vad_trim([np.zeros(100, np.float32)], ms_per_frame=0)
    with pytest.raises(RuntimeError):
        gpu_ctx.vad_trim([np.zeros(100, np.float32)], sample_rate=50.0, ms_per_frame=10)
