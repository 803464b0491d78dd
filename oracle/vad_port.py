"""oracle/vad_port.py -- TEST INFRASTRUCTURE (the product never imports it): numpy restatement of the reference's trailing-silence trim,
apply_energy_voice_inactivity_detection + energy (reference examples/cli/vad.cpp:3-68).  Pinned to the compiled unmodified reference
(oracle/_ref/vad_ref) by tests/golden/vad_vectors.npz (tests/test_oracle_port.py): trimmed lengths and frame energies bit for bit.

The frame energy is a running fp32 sum in sample order (vad.cpp:5-7).  How the squares are rounded is a property of the reference BUILD: gcc -O2 -march=x86-64-v3
(oracle/Makefile) squares the first (count & ~3) samples with a vector multiply and adds them in order (two roundings), and contracts the last (count & 3) into
fused multiply-adds (one rounding); `fused_tail=True` restates that."""
import numpy as np


def samples_per_frame(sample_rate, ms_per_frame):
    return int(np.float32(ms_per_frame) * np.float32(sample_rate) / np.float32(1000.0))            # vad.cpp:20 (float arithmetic, truncated)


def frame_energies(pcm, spf, fused_tail=True):
    """energy() of every whole frame (vad.cpp:3-9, 33-34), vectorised ACROSS frames; the sum inside a frame stays sequential"""
    pcm = np.asarray(pcm, np.float32)
    nf = pcm.size // spf
    x = pcm[:nf * spf].reshape(nf, spf)
    en = np.zeros(nf, np.float32)
    split = spf & ~3 if fused_tail else spf
    for s in range(spf):
        if s < split:
            en = (en + x[:, s] * x[:, s]).astype(np.float32)                                       # float32 ops: product rounded, then the sum rounded
        else:
            en = (x[:, s].astype(np.float64) * x[:, s].astype(np.float64) + en.astype(np.float64)).astype(np.float32)   # fma: the product is exact in double; one rounding
    return en


def vad_trim(pcm, sample_rate=44100.0, ms_per_frame=10, frame_threshold=20, normalized_energy_threshold=0.01, trailing_silent_frames=5,
             early_cutoff_seconds_threshold=3, early_cutoff_energy_threshold=0.1, fused_tail=True):
    """-> (n_outputs, energies): what the reference leaves in data.n_outputs (its size_t arithmetic as Python ints, wrapped to int64)"""
    pcm = np.asarray(pcm, np.float32)
    spf = samples_per_frame(sample_rate, ms_per_frame)
    n = int(pcm.size)
    nf = n // spf
    early_frames = int((early_cutoff_seconds_threshold * 1000) / ms_per_frame)                     # vad.cpp:22 (int division of non-negative ints)
    e = frame_energies(pcm, spf, fused_tail)
    thr_e, thr_n = np.float32(early_cutoff_energy_threshold), np.float32(normalized_energy_threshold)
    mx = mn = np.float32(0)
    silent = 0
    for i in range(nf):                                                                            # vad.cpp:31-51
        v = e[i]
        if i == 0:
            mx = mn = v
        elif v > mx:
            mx = v
        elif v < mn:
            mn = v
        silent = silent + 1 if v <= thr_e else 0
        if silent >= early_frames:
            return _wrap((i + trailing_silent_frames - silent) * spf), e
    run = 0
    with np.errstate(all="ignore"):
        for i in range(nf, 0, -1):                                                                 # vad.cpp:55-62
            fe = np.float32(np.float32(e[i - 1] - mn) / np.float32(mx - mn))
            if fe < thr_n:
                run += 1
            else:
                break
    if run >= frame_threshold:
        n -= (run - trailing_silent_frames) * spf                                                  # size_t -= int: modulo 2^64 upstream
    return _wrap(n), e


def _wrap(v):
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v
