"""tests/golden/vad_cases.py -- the inputs of the VAD golden vectors, generated (not stored): every case is a list of (seconds, amplitude) segments of hash noise,
deterministic in integer arithmetic (independent of numpy's random generators), plus the arguments of apply_energy_voice_inactivity_detection
(reference examples/cli/vad.h:13-22).  Used by make_golden_vad.py (which runs the compiled reference on them) and by the tests (which run the port, the emulated
CUDA path and the GPU on them)."""
import numpy as np

DEFAULTS = dict(sample_rate=44100.0, ms_per_frame=10, frame_threshold=20, normalized_energy_threshold=0.01, trailing_silent_frames=5,
                early_cutoff_seconds_threshold=3, early_cutoff_energy_threshold=0.1)


def hash_noise(n, seed):
    """n floats in [-1, 1), full 24-bit mantissas (so that the squares are NOT exact in fp32: the rounding of the energy sum is exercised)"""
    i = np.arange(n, dtype=np.uint64) + np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        i ^= i >> np.uint64(33); i *= np.uint64(0xFF51AFD7ED558CCD); i ^= i >> np.uint64(33); i *= np.uint64(0xC4CEB9FE1A85EC53); i ^= i >> np.uint64(33)
    return ((i >> np.uint64(40)).astype(np.float64) / float(1 << 23) - 1.0).astype(np.float32)


def pcm_of(segments, sample_rate, seed, extra_samples=0):
    parts = []
    for k, (seconds, amp) in enumerate(segments):
        n = int(round(seconds * sample_rate))
        parts.append(hash_noise(n, seed * 131 + k) * np.float32(amp))
    if extra_samples:
        parts.append(hash_noise(extra_samples, seed * 131 + 99) * np.float32(0.3))
    return np.concatenate(parts) if parts else np.zeros(0, np.float32)


def cases():
    """-> list of (name, kwargs of the reference function, [pcm per utterance])"""
    out = []
    d = dict(DEFAULTS)
    # 1. default arguments, a batch of five: trailing silence trimmed; loud to the end (kept); a ragged tail; three seconds of digital silence in the middle
    #    (early cut-off); a 0.4 s utterance with a quiet tail shorter than frame_threshold
    out.append(("defaults", d, [
        pcm_of([(1.2, 0.5), (0.6, 0.001)], 44100.0, 1),
        pcm_of([(0.9, 0.4)], 44100.0, 2),
        pcm_of([(0.7, 0.6), (0.45, 0.0005)], 44100.0, 3, extra_samples=137),
        pcm_of([(0.5, 0.5), (3.2, 0.0), (0.5, 0.5)], 44100.0, 4),
        pcm_of([(0.3, 0.5), (0.1, 0.001)], 44100.0, 5),
    ]))
    # 2. degenerate lengths: empty, shorter than one frame, exactly one frame, a constant signal (max == min: the normalisation divides by zero)
    out.append(("degenerate", d, [
        np.zeros(0, np.float32),
        pcm_of([(0.005, 0.5)], 44100.0, 6),
        hash_noise(441, 7) * np.float32(0.5),
        np.full(441 * 30, 0.25, np.float32),
        np.zeros(441 * 25, np.float32),
    ]))
    # 3. Kokoro's rate, longer frames, and a trailing allowance LARGER than the silent run with a small frame_threshold: the reference subtracts a negative
    #    int from its size_t, i.e. n_outputs GROWS -- reproduced as is
    k = dict(sample_rate=24000.0, ms_per_frame=20, frame_threshold=2, normalized_energy_threshold=0.02, trailing_silent_frames=5, early_cutoff_seconds_threshold=1,
             early_cutoff_energy_threshold=0.05)
    out.append(("kokoro_rate_negative_trim", k, [
        pcm_of([(0.8, 0.5), (0.06, 0.0002)], 24000.0, 8),
        pcm_of([(0.5, 0.5), (1.3, 0.00001), (0.2, 0.5)], 24000.0, 9),
        pcm_of([(0.6, 0.3), (0.5, 0.0003)], 24000.0, 10, extra_samples=77),
    ]))
    # 4. a frame length that is a multiple of 4 (no fused tail in the reference build) and one with the longest fused tail (count & 3 == 3)
    m = dict(DEFAULTS, sample_rate=16000.0, ms_per_frame=10)          # 160 samples per frame
    out.append(("frame_160", m, [pcm_of([(0.5, 0.5), (0.4, 0.001)], 16000.0, 11)]))
    t = dict(DEFAULTS, sample_rate=44300.0, ms_per_frame=10)          # 443 samples per frame
    out.append(("frame_443", t, [pcm_of([(0.5, 0.5), (0.4, 0.001)], 44300.0, 12)]))
    # 5. BASELINE config 3's codec output size: 16 utterances x ~10 s at 44.1 kHz (7 M samples), ragged, each with its own quiet tail (0.3 .. 1.8 s); one is cut
    #    early by three seconds of digital silence in the middle
    big = []
    for b in range(16):
        tail = 0.3 + 0.1 * b
        if b == 5:
            big.append(pcm_of([(3.0, 0.4), (3.1, 0.0), (3.0, 0.4), (0.9, 0.0004)], 44100.0, 100 + b))
        else:
            big.append(pcm_of([(10.0 - tail - 0.003 * b, 0.2 + 0.02 * b), (tail, 0.0004)], 44100.0, 100 + b, extra_samples=17 * b))
    out.append(("config3_size", d, big))
    return out


def pack_input(kw, utts):
    """the input file of oracle/_ref/vad_ref and tests/emu vad_emu"""
    import struct
    head = struct.pack("<Ifiifiif", len(utts), kw["sample_rate"], kw["ms_per_frame"], kw["frame_threshold"], kw["normalized_energy_threshold"],
                       kw["trailing_silent_frames"], kw["early_cutoff_seconds_threshold"], kw["early_cutoff_energy_threshold"])
    n = np.asarray([u.size for u in utts], np.int64)
    return head + n.tobytes() + (np.concatenate(utts).astype(np.float32).tobytes() if utts else b"")


def unpack_output(raw, kw, utts):
    from oracle.vad_port import samples_per_frame
    spf = samples_per_frame(kw["sample_rate"], kw["ms_per_frame"])
    B = len(utts)
    n_out = np.frombuffer(raw[:8 * B], np.int64).copy()
    en = np.frombuffer(raw[8 * B:], np.float32)
    nf = [u.size // spf for u in utts]
    assert en.size == sum(nf), (en.size, sum(nf))
    cuts = np.concatenate([[0], np.cumsum(nf)]).astype(int)
    return n_out, [en[cuts[b]:cuts[b + 1]].copy() for b in range(B)]
