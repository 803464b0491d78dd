"""GPU parity of the Kokoro forward (through the C-ABI) against the oracle port, stage by stage.

Why stage-wise (DESIGN.md "parity floor"): two builds of the reference itself differ by 0.07 RMS in PCM on this model, because
every F16 matmul re-rounds activations to fp16 and the harmonic source integrates f0 into a phase.  So each CUDA stage is
checked on IDENTICAL inputs: all stage inputs are overridden with the oracle's tensors (teacher forcing) in one run, and
every stage output is compared with the oracle's output for the same inputs.  Integers (durations) must be bit-exact.
"""
import numpy as np
import pytest

from conftest import report, rms

pytestmark = pytest.mark.gpu


def _utts(ns, seed0=500):
    from tts_cpp_b200.synth import synthetic_prompts
    return [synthetic_prompts(1, n_phonemes=n - 2, seed0=seed0 + i)[0] for i, n in enumerate(ns)]


def _port_run(port, toks, skip=0):
    taps = {}
    lens, pcm = port.run(toks, skip, taps)
    return lens, pcm, taps


def _cl(t):  # oracle [C, L] -> channels-last [L, C]
    return np.ascontiguousarray(t.numpy().T)


def test_free_running_durations_and_shapes(runner, port):
    """No teacher forcing: durations bit-exact, sample counts 600*sum(durations), PCM finite and the right scale."""
    utts = _utts([18, 9, 31])
    pcms, durs = runner.run_batch(utts)
    for u, p, d in zip(utts, pcms, durs):
        lens, ppcm, _ = _port_run(port, u)
        assert np.array_equal(d, lens), (d, lens)
        assert p.shape[0] == 600 * int(lens.sum())
        assert np.isfinite(p).all()
        r = rms(p) / rms(ppcm)
        print("free-running pcm rms ratio", r, "diff rms", rms(p - ppcm))
        assert 0.7 < r < 1.4


def test_stagewise_teacher_forced(runner, port):
    utts = _utts([18, 11])
    runner.set_taps(True)
    ref = [_port_run(port, u, skip) for u, skip in zip(utts, (0, 777))]
    B = len(utts)
    nmax = max(len(u) for u in utts)
    T = [int(r[0].sum()) for r in ref]
    tmax = max(T)

    def pad(arrs, L, C):
        out = np.zeros((B, L, C), np.float32)
        for b, a in enumerate(arrs):
            out[b, :a.shape[0]] = a.reshape(a.shape[0], C)
        return out

    ov = {
        "d": pad([r[2]["d"].numpy() for r in ref], nmax, 640),
        "lens": pad([r[0][:, None] for r in ref], nmax, 1),
        "shared": pad([r[2]["shared"].numpy() for r in ref], tmax, 512),
        "f0": pad([r[2]["f0"].numpy()[:, None] for r in ref], 2 * tmax, 1),
        "n": pad([r[2]["n"].numpy()[:, None] for r in ref], 2 * tmax, 1),
        "t_en": pad([r[2]["t_en"].numpy() for r in ref], nmax, 512),
        "dec": pad([_cl(r[2]["dec"]) for r in ref], 2 * tmax, 512),
        "har_spec": pad([np.concatenate([r[2]["mag"].numpy(), r[2]["ph"].numpy()], axis=1) for r in ref], 120 * tmax + 1, 22),
        # generator stage buffers use the polyphase ConvTranspose pitch: (2*Tmax + 1) * 10 rows per utterance
        "gen_out0": pad([_cl(r[2]["gen_out0"]) for r in ref], (2 * tmax + 1) * 10, 256),
    }
    har = np.zeros((B, 600 * tmax), np.float32)
    for b, r in enumerate(ref):
        har[b, :600 * T[b]] = r[2]["har"].numpy()
    ov["har"] = har
    try:
        for k, v in ov.items():
            runner.override(k, v)
        pcms, durs = runner.run_batch(utts, noise_skip=[0, 777])
    finally:
        for k in ov:
            runner.override(k, None)

    def got(name, b, L):
        return runner.tap(name)[b, :L]

    worst = {}
    for b, (lens, ppcm, tp) in enumerate(ref):
        n = len(utts[b])
        checks = [
            ("albert", got("albert", b, n), tp["albert"].numpy(), 5e-3),          # free-running 12-layer stack with fp16 re-rounding
            ("shared", got("shared", b, T[b]), tp["shared"].numpy(), 2e-3),        # LSTM over T steps from identical d
            ("f0", got("f0", b, 2 * T[b])[:, 0], tp["f0"].numpy(), 2e-5),          # relative (f0 ~ 120)
            ("n", got("n", b, 2 * T[b])[:, 0], tp["n"].numpy(), 5e-3),
            ("t_en", got("t_en", b, n), tp["t_en"].numpy(), 2e-3),
            ("dec", got("dec", b, 2 * T[b]), _cl(tp["dec"]), 3e-3),
            ("har", runner.tap("har")[b, :600 * T[b]], tp["har"].numpy(), 2e-5),
            ("gen_out0", got("gen_out0", b, 20 * T[b]), _cl(tp["gen_out0"]), 3e-3),
            ("gen_out1", got("gen_out1", b, 120 * T[b] + 1), _cl(tp["gen_out1"]), 3e-3),
            ("pcm", pcms[b], ppcm, None),
        ]
        for name, g, w, tol in checks:
            d, r, mx = report(f"u{b} {name}", g, w)
            worst[name] = max(worst.get(name, 0.0), d / max(r, 1e-30))
            if tol is not None:
                assert d <= tol * max(r, 1e-30), (name, d, r)
        # durations: bit-exact given identical d
        assert np.array_equal(durs[b], lens)
        # PCM given identical generator inputs up to stage 0: the north-star tolerance
        dp = rms(pcms[b] - ppcm)
        print(f"u{b} teacher-forced PCM diff rms = {dp:.3g} vs the PORT (itself 8.6e-5 from the reference; CUDA vs the reference directly: 1e-4 asserted in test_golden_gpu / test_bench_size_gpu)")
        assert dp < 2e-4
    print("WORST", worst)


def test_stft_stage_phase_convention(runner, port):
    """har -> (mag, phase): bins 0 / 10 are exactly 0 or +pi, other phases match where the bin is not noise."""
    utts = _utts([12])
    runner.set_taps(True)
    lens, ppcm, tp = _port_run(port, utts[0])
    T = int(lens.sum())
    har = tp["har"].numpy()[None]
    try:
        runner.override("har", np.ascontiguousarray(har))
        runner.override("lens", lens[None, :, None])
        runner.run_batch(utts)
    finally:
        runner.override("har", None); runner.override("lens", None)
    hs = runner.tap("har_spec")[0, :120 * T + 1]
    mag, ph = hs[:, :11], hs[:, 11:]
    wm, wp = tp["mag"].numpy(), tp["ph"].numpy()
    d, r, mx = report("stage stft mag", mag, wm)
    assert mx < 5e-6
    ok = wm > 1e-3
    circ = np.abs(np.angle(np.exp(1j * (ph.astype(np.float64) - wp))))[ok]
    print("stage stft phase: max circular diff", circ.max(), " +-pi wraps:", int((np.abs(ph - wp)[ok] > 6).sum()), "of", int(ok.sum()))
    assert circ.max() < 2e-3
    assert set(np.unique(ph[:, [0, 10]])) <= {np.float32(0.0), np.float32(np.pi)}


def test_ragged_batch_matches_single(runner):
    """An utterance computes the same thing alone and inside a ragged batch (padding / masking is inert)."""
    utts = _utts([18, 5, 27, 3])
    pb, db = runner.run_batch(utts)
    for i, u in enumerate(utts):
        p1, d1 = runner.run(u)
        assert np.array_equal(d1, db[i])
        assert p1.shape == pb[i].shape
        d = rms(p1 - pb[i])
        print(f"ragged u{i} n={len(u)} single-vs-batch pcm diff rms {d:.3g}")
        assert d < 5e-3   # same kernels, same summation order per utterance except double atomics in the norm statistics


def test_noise_skip_changes_only_noise(runner):
    u = _utts([10])[0]
    a, _ = runner.run(u, noise_skip=0)
    b, _ = runner.run(u, noise_skip=0)
    c, _ = runner.run(u, noise_skip=12345)
    assert rms(a - b) < 5e-3
    assert rms(a - c) > 1e-4


def test_errors_are_loud(runner):
    from tts_cpp_b200.binding import B2TTSError
    with pytest.raises(B2TTSError):
        runner.run_batch([[0, 5]])            # < 3 tokens: voice row n-3 does not exist
    with pytest.raises(B2TTSError):
        runner.run([0, 5, 6, 0], voice="no_such_voice")
