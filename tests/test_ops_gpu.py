"""GPU parity of the stand-alone kernels (through the C-ABI) against the oracle on identical inputs.

Byte/integer-like ops (uniform stream, cumsum, mod, round) must be bit-exact; floating-point ops are held to the
tolerance written at each assert (fp32 re-association / libm-vs-CUDA transcendental differences only).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import report

pytestmark = pytest.mark.gpu


def test_uniform_stream_bit_exact(gpu_ctx):
    from oracle.kokoro_port import minstd_uniform
    for skip, n in ((0, 6), (0, 100003), (12345678901, 4099), (9 * 600 * 198, 9 * 600 * 20)):
        got = gpu_ctx.uniform(n, skip)
        want = minstd_uniform(n, skip)
        assert np.array_equal(got, want), (skip, n, np.abs(got - want).max())
    # the reference's own first six draws (measured by linking the reference)
    kat = np.array([7.82590359e-06, 0.131537795, 0.75560534, 0.458650142, 0.532767236, 0.218959183], np.float32)
    assert np.array_equal(gpu_ctx.uniform(6, 0), kat)


def test_cumsum_mod_round_reciprocal_bit_exact(gpu_ctx):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((9, 400)).astype(np.float32) * 3
    want = np.zeros_like(x)
    run = np.zeros(9, np.float32)
    for t in range(400):
        run = (run + x[:, t]).astype(np.float32)
        want[:, t] = run
    assert np.array_equal(gpu_ctx.cumsum(x), want)
    assert np.array_equal(gpu_ctx.mod(x, 1.0), np.fmod(x, np.float32(1.0)))
    assert np.array_equal(gpu_ctx.round(x), np.trunc(x + np.float32(0.5)).astype(np.float32))
    assert np.array_equal(gpu_ctx.reciprocal(x), (np.float32(1.0) / x).astype(np.float32))
    # edge cases: empty-ish and single element
    assert np.array_equal(gpu_ctx.cumsum(np.ones((1, 1), np.float32)), np.ones((1, 1), np.float32))


def test_upscale_linear(gpu_ctx):
    from oracle.kokoro_port import upscale_linear
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((9, 57)).cumsum(axis=1) * 100).astype(np.float32)
    got = gpu_ctx.upscale_linear(x, 300)
    want = upscale_linear(x, 300)
    d, r, mx = report("upscale_linear", got, want)
    assert mx <= 1e-6 * float(np.abs(want).max()) + 1e-6   # only FMA-vs-separate rounding of the interpolation term


def test_snake(gpu_ctx):
    rng = np.random.default_rng(3)
    a = rng.uniform(0.3, 2.0, 64).astype(np.float32)
    x = rng.standard_normal((64, 333)).astype(np.float32) * 2
    want = x + np.sin(x * a[:, None]) ** 2 * (np.float32(1.0) / a[:, None])
    d, r, mx = report("snake", gpu_ctx.snake(a, x), want)
    assert mx < 2e-6


def test_stft_istft(gpu_ctx):
    from oracle.kokoro_port import stft_ref, istft_ref
    rng = np.random.default_rng(4)
    t = np.arange(3000)
    x = np.tanh(0.3 * np.sin(2 * np.pi * 120 * t / 24000) + 0.2 * np.sin(2 * np.pi * 960 * t / 24000) + 0.003 * rng.random(3000)).astype(np.float32)
    mag, ph = gpu_ctx.stft(x)
    wm, wp = stft_ref(x)
    assert mag.shape == wm.shape == (601, 11)
    d, r, mx = report("stft mag", mag, wm)
    assert mx < 2e-6
    # phase: compare where the magnitude is well above rounding noise; bins 0 and 10 must be exactly 0 or +pi like the reference
    # atan2 is discontinuous at +-pi: compare on the circle, and count the wraps (imag ~ 0, re < 0) separately
    ok = wm > 1e-4
    dp = np.abs(np.angle(np.exp(1j * (ph.astype(np.float64) - wp))))[ok]
    print("stft phase max circular diff (|X|>1e-4):", dp.max(), " +-pi wraps:", int((np.abs(ph - wp)[ok] > 6).sum()))
    assert dp.max() < 1e-3
    assert set(np.unique(ph[:, [0, 10]])) <= {np.float32(0.0), np.float32(np.pi)}
    # inverse on arbitrary (mag, sin-phase) like the generator feeds it
    m2 = np.exp(rng.standard_normal((601, 11)).astype(np.float32) * 0.5)
    p2 = np.sin(rng.standard_normal((601, 11)).astype(np.float32))
    got = gpu_ctx.istft(m2, p2)
    want = istft_ref(m2, p2)
    d, r, mx = report("istft", got, want)
    assert d < 2e-6 * max(r, 1.0) and mx < 2e-5


def test_conv_transpose_1d(gpu_ctx):
    rng = np.random.default_rng(5)
    for (K, cout, cin, L, s, p, op, g) in ((4, 2, 3, 5, 2, 1, 0, 1), (20, 16, 32, 9, 10, 5, 0, 1), (12, 8, 16, 7, 6, 3, 0, 1), (3, 6, 6, 5, 2, 1, 1, 6)):
        W = rng.standard_normal((cin, cout // g, K)).astype(np.float32)
        x = rng.standard_normal((cin, L)).astype(np.float32)
        want = F.conv_transpose1d(torch.from_numpy(x)[None], torch.from_numpy(W), None, stride=s, padding=p, output_padding=op, groups=g)[0].numpy()
        got = gpu_ctx.conv_transpose_1d(W, x, s, p, op, g)
        d, r, mx = report(f"convT K{K} s{s} g{g}", got, want)
        assert got.shape == want.shape and mx < 1e-4


def test_conv_1d_f16(gpu_ctx):
    rng = np.random.default_rng(6)
    for (K, cin, cout, L, s, p, dl) in ((3, 64, 96, 50, 1, 1, 1), (7, 128, 128, 300, 1, 9, 3), (11, 128, 128, 257, 1, 25, 5), (12, 22, 256, 601, 6, 3, 1), (5, 512, 512, 66, 1, 2, 1), (1, 514, 1024, 33, 1, 0, 1),
                                      # tcgen05 kernel corners: 128x256 tiles with dilation, 256x128 tiles over several row tiles, an unaligned Cout
                                      # (generic epilogue), a narrow head (Cout < 128, weight rows zero-filled by TMA) and a long pointwise layer
                                      (7, 256, 256, 700, 1, 9, 3), (3, 128, 128, 900, 1, 1, 1), (3, 64, 130, 300, 1, 1, 1), (7, 128, 22, 33000, 1, 3, 1),
                                      (1, 768, 2048, 400, 1, 0, 1)):
        W = (rng.standard_normal((cout, cin, K)) / np.sqrt(cin * K)).astype(np.float16).astype(np.float32)
        x = rng.standard_normal((cin, L)).astype(np.float32)
        xh = torch.from_numpy(x).half().float()
        want = F.conv1d(xh[None].double(), torch.from_numpy(W).double(), None, stride=s, padding=p, dilation=dl)[0].float().numpy()
        got = gpu_ctx.conv_1d(W, x, s, p, dl)
        d, r, mx = report(f"conv1d K{K} cin{cin} d{dl} s{s}", got, want)
        assert got.shape == want.shape and d < 2e-6 * max(r, 1.0) * np.sqrt(cin * K) / 8 and mx < 1e-4


def _torch_bilstm(w_ih, w_hh, b_ih, b_hh, x, lens):
    """the reference recurrence (model.cpp:53-86) with fp16 re-rounding of the activations (F16 weights)."""
    B, Lmax, In = x.shape
    H = w_hh.shape[-1]
    y = np.zeros((B, Lmax, 2 * H), np.float32)
    h16 = lambda t: t.half().float()
    for b in range(B):
        n = int(lens[b])
        xb = torch.from_numpy(x[b, :n])
        for d in range(2):
            Wi, Wh = torch.from_numpy(w_ih[d]), torch.from_numpy(w_hh[d])
            bi, bh = torch.from_numpy(b_ih[d]), torch.from_numpy(b_hh[d])
            pre = h16(xb) @ Wi.t() + bi
            h = torch.zeros(H); c = torch.zeros(H)
            for s in range(n):
                t = s if d == 0 else n - 1 - s
                g = pre[t] + (h16(h) @ Wh.t() + bh)
                i, f, gg, o = torch.sigmoid(g[:H]), torch.sigmoid(g[H:2 * H]), torch.tanh(g[2 * H:3 * H]), torch.sigmoid(g[3 * H:])
                c = f * c + i * gg
                h = torch.tanh(c) * o
                y[b, t, d * H:(d + 1) * H] = h.numpy()
    return y


@pytest.mark.parametrize("B,In,lens", [(1, 64, [7]), (3, 640, [66, 1, 40]), (33, 512, list(range(1, 34)))])
def test_bilstm_ragged(gpu_ctx, B, In, lens):
    rng = np.random.default_rng(7 + B)
    H = 256
    f16 = lambda a: a.astype(np.float16).astype(np.float32)
    w_ih = f16(rng.standard_normal((2, 4 * H, In)) / np.sqrt(In)); w_hh = f16(rng.standard_normal((2, 4 * H, H)) / np.sqrt(H))
    b_ih = f16(rng.standard_normal((2, 4 * H)) * 0.05); b_hh = f16(rng.standard_normal((2, 4 * H)) * 0.05)
    Lmax = max(lens)
    x = rng.standard_normal((B, Lmax, In)).astype(np.float32)
    got = gpu_ctx.bilstm(w_ih, w_hh, b_ih, b_hh, x, lens)
    want = _torch_bilstm(w_ih, w_hh, b_ih, b_hh, x, lens)
    for b in range(B):
        got[b, lens[b]:] = 0
    d, r, mx = report(f"bilstm B{B} In{In}", got, want)
    # fp16 re-rounding of h makes 1-ulp(fp16) flips possible: tolerance = a few fp16 ulps of O(1) values
    assert d < 2e-4 and mx < 5e-3
