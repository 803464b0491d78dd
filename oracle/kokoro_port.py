"""oracle/kokoro_port.py -- TEST INFRASTRUCTURE.  CPU restatement of the reference Kokoro forward.

This file is the *checker*, never the product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import it.  It restates, in numpy + CPU torch (fp32), what the
reference's two GGML graphs compute for one utterance, stage by stage, so every CUDA stage can
be compared against it on identical inputs on the GPU box (where /root/reference does not exist).

Pinned against the compiled reference (oracle/_ref/kokoro_ref, built from the unmodified
sources by oracle/Makefile): tests/golden/make_golden.py runs both on the same synthetic GGUF
and tests/test_oracle_port.py checks the committed vectors.  NOTE (DESIGN.md "parity floor"):
two builds of the *reference itself* (x86-64-v3 vs x86-64-v2) differ by 4e-3 in the duration
hidden states and 0.07 RMS in PCM on this model, because every F16-weight matmul re-rounds its
activations to fp16 (ggml-cpu.c:262-267) and the harmonic source integrates f0 into a phase --
so tolerances against the reference are set from that measured floor, and tight tolerances are
used only between this port and the CUDA path on identical stage inputs.

Reference files followed (file:line of /root/reference):
  src/models/kokoro/model.cpp:10-31     ALBERT embeddings / LayerNorm
  src/models/kokoro/model.cpp:35-86     bi-LSTM (gate order i,f,g,o; even = input side, odd = hidden side)
  src/models/kokoro/model.cpp:88-134    AdaIN residual conv block
  src/models/kokoro/model.cpp:136-171   generator residual block / noise block
  src/models/kokoro/model.cpp:173-244   sine source, generator, STFT/iSTFT plumbing
  src/models/kokoro/model.cpp:938-1047  duration graph
  src/models/kokoro/model.cpp:1141-1275 generation graph and its host-side inputs
  src/util.cpp:66-72,86-137,140-172,203-217   RNG, snake, stft/istft wrappers, uv/noise op, window sums
  ggml/src/ggml-cpu/ggml-cpu.c:262-267,1797-1830,5570-5600,7114-7163,8476-8760,10104-10200,10912-10964
"""
from __future__ import annotations

import math
import numpy as np
import torch
import torch.nn.functional as F

F32 = np.float32


# --------------------------------------------------------------------------------------
# the reference's process-wide uniform generator (src/util.cpp:66-72): libstdc++
# std::default_random_engine == minstd_rand0 (a=16807, m=2^31-1, seed 1) driving
# uniform_real_distribution<float>(0,1) == generate_canonical<float,24>: one draw,
# (float)(x-1) / 2147483648.0f, clamped below 1.
# --------------------------------------------------------------------------------------
_M = (1 << 31) - 1
_A = 16807


def minstd_uniform(count: int, skip: int = 0) -> np.ndarray:
    """Draws skip+1 .. skip+count of the reference's static uniform engine."""
    x0 = pow(_A, skip, _M)  # state after `skip` draws from seed 1
    blk = 1 << 16
    tab = np.empty(blk, dtype=np.uint64)  # tab[i] = A^(i+1) mod M
    tab[0] = _A
    k = 1
    while k < blk:
        ak = int(tab[k - 1])
        n = min(k, blk - k)
        tab[k:k + n] = (tab[:n] * np.uint64(ak)) % np.uint64(_M)
        k += n
    out = np.empty(count, dtype=np.uint64)
    pos = 0
    while pos < count:
        n = min(blk, count - pos)
        out[pos:pos + n] = (tab[:n] * np.uint64(x0)) % np.uint64(_M)
        x0 = int(out[pos + n - 1])
        pos += n
    u = (out - np.uint64(1)).astype(np.float32) / np.float32(2147483648.0)
    return np.minimum(u, np.nextafter(np.float32(1.0), np.float32(0.0)))


def hann20(n_fft: int = 20) -> np.ndarray:
    """src/util.cpp:132-137: (float) pow(sin(pi*i/n), 2) evaluated in double."""
    return np.array([math.pow(math.sin(math.pi * i / n_fft), 2.0) for i in range(n_fft)], dtype=np.float32)


def window_sq_sum(n_fft: int, hop: int, n_frames: int, w: np.ndarray) -> np.ndarray:
    """src/util.cpp:203-217 (note: iterates n_frames + half/hop frames -- one more than the iSTFT has)."""
    cutoff = n_frames * hop
    half = n_fft // 2
    tgt = np.zeros(cutoff, dtype=np.float32)
    w64 = w.astype(np.float32).astype(np.float64)
    for i in range(n_frames + half // hop):
        lo = i * hop - half
        a, b = max(lo, 0), min(lo + n_fft, cutoff)
        if a < b:
            # `tgt[index] += powf(window[ii], 2)` compiles to one FMA in the reference build (gcc -O3, -ffp-contract=fast, FMA ISA)
            tgt[a:b] = (w64[a - lo:b - lo] * w64[a - lo:b - lo] + tgt[a:b].astype(np.float64)).astype(np.float32)
    return tgt


# ------------------------------------------------------------------ ggml numerics helpers
def _h(x: torch.Tensor) -> torch.Tensor:
    """Activation re-rounding ggml applies before an F16-weight matmul / im2col (ggml-cpu.c:262-267, ggml.c:3878-3882)."""
    return x.to(torch.float16).to(torch.float32)


def ggml_norm(x: torch.Tensor, eps: float, dim: int = -1) -> torch.Tensor:
    """ggml-cpu.c:7114-7163: double-accumulated mean / variance, fp32 elsewhere."""
    mean = x.double().mean(dim=dim, keepdim=True).float()
    v = x - mean
    var = (v * v).double().mean(dim=dim, keepdim=True).float()
    scale = 1.0 / torch.sqrt(var + eps)
    return v * scale


def gelu_f16_lut(x: torch.Tensor) -> torch.Tensor:
    """ggml-cpu.c:1816-1830: y = fp16(gelu_tanh(fp16(x))) for |x| < 10."""
    xh = _h(x)
    y = 0.5 * xh * (1.0 + torch.tanh(0.79788456080286535587989211986876 * xh * (1.0 + 0.044715 * xh * xh)))
    y = _h(y)
    return torch.where(x <= -10.0, torch.zeros_like(x), torch.where(x >= 10.0, x, y))


def leaky(x: torch.Tensor, ns: float) -> torch.Tensor:
    return torch.where(x > 0, x, torch.zeros_like(x)) + np.float32(ns) * torch.where(x < 0, x, torch.zeros_like(x))


def ggml_round(x: torch.Tensor) -> torch.Tensor:
    """ggml-cpu.c:1797: (float)(int)(x + 0.5f)."""
    return torch.trunc(x + 0.5)


def upscale_linear(x: np.ndarray, factor: int) -> np.ndarray:
    """ggml-cpu.c:10912-10964 along the last axis (bespoke, edge clamped)."""
    n = x.shape[-1]
    ne0 = n * factor
    sf0 = np.float32(ne0) / np.float32(n)
    hsf0 = sf0 / np.float32(2.0)
    sf, hsf = int(sf0), int(hsf0)
    i0 = np.arange(ne0)
    i00 = ((i0.astype(np.float32) - hsf0) / sf0).astype(np.int64)
    i00 = np.clip(i00, 0, n - 2)
    base = x[..., i00]
    top = x[..., i00 + 1]
    diff_adj = (top - base) / sf0
    adj = ((i0 - hsf) % sf).astype(np.float32) * diff_adj + diff_adj / np.float32(2.0)
    y = (base + adj).astype(np.float32)
    y[..., :hsf] = x[..., :1]
    y[..., ne0 - hsf:] = x[..., -1:]
    return y


def stft_ref(x: np.ndarray, n_fft: int = 20, hop: int = 5) -> tuple[np.ndarray, np.ndarray]:
    """ggml-cpu.c:8560-8640: centre (reflect) framing, Hann, DFT, |X| and atan2.  Returns one-sided (mag, phase) [frames, bins].
    The reference's radix-2/DFT leaves imag == +0.0 exactly for bins 0 and n_fft/2 of a real frame, so their phase is 0 or +pi."""
    w = hann20(n_fft).astype(np.float64)
    half = n_fft // 2
    L = x.shape[0]
    frames = L // hop + 1
    idx = (np.arange(frames) * hop - half)[:, None] + np.arange(n_fft)[None, :]
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= L, L - (idx - L + 1) - 0, idx)
    idx = np.where(idx >= L, 2 * L - idx - 1, idx)
    fr = x.astype(np.float64)[idx] * w[None, :]
    k = np.arange(n_fft // 2 + 1)
    ang = -2.0 * np.pi * np.outer(np.arange(n_fft), k) / n_fft
    re = fr @ np.cos(ang)
    im = fr @ np.sin(ang)
    im[:, 0] = 0.0
    im[:, -1] = 0.0
    mag = np.sqrt(re * re + im * im).astype(np.float32)
    ph = np.arctan2(im.astype(np.float32), re.astype(np.float32)).astype(np.float32)
    return mag, ph


def istft_ref(mag: np.ndarray, ph: np.ndarray, n_fft: int = 20, hop: int = 5) -> np.ndarray:
    """ggml-cpu.c:8665-8760 + src/util.cpp:123-130: one-sided (mag, phase) [frames, bins] -> signal / window-square-sum."""
    frames = mag.shape[0]
    half = n_fft // 2
    n_out = (frames - 1) * hop
    re = (mag.astype(np.float32) * np.cos(ph.astype(np.float32)).astype(np.float32)).astype(np.float64)
    im = (mag.astype(np.float32) * np.sin(ph.astype(np.float32)).astype(np.float32)).astype(np.float64)
    full = np.zeros((frames, n_fft), dtype=np.complex128)
    full[:, :half + 1] = re + 1j * im
    full[:, half + 1:] = (re - 1j * im)[:, 1:half][:, ::-1]
    # the reference keeps only Re(IDFT) (it reads mdst of the reversed forward FFT)
    t = np.fft.ifft(full, axis=1).real
    w = hann20(n_fft).astype(np.float64)
    out = np.zeros(n_out + 2 * n_fft, dtype=np.float64)
    for f in range(frames):
        out[f * hop:f * hop + n_fft] += t[f] * w
    sig = out[half:half + n_out].astype(np.float32)
    return sig / window_sq_sum(n_fft, hop, n_out // hop, hann20(n_fft))


# --------------------------------------------------------------------------------------
class KokoroPort:
    def __init__(self, gguf_path: str, threads: int = 8):
        import gguf
        torch.set_num_threads(threads)
        rd = gguf.GGUFReader(gguf_path)
        self.w: dict[str, torch.Tensor] = {}
        self.f16: dict[str, bool] = {}
        for t in rd.tensors:
            name = t.name[len("kokoro."):] if t.name.startswith("kokoro.") else t.name
            arr = np.array(t.data)
            self.f16[name] = arr.dtype == np.float16
            self.w[name] = torch.from_numpy(arr.astype(np.float32))
        self.kv = {}
        for k, f in rd.fields.items():
            if k.startswith("kokoro.") and len(f.data) == 1 and f.types and f.types[0].name in ("UINT32",):
                self.kv[k] = int(f.parts[f.data[0]][0])
        self.voice = self.w["voice_tensors.af_heart"]  # [510, 256]

    # -- primitive layers -------------------------------------------------------------
    def lin(self, name: str, x: torch.Tensor, bias: str | None = None) -> torch.Tensor:
        """y[..., out] = x[..., in] @ W[out,in]^T (+ b); activations re-rounded to fp16 when W is stored F16."""
        W = self.w[name]
        W = W.reshape(W.shape[0], -1)
        xi = _h(x) if self.f16[name] else x
        y = xi @ W.t()
        if bias is not None:
            y = y + self.w[bias]
        return y

    def conv(self, name: str, x: torch.Tensor, bias: str | None, stride=1, pad=0, dil=1) -> torch.Tensor:
        """x [C, L] -> [Cout, Lout]  (ggml_conv_1d = im2col + mul_mat, ggml.c:3870-3894)."""
        W = self.w[name]
        xi = _h(x) if self.f16[name] else x
        y = F.conv1d(xi[None], W, None, stride=stride, padding=pad, dilation=dil)[0]
        if bias is not None:
            y = y + self.w[bias][:, None]
        return y

    def lstm(self, base: str, x: torch.Tensor) -> torch.Tensor:
        """x [len, in] -> [len, 2*hid]  (model.cpp:35-86)."""
        outs = []
        for part, bpart, rev in (("weights", "biases", False), ("reverse_weights", "reverse_biases", True)):
            pre = [self.lin(f"{base}.0.{part}.{2 * g}", x, f"{base}.0.{bpart}.{2 * g}") for g in range(4)]
            hid = pre[0].shape[1]
            h = torch.zeros(hid)
            c = torch.zeros(hid)
            res = [None] * x.shape[0]
            order = range(x.shape[0] - 1, -1, -1) if rev else range(x.shape[0])
            for t in order:
                g4 = [pre[g][t] + (self.lin(f"{base}.0.{part}.{2 * g + 1}", h) + self.w[f"{base}.0.{bpart}.{2 * g + 1}"]) for g in range(4)]
                i_g = torch.sigmoid(g4[0]); f_g = torch.sigmoid(g4[1]); g_g = torch.tanh(g4[2]); o_g = torch.sigmoid(g4[3])
                c = f_g * c + i_g * g_g
                h = torch.tanh(c) * o_g
                res[t] = h
            outs.append(torch.stack(res, 0))
        return torch.cat(outs, dim=1)

    def adain(self, x: torch.Tensor, gw: str, gb: str, bw: str, bb: str, style: torch.Tensor) -> torch.Tensor:
        """x [C, L]: InstanceNorm over time (eps 1e-5) then x + x*gamma + beta (model.cpp:93-100)."""
        gamma = self.lin(gw, style, gb)
        beta = self.lin(bw, style, bb)
        n = ggml_norm(x, 1e-5, dim=1)
        return (n + n * gamma[:, None]) + beta[:, None]

    def ada_block(self, base: str, x: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
        """AdaIN residual conv block, x [Cin, L] -> [Cout, L or 2L]  (model.cpp:88-134)."""
        w = self.w
        cur = self.adain(x, f"{base}.norm1_gamma_weight", f"{base}.norm1_gamma_bias", f"{base}.norm1_beta_weight", f"{base}.norm1_beta_bias", style)
        cur = leaky(cur, 0.2)
        has_pool = f"{base}.pool_weight" in w
        if has_pool:
            C = cur.shape[0]
            cur = F.conv_transpose1d(cur[None], w[f"{base}.pool_weight"], None, stride=2, padding=1, output_padding=1, groups=C)[0]
            cur = cur + w[f"{base}.pool_bias"][:, None]
        cur = self.conv(f"{base}.conv1_weight", cur, f"{base}.conv1_bias", pad=1)
        cur = self.adain(cur, f"{base}.norm2_gamma_weight", f"{base}.norm2_gamma_bias", f"{base}.norm2_beta_weight", f"{base}.norm2_beta_bias", style)
        cur = leaky(cur, 0.2)
        res = self.conv(f"{base}.conv2_weight", cur, f"{base}.conv2_bias", pad=1)
        sc = x
        if f"{base}.conv1x1_weight" in w:
            if has_pool:
                sc = sc.repeat_interleave(2, dim=1)            # nearest x2 (ggml_upscale_ext)
            sc = self.lin(f"{base}.conv1x1_weight", sc.t()).t()  # bias is loaded but never applied (model.cpp:129)
        return (res + sc) / np.float32(math.sqrt(2.0))

    def snake(self, alpha: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """x + sin^2(alpha x) * (1/alpha), x [C, L] (src/util.cpp:86-101)."""
        a = alpha.reshape(-1, 1)
        s = torch.sin(x * a)
        return x + (s * s) * (1.0 / a)

    def gen_resblock(self, base: str, x: torch.Tensor, style: torch.Tensor, pads, dils) -> torch.Tensor:
        """model.cpp:136-165, x [C, L]."""
        inp = x
        for i in range(3):
            cur = self.adain(inp, f"{base}.{i}.gamma1_weight", f"{base}.{i}.gamma1_bias", f"{base}.{i}.beta1_weight", f"{base}.{i}.beta1_bias", style)
            cur = self.snake(self.w[f"{base}.{i}.alpha1"], cur)
            cur = self.conv(f"{base}.{i}.convs1_weight", cur, f"{base}.{i}.convs1_bias", pad=pads[i], dil=dils[i])
            cur = self.adain(cur, f"{base}.{i}.gamma2_weight", f"{base}.{i}.gamma2_bias", f"{base}.{i}.beta2_weight", f"{base}.{i}.beta2_bias", style)
            cur = self.snake(self.w[f"{base}.{i}.alpha2"], cur)
            cur = self.conv(f"{base}.{i}.convs2_weight", cur, f"{base}.{i}.convs2_bias", pad=pads[0], dil=1)
            inp = inp + cur
        return inp

    # -- stages -------------------------------------------------------------------------
    def albert(self, tokens) -> torch.Tensor:
        """tokens [n] -> [n, 768]  (model.cpp:10-31, 967-1008)."""
        w = self.w
        tok = torch.as_tensor(np.asarray(tokens, dtype=np.int64))
        n = tok.shape[0]
        x = (w["albert.token_embd"][tok] + w["albert.position_embd"][:n]) + w["albert.token_type_embd"]
        x = ggml_norm(x, 1e-12) * w["albert.norm"] + w["albert.norm_bias"]
        x = self.lin("albert.embd", x, "albert.embd_bias")
        L = "albert.layer.0."
        rec = self.kv.get("kokoro.duration_predictor.albert.recurrence", 12)
        heads = self.kv.get("kokoro.duration_predictor.albert.attn_heads", 12)
        hd = x.shape[1] // heads
        for _ in range(rec):
            q = self.lin(L + "q", x, L + "q_bias").reshape(n, heads, hd).permute(1, 0, 2)
            k = self.lin(L + "k", x, L + "k_bias").reshape(n, heads, hd).permute(1, 0, 2)
            v = self.lin(L + "v", x, L + "v_bias").reshape(n, heads, hd).permute(1, 0, 2)
            kq = (q @ k.transpose(1, 2)) * np.float32(0.125)
            kq = kq - kq.max(dim=-1, keepdim=True).values
            p = torch.exp(kq)
            p = p * (1.0 / p.double().sum(dim=-1, keepdim=True)).float()
            att = (p @ v).permute(1, 0, 2).reshape(n, heads * hd)
            att = self.lin(L + "o", att, L + "o_bias")
            x = att + x
            x = ggml_norm(x, 1e-12) * w[L + "ffn_norm"] + w[L + "ffn_norm_bias"]       # names are crossed in the reference (model.cpp:765-770)
            f = gelu_f16_lut(self.lin(L + "ffn", x, L + "ffn_bias"))
            f = self.lin(L + "ffn_out", f, L + "ffn_out_bias")
            x = f + x
            x = ggml_norm(x, 1e-12) * w[L + "attn_norm"] + w[L + "attn_norm_bias"]
        return x

    def styles(self, n_tokens: int):
        row = self.voice[n_tokens - 3]
        return row[128:256].clone(), row[0:128].clone()   # (prosody style, decoder style)  model.cpp:1013,1213

    def prosody(self, albert_out: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
        """-> d [n, 640]  (model.cpp:1011-1033)."""
        n = albert_out.shape[0]
        cur = self.lin("duration_predictor.encode", albert_out, "duration_predictor.encode_bias")
        cur = torch.cat([cur, style[None].expand(n, -1)], dim=1)
        for i in range(3):
            cur = self.lstm(f"duration_predictor.layers.{2 * i}.lstm", cur)
            p = f"duration_predictor.layers.{2 * i + 1}."
            gamma = self.lin(p + "gamma_weight", style, p + "gamma_bias")
            beta = self.lin(p + "beta_weight", style, p + "beta_bias")
            nn = ggml_norm(cur, 1e-5)
            cur = (nn + nn * gamma) + beta
            cur = torch.cat([cur, style[None].expand(n, -1)], dim=1)
        return cur

    def durations(self, d: torch.Tensor) -> torch.Tensor:
        """-> lens [n] (float, integral)  (model.cpp:1036-1040)."""
        cur = self.lstm("duration_predictor.duration_lstm", d)
        cur = torch.sigmoid(self.lin("duration_predictor.duration_proj", cur, "duration_predictor.duration_proj_bias"))
        self.last_dur_sum = cur.sum(dim=1)
        return torch.clamp(ggml_round(cur.sum(dim=1)), 1.0, 50.0)

    def duration_pass(self, tokens):
        n = len(tokens)
        s_pros, _ = self.styles(n)
        a = self.albert(tokens)
        d = self.prosody(a, s_pros)
        lens = self.durations(d)
        return lens, d, a

    @staticmethod
    def alignment(lens) -> np.ndarray:
        """token index of every frame (the one-hot duration mask of model.cpp:1265-1274 as a gather)."""
        li = np.asarray(lens).astype(np.int64)
        return np.repeat(np.arange(li.shape[0]), li)

    def f0n(self, shared: torch.Tensor, style: torch.Tensor):
        """shared [T, 512] -> f0 [2T], n [2T]  (model.cpp:1169-1190)."""
        outs = []
        for br in ("f0", "n"):
            x = shared.t()
            for i in range(3):
                x = self.ada_block(f"duration_predictor.{br}_blocks.{i}", x, style)
            y = self.lin(f"duration_predictor.{br}_proj_kernel", x.t())[:, 0] + self.w[f"duration_predictor.{br}_proj_bias"]
            outs.append(y)
        return outs[0], outs[1]

    def text_encoder(self, tokens) -> torch.Tensor:
        """-> [n, 512]  (model.cpp:1196-1205)."""
        w = self.w
        tok = torch.as_tensor(np.asarray(tokens, dtype=np.int64))
        cur = w["text_encoder.embedding_weight"][tok]        # [n, 512]
        for i in range(3):
            p = f"text_encoder.layers.{i}."
            y = self.conv(p + "weight", cur.t(), p + "bias", pad=2).t()
            y = ggml_norm(y, 1e-5) * w[p + "gamma"] + w[p + "beta"]
            cur = leaky(y, 0.2)
        return self.lstm("text_encoder.lstm", cur)

    def decoder(self, asr: torch.Tensor, f0: torch.Tensor, n: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
        """asr [T,512], f0/n [2T] -> [512, 2T]  (model.cpp:1215-1231)."""
        f0d = self.conv("decoder.f0_conv_weight", f0[None], "decoder.f0_conv_bias", stride=2, pad=1)     # [1, T]
        nd = self.conv("decoder.n_conv_weight", n[None], "decoder.n_conv_bias", stride=2, pad=1)
        cur = torch.cat([asr.t(), f0d, nd], dim=0)                                                       # [514, T]
        cur = self.ada_block("decoder.encoder_block", cur, style)
        asr_res = (self.lin("decoder.asr_conv_weight", asr) + self.w["decoder.asr_conv_bias"]).t()       # [64, T]
        for i in range(4):
            cur = torch.cat([cur, asr_res, f0d, nd], dim=0)
            cur = self.ada_block(f"decoder.decoder_blocks.{i}", cur, style)
        return cur

    def source(self, f0: torch.Tensor, noise: np.ndarray):
        """f0 [2T] -> har [600T], (mag, phase) [120T+1, 11]  (model.cpp:173-206; util.cpp:140-172)."""
        f0n = f0.numpy().astype(np.float32)
        L = f0n.shape[0]
        hnorm = ((np.arange(9, dtype=np.float32) + np.float32(1.0)) / np.float32(24000.0)).astype(np.float32)
        cur = (f0n[None, :] * hnorm[:, None]).astype(np.float32)                 # [9, 2T]
        cur = np.fmod(cur, np.float32(1.0)).astype(np.float32)
        cs = np.zeros_like(cur)
        run = np.zeros(9, dtype=np.float32)
        for t in range(L):                                                       # serial fp32 prefix sum (ggml-cpu.c:5590-5600)
            run = (run + cur[:, t]).astype(np.float32)
            cs[:, t] = run
        scal = np.float32(300.0 * 2.0 * math.pi)
        cs = (cs * scal).astype(np.float32)
        up = upscale_linear(cs, 300)                                              # [9, 600T]
        f0u = np.repeat(f0n, 300)                                                # nearest
        S = f0u.shape[0]
        nz = noise.reshape(9, S).astype(np.float32)
        voiced = f0u > np.float32(10.0)
        uv = np.where(voiced, np.float32(0.1), np.float32(0.0)).astype(np.float32)
        nsd = np.where(voiced[None, :], np.float32(0.003) * nz, (np.float32(0.1) / np.float32(3.0)) * nz).astype(np.float32)
        sing = (np.sin(up).astype(np.float32) * uv[None, :] + nsd).astype(np.float32)     # [9, S]
        har = torch.tanh(self.lin("decoder.generator.m_source_weight", torch.from_numpy(sing.T.copy()), "decoder.generator.m_source_bias"))[:, 0]
        mag, ph = stft_ref(har.numpy())
        return har, torch.from_numpy(mag), torch.from_numpy(ph), torch.from_numpy(sing)

    def generator(self, x: torch.Tensor, mag: torch.Tensor, ph: torch.Tensor, style: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
        """x [512, 2T], (mag, ph) [120T+1, 11] -> pcm [600T]  (model.cpp:195-244)."""
        w, kv = self.w, self.kv
        g = "decoder.generator."
        G = "kokoro.decoder.generator."
        har = torch.cat([mag, ph], dim=1).t().contiguous()           # [22, frames]
        cur = x
        for i in range(2):
            cur = leaky(cur, 0.1)
            s, p = kv[f"{G}up_convs.{i}.stride"], kv[f"{G}up_convs.{i}.padding"]
            cur = F.conv_transpose1d(cur[None], w[f"{g}ups.{i}.weight"], None, stride=s, padding=p)[0] + w[f"{g}ups.{i}.bias"][:, None]
            if i == 1:
                cur = torch.cat([cur[:, 1:2], cur], dim=1)            # 1-sample reflect pad on the left
            ns, npad = kv[f"{G}noise_blocks.{i}.stride"], kv[f"{G}noise_blocks.{i}.padding"]
            xs = self.conv(f"{g}noise_blocks.{i}.conv_weight", har, f"{g}noise_blocks.{i}.conv_bias", stride=ns, pad=npad)
            pads = [kv[f"{G}noise_blocks.{i}.res_block.{j}.padding"] for j in range(3)]
            dils = [kv[f"{G}noise_blocks.{i}.res_block.{j}.dilation"] for j in range(3)]
            xs = self.gen_resblock(f"{g}noise_blocks.{i}.resblock", xs, style, pads, dils)
            cur = cur + xs
            if taps is not None:
                taps[f"gen_in{i}"] = cur.clone()
            acc = None
            for j in range(3):
                r = i * 3 + j
                pads = [kv[f"{G}res_blocks.{r}.{q}.padding"] for q in range(3)]
                dils = [kv[f"{G}res_blocks.{r}.{q}.dilation"] for q in range(3)]
                o = self.gen_resblock(f"{g}resblocks.{r}", cur, style, pads, dils)
                acc = o if acc is None else acc + o
            cur = acc / np.float32(3.0)
            if taps is not None:
                taps[f"gen_out{i}"] = cur.clone()
        cur = leaky(cur, 0.01)
        cur = self.conv(g + "conv_post_weight", cur, g + "conv_post_bias", pad=kv.get(G + "padding", 3))   # [22, frames]
        spec = torch.exp(cur[:11]).t().numpy()
        phase = torch.sin(cur[11:]).t().numpy()
        if taps is not None:
            taps["spec"] = torch.from_numpy(spec.copy()); taps["phase"] = torch.from_numpy(phase.copy())
        return torch.from_numpy(istft_ref(spec, phase))

    def generation_pass(self, tokens, lens, d: torch.Tensor, noise_skip: int = 0, taps: dict | None = None) -> torch.Tensor:
        n = len(tokens)
        s_pros, s_dec = self.styles(n)
        al = torch.from_numpy(self.alignment(lens))
        T = al.shape[0]
        en = d[al]                                                   # [T, 640]
        shared = self.lstm("duration_predictor.shared_lstm", en)     # [T, 512]
        f0, nn = self.f0n(shared, s_pros)
        t_en = self.text_encoder(tokens)
        asr = t_en[al]                                               # [T, 512]
        dec = self.decoder(asr, f0, nn, s_dec)                       # [512, 2T]
        noise = minstd_uniform(9 * 600 * T, noise_skip)
        har, mag, ph, sing = self.source(f0, noise)
        if taps is not None:
            taps.update(dict(en=en, shared=shared, f0=f0, n=nn, t_en=t_en, asr=asr, dec=dec, har=har, mag=mag, ph=ph, sing=sing))
        return self.generator(dec, mag, ph, s_dec, taps)

    def run(self, tokens, noise_skip: int = 0, taps: dict | None = None):
        """kokoro_runner::run (model.cpp:1277-1325): tokens -> (lens, pcm)."""
        lens, d, a = self.duration_pass(tokens)
        if taps is not None:
            taps.update(dict(albert=a, d=d, lens=lens))
        pcm = self.generation_pass(tokens, lens.numpy(), d, noise_skip, taps)
        return lens.numpy().astype(np.float32), pcm.numpy().astype(np.float32)
