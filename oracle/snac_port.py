"""oracle/snac_port.py -- TEST INFRASTRUCTURE.  CPU restatement of the reference SNAC codec decoder.

The checker for a future CUDA SNAC path (never imported by the product): what snac_runner::run computes for one utterance
(reference src/decoder/snac_model.cpp:86-208, src/decoder/general_neural_audio_codec.cpp:133-172), in CPU torch fp32, including the
reference's noise source: a process-wide default-seeded std::default_random_engine (minstd_rand0) driving
std::normal_distribution<float>(0, 1) (src/util.cpp:74-80), restated from libstdc++'s polar Box-Muller
(bits/random.tcc, normal_distribution::operator()): pairs (x, y) uniform in (-1, 1) rejected until 0 < r2 <= 1,
mult = sqrt(-2 log(r2) / r2); the call returns y*mult and saves x*mult for the next call.
Pinned against oracle/_ref/snac_ref by tests/golden/make_golden.py + tests/test_oracle_port.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

_M = (1 << 31) - 1
_A = 16807


class MinstdNormal:
    """The reference's static normal generator: state persists across calls (utterances decoded in one process share the stream)."""

    def __init__(self):
        self.x = 1                # std::default_random_engine default seed
        self.saved = None

    def _canonical(self) -> np.float32:
        self.x = (self.x * _A) % _M
        u = np.float32(self.x - 1) / np.float32(2147483646.0)      # generate_canonical<float, 24>: one draw; the divisor is 2^31 in fp32
        return np.float32(0.99999994) if u >= np.float32(1.0) else u

    def draw(self, count: int) -> np.ndarray:
        out = np.empty(count, np.float32)
        f2, f1 = np.float32(2.0), np.float32(1.0)
        for i in range(count):
            if self.saved is not None:
                out[i] = self.saved
                self.saved = None
                continue
            while True:
                x = np.float32(f2 * self._canonical() - f1)
                y = np.float32(f2 * self._canonical() - f1)
                r2 = np.float32(x * x + y * y)
                if not (r2 > f1 or r2 == np.float32(0.0)):
                    break
            mult = np.float32(np.sqrt(np.float32(np.float32(-2.0) * np.float32(np.log(r2)) / r2)))
            self.saved = np.float32(x * mult)
            out[i] = np.float32(y * mult)
        return out


class SnacPort:
    REPEATS = (4, 2, 1)
    NOISE_STEPS = (8, 64, 256, 512)

    def __init__(self, gguf_path: str, threads: int = 8):
        import gguf
        torch.set_num_threads(threads)
        rd = gguf.GGUFReader(gguf_path)
        self.w = {}
        for t in rd.tensors:
            name = t.name[len("snac."):] if t.name.startswith("snac.") else t.name
            self.w[name] = torch.from_numpy(np.array(t.data).astype(np.float32))
        self.kv = {}
        for k, f in rd.fields.items():
            if len(f.data) == 1 and f.types and f.types[0].name in ("UINT32",):
                self.kv[k] = int(f.parts[f.data[0]][0])
        self.strides = [self.kv[f"snac.snac_layer_stride_{i}"] for i in range(4)]
        self.pads = [self.kv[f"snac.snac_layer_padding_{i}"] for i in range(4)]
        self.rng = MinstdNormal()

    @staticmethod
    def snake(alpha, x):
        a = alpha.reshape(-1, 1)
        s = torch.sin(x * a)
        return x + (s * s) * (1.0 / a)

    def conv(self, name, x, pad=0, dil=1, groups=1, bias=True):
        y = F.conv1d(x[None], self.w[name + ".weight"] if name + ".weight" in self.w else self.w[name], None, padding=pad, dilation=dil, groups=groups)[0]
        return y + self.w[name + ".bias"][:, None] if bias else y

    def embed(self, codes) -> torch.Tensor:
        x = None
        for i, rep in enumerate(self.REPEATS):
            rows = self.w[f"quantizers.{i}.codebook.weight"][torch.from_numpy(np.asarray(codes[i]).astype(np.int64))]   # [len, 8]
            e = self.conv(f"quantizers.{i}.out_proj", rows.t().contiguous())
            if rep > 1:
                e = e.repeat_interleave(rep, dim=1)
            x = e if x is None else x + e
        return x

    def decode(self, codes, noise: np.ndarray | None = None) -> np.ndarray:
        """codes: [coarse L/4, medium L/2, fine L].  noise: the 840*L normal draws (None -> drawn from this port's persistent generator)."""
        L = len(codes[2])
        if noise is None:
            noise = self.rng.draw(sum(self.NOISE_STEPS) * L)
        x = self.embed(codes)
        C = x.shape[0]
        x = self.conv("in", x, pad=3, groups=C)
        x = self.conv("up", x)
        off = 0
        for l in range(4):
            b = f"layers.{l}"
            x = self.snake(self.w[b + ".alpha"], x)
            x = F.conv_transpose1d(x[None], self.w[b + ".weight"], None, stride=self.strides[l], padding=self.pads[l])[0] + self.w[b + ".bias"][:, None]
            n = torch.from_numpy(noise[off:off + self.NOISE_STEPS[l] * L]); off += self.NOISE_STEPS[l] * L
            x = x + F.conv1d(x[None], self.w[b + ".noise_weight"], None)[0] * n[None, :]
            C = x.shape[0]
            for i in range(3):
                r = f"{b}.residual_unit.{i}.res"
                y = self.snake(self.w[r + ".initial.alpha"], x)
                y = self.conv(r + ".initial", y, pad=3 ** (i + 1), dil=3 ** i, groups=C)
                y = self.snake(self.w[r + ".final.alpha"], y)
                y = self.conv(r + ".final", y)
                x = y + x
        x = self.snake(self.w["alpha_out"], x)
        x = self.conv("final", x, pad=3)
        return torch.tanh(x)[0].numpy()
