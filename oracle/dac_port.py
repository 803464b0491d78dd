"""oracle/dac_port.py -- TEST INFRASTRUCTURE.  CPU restatement of the reference DAC codec decoder.

The checker for the CUDA codec path (never imported by the product): what dac_runner::run computes for one utterance
(reference src/decoder/dac_model.cpp:100-123,146-170 and src/decoder/general_neural_audio_codec.cpp:133-172), in CPU torch fp32.
Pinned against the compiled reference (oracle/_ref/dac_ref) by tests/golden/make_golden.py + tests/test_oracle_port.py.

ggml numerics that matter here (all file:line of /root/reference):
  * ggml_conv_1d picks the im2col type from the operand types (ggml/src/ggml.c:3877-3881): with an F32 kernel and F32 input the whole
    convolution is fp32; with an F16 kernel the activations are re-rounded to fp16 (products accumulated in fp32);
  * ggml_conv_transpose_1d with an F32 kernel is fp32 (ggml/src/ggml-cpu/ggml-cpu.c:10101-10200);
  * snake_1d: x + sin^2(alpha x) * (1 / alpha)  (src/util.cpp:86-101);
  * the quantizer: codebook rows (get_rows) -> 1x1 conv + bias per head, heads summed in order (dac_model.cpp:100-123).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _h(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).to(torch.float32)


class DacPort:
    def __init__(self, gguf_path: str, threads: int = 8):
        import gguf
        torch.set_num_threads(threads)
        rd = gguf.GGUFReader(gguf_path)
        self.w: dict[str, torch.Tensor] = {}
        self.f16: dict[str, bool] = {}
        for t in rd.tensors:
            name = t.name[len("audio_encoder."):] if t.name.startswith("audio_encoder.") else t.name
            arr = np.array(t.data)
            self.f16[name] = arr.dtype == np.float16
            self.w[name] = torch.from_numpy(arr.astype(np.float32))
        self.kv = {}
        for k, f in rd.fields.items():
            if len(f.data) == 1 and f.types and f.types[0].name in ("UINT32",):
                self.kv[k] = int(f.parts[f.data[0]][0])
        self.n_heads = self.kv.get("output_heads", 9)
        self.strides = [self.kv[f"dac.dac_layer_stride_{i}"] for i in range(4)]
        self.pads = [self.kv[f"dac.dac_layer_padding_{i}"] for i in range(4)]

    def conv(self, name: str, x: torch.Tensor, pad=0, dil=1) -> torch.Tensor:
        W = self.w[name + ".weight"]
        xi = _h(x) if self.f16[name + ".weight"] else x
        return F.conv1d(xi[None], W, None, padding=pad, dilation=dil)[0] + self.w[name + ".bias"][:, None]

    @staticmethod
    def snake(alpha: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        a = alpha.reshape(-1, 1)
        s = torch.sin(x * a)
        return x + (s * s) * (1.0 / a)

    def embed(self, codes: np.ndarray) -> torch.Tensor:
        """codes [frames, n_heads] -> [latent, frames]  (dac_build_audio_inputs, dac_model.cpp:100-123)."""
        x = None
        for i in range(self.n_heads):
            rows = self.w[f"quantizers.{i}.codebook.weight"][torch.from_numpy(codes[:, i].astype(np.int64))]    # [frames, 8]
            e = self.conv(f"quantizers.{i}.out_proj", rows.t().contiguous())
            x = e if x is None else x + e
        return x

    def decode(self, codes: np.ndarray, taps: dict | None = None) -> np.ndarray:
        x = self.embed(codes)
        if taps is not None: taps["embd"] = x.numpy().copy()
        x = self.conv("initial", x, pad=3)
        if taps is not None: taps["initial"] = x.numpy().copy()
        for l in range(1, 5):
            b = f"decoder_block.{l}"
            x = self.snake(self.w[b + ".final.alpha"], x)
            x = F.conv_transpose1d(x[None], self.w[b + ".final.weight"], None, stride=self.strides[l - 1], padding=self.pads[l - 1])[0]
            x = x + self.w[b + ".final.bias"][:, None]
            for i in range(3):
                r = f"{b}.residual_unit.{i}.res"
                y = self.snake(self.w[r + ".initial.alpha"], x)
                y = self.conv(r + ".initial", y, pad=3 ** (i + 1), dil=3 ** i)
                y = self.snake(self.w[r + ".final.alpha"], y)
                y = self.conv(r + ".final", y)
                x = y + x
            if taps is not None: taps[f"layer{l}"] = x.numpy().copy()
        x = self.snake(self.w["final.alpha"], x)
        x = self.conv("final", x, pad=3)
        return torch.tanh(x)[0].numpy()
