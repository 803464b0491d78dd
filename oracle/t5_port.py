"""oracle/t5_port.py -- TEST INFRASTRUCTURE (the product never imports it).  CPU restatement of the reference's T5 conditional-prompt encoder
(src/models/parler/t5/model.cpp:183-320: build_t5_norm, build_t5_pos_bias, build_t5_graph, set_inputs), in torch fp32.
Pinned against oracle/_ref/t5_ref (the unmodified t5_runner::run on explicit token ids) by tests/golden/t5_vectors.npz + tests/test_oracle_port.py.

Restated semantics: token embedding rows (ggml_get_rows); per layer: RMS norm (eps 1e-6, no mean subtraction, double-accumulated like ggml_rms_norm) x weight;
bias-free q / k / v projections, head size fixed at 64; scores q.k WITHOUT 1/sqrt(d) plus the relative-position bias of layer 0's table, shared by all layers,
looked up by bucket(key - query) with the reference's own bucket arithmetic -- INCLUDING its integer division inside the logarithm (model.cpp:330: ab_rpos /
max_exact are ints); softmax over all positions (the mask is all zeros: bidirectional, no padding); o projection + residual; gated GELU feed-forward
gelu(wi_0 x) * (wi_1 x) -> wo (ggml_gelu = the fp16 table) + residual; final RMS norm; optional down projection + bias to the Parler decoder's width."""
from __future__ import annotations

import math

import numpy as np
import torch

try:
    from .kokoro_port import gelu_f16_lut
    from .parler_port import quant_mm, unpack_blocks
except ImportError:
    from kokoro_port import gelu_f16_lut
    from parler_port import quant_mm, unpack_blocks


def relative_bucket(key_pos: int, query_pos: int, relative_attn_buckets: int = 32) -> int:
    """t5_runner::set_inputs (model.cpp:318-332), one entry: i = key position, ii = query position (the bias is added to kq[key, query], model.cpp:259-260)"""
    n_buckets = relative_attn_buckets // 2
    max_exact = n_buckets // 2
    log_den = np.float32(math.log(128.0 / max_exact))                       # `float logarithmic_denominator`
    rpos = key_pos - query_pos
    ab = abs(rpos)
    if ab < max_exact:
        v = ab
    else:
        v = min(n_buckets - 1, max_exact + int((math.log(ab // max_exact) / float(log_den)) * max_exact))      # ab // max_exact: the reference divides ints
    return (n_buckets if rpos > 0 else 0) + v


class T5Port:
    def __init__(self, gguf_path: str, threads: int = 8):
        import gguf
        torch.set_num_threads(threads)
        rd = gguf.GGUFReader(gguf_path)
        self.w, self.f16, self.q = {}, set(), {}
        for t in rd.tensors:
            if t.tensor_type.name in ("Q8_0", "Q5_0", "Q4_0"):            # block-quantised matrix: (scales, integer values), see parler_port.quant_mm
                d, qv = unpack_blocks(np.array(t.data), t.tensor_type.name)
                self.q[t.name] = (torch.from_numpy(d), torch.from_numpy(qv))
                continue
            self.w[t.name] = torch.from_numpy(np.array(t.data).astype(np.float32))
            if t.tensor_type.name == "F16":
                self.f16.add(t.name)
        self.kv = {}
        for k, f in rd.fields.items():
            if len(f.data) == 1 and f.types and f.types[0].name in ("UINT32",):
                self.kv[k] = int(f.parts[f.data[0]][0])
        self.layers = self.kv["t5encoder.block_count"]; self.hidden = self.kv["t5encoder.embedding_length"]; self.heads = self.kv["t5encoder.attention.head_count"]
        self.hd = 64                                                            # model.h:46, not a GGUF key
        assert self.heads * self.hd == self.hidden

    def mm(self, x, name):
        if name in self.q:
            return quant_mm(x, *self.q[name])
        if name in self.f16:
            x = x.half().float()
        return x @ self.w[name].t()

    @staticmethod
    def norm(x, w):
        ms = (x.double() * x.double()).mean(dim=-1, keepdim=True).float()      # ggml_rms_norm: ggml_float sum of squares, mean, 1/sqrtf(mean + eps)
        return x * (1.0 / torch.sqrt(ms + np.float32(1e-6))) * w

    def run(self, tokens, taps=None):
        """t5_runner::run: [n] token ids -> [n, output_size] encoding"""
        ids = torch.as_tensor(np.asarray(tokens, np.int64))
        n = ids.numel()
        x = self.w["t5encoder.token_embd"][ids]
        rel = self.w["t5encoder.enc.blk.0.attn_rel_b"]                          # [buckets, heads]
        bucket = torch.tensor([[relative_bucket(k, q) for k in range(n)] for q in range(n)])       # [query, key]
        bias = rel[bucket].permute(2, 0, 1)                                     # [heads, query, key]
        for l in range(self.layers):
            b = f"t5encoder.enc.blk.{l}"
            cur = self.norm(x, self.w[b + ".attn_norm"])
            q = self.mm(cur, b + ".attn_q").view(n, self.heads, self.hd).transpose(0, 1)
            k = self.mm(cur, b + ".attn_k").view(n, self.heads, self.hd).transpose(0, 1)
            v = self.mm(cur, b + ".attn_v").view(n, self.heads, self.hd).transpose(0, 1)
            s = q @ k.transpose(1, 2) + bias                                    # softmax scale 1.0 (model.cpp:262)
            p = torch.softmax(s.double(), dim=-1).float()
            att = (p @ v).transpose(0, 1).reshape(n, self.hidden)
            x = self.mm(att, b + ".attn_o") + x
            cur = self.norm(x, self.w[b + ".ffn_norm"])
            cur = gelu_f16_lut(self.mm(cur, b + ".ffn_up")) * self.mm(cur, b + ".ffn_gate")         # wi_0 = ffn_up (GELU'd), wi_1 = ffn_gate (model.cpp:14-16,278-279)
            x = self.mm(cur, b + ".ffn_down") + x
            if taps is not None:
                taps[f"layer{l}"] = x.clone()
        x = self.norm(x, self.w["t5encoder.enc.final_layer_norm"])
        if "t5encoder.down_proj" in self.w:
            x = self.mm(x, "t5encoder.down_proj")
        if "t5encoder.down_proj_bias" in self.w:
            x = x + self.w["t5encoder.down_proj_bias"]
        return x.numpy()
