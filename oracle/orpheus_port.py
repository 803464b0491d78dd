"""oracle/orpheus_port.py -- TEST INFRASTRUCTURE.  CPU restatement of the reference Orpheus decode step (llama-3 style).

The checker for a future CUDA AR-decode path (never imported by the product): what orpheus_runner::decode computes per call
(reference src/models/orpheus/model.cpp:122-131,196-312,342-353) plus the greedy sampler (src/sampler.cpp: argmax, first maximum wins),
in CPU torch fp32 with an explicit KV cache.  Pinned against oracle/_ref/orpheus_ref by tests/golden/make_golden.py +
tests/test_oracle_port.py.

Reference semantics restated: RMSNorm eps 1e-5 times weight; q/k/v/o and SwiGLU projections without bias; NeoX-style RoPE
(ggml_rope_ext mode 2 over the whole head, theta base 5e5, per-pair frequency factors from the `rope_frequencies` tensor, theta advanced
by repeated multiplication like ggml); K and V cached already expanded to the query heads (each kv head repeated 3x, model.cpp:251);
scores scaled by 1/sqrt(head) with a causal -inf mask inside the softmax; logits for the last position only.
"""
from __future__ import annotations

import numpy as np
import torch


class OrpheusPort:
    def __init__(self, gguf_path: str, threads: int = 8):
        import gguf
        torch.set_num_threads(threads)
        rd = gguf.GGUFReader(gguf_path)
        self.w = {}
        for t in rd.tensors:
            if t.name.startswith("orpheus."):
                self.w[t.name[len("orpheus."):]] = torch.from_numpy(np.array(t.data).astype(np.float32))
        self.kv = {}
        for k, f in rd.fields.items():
            if len(f.data) == 1 and f.types and f.types[0].name in ("UINT32",):
                self.kv[k] = int(f.parts[f.data[0]][0])
        self.layers = self.kv["orpheus.layers"]
        self.heads = self.kv["orpheus.attn_heads"]; self.kv_heads = self.kv["orpheus.kv_attn_heads"]; self.hd = self.kv["orpheus.head_dim"]
        self.vocab = self.kv["orpheus.vocab_size"]
        self.reset()

    def reset(self):
        self.k = [None] * self.layers
        self.v = [None] * self.layers
        self.pos = 0

    @staticmethod
    def rms(x, w):
        ms = (x.double() ** 2).mean(dim=-1, keepdim=True)                      # ggml_rms_norm accumulates in ggml_float (double)
        scale = (1.0 / torch.sqrt(ms.float() + 1e-5))
        return x * scale * w

    def rope(self, x, pos0):
        """x [n, heads, hd]; NeoX pairs (i, i + hd/2)."""
        n, H, hd = x.shape
        half = hd // 2
        ff = self.w["rope_frequencies"].numpy().astype(np.float32)
        theta_scale = np.float32(np.power(np.float32(500000.0), np.float32(-2.0) / np.float32(hd)))
        cos = np.empty((n, half), np.float32); sin = np.empty((n, half), np.float32)
        for t in range(n):
            theta = np.float32(pos0 + t)
            for i in range(half):
                th = np.float32(theta / ff[i])
                cos[t, i] = np.cos(th); sin[t, i] = np.sin(th)
                theta = np.float32(theta * theta_scale)
        c = torch.from_numpy(cos)[:, None, :]; s = torch.from_numpy(sin)[:, None, :]
        x0, x1 = x[..., :half], x[..., half:]
        return torch.cat([x0 * c - x1 * s, x0 * s + x1 * c], dim=-1)

    def step(self, tokens) -> torch.Tensor:
        """tokens: the new token ids (whole prompt at first, then one); returns logits [vocab] of the last position."""
        tok = torch.from_numpy(np.asarray(tokens).astype(np.int64))
        n = tok.numel()
        x = self.w["embed_tokens"][tok]
        rep = self.heads // self.kv_heads
        for l in range(self.layers):
            b = f"layers.{l}"
            res = x
            cur = self.rms(x, self.w[b + ".input_layernorm"])
            q = (cur @ self.w[b + ".self_attn.q_proj"].t()).reshape(n, self.heads, self.hd)
            k = (cur @ self.w[b + ".self_attn.k_proj"].t()).reshape(n, self.kv_heads, self.hd)
            v = (cur @ self.w[b + ".self_attn.v_proj"].t()).reshape(n, self.kv_heads, self.hd)
            k = self.rope(k, self.pos).repeat_interleave(rep, dim=1)
            v = v.repeat_interleave(rep, dim=1)
            q = self.rope(q, self.pos)
            self.k[l] = k if self.k[l] is None else torch.cat([self.k[l], k], 0)
            self.v[l] = v if self.v[l] is None else torch.cat([self.v[l], v], 0)
            K, V = self.k[l], self.v[l]                                        # [T, heads, hd]
            T = K.shape[0]
            s = torch.einsum("nhd,thd->hnt", q, K) * (1.0 / np.sqrt(np.float32(self.hd)))
            mask = torch.full((n, T), 0.0)
            for i in range(n):
                mask[i, self.pos + i + 1:] = float("-inf")
            p = torch.softmax((s + mask[None]).double(), dim=-1).float()
            o = torch.einsum("hnt,thd->nhd", p, V).reshape(n, self.heads * self.hd)
            x = o @ self.w[b + ".self_attn.o_proj"].t() + res
            res2 = x
            cur = self.rms(x, self.w[b + ".post_attention_layernorm"])
            g = cur @ self.w[b + ".mlp.gate_proj"].t()
            u = cur @ self.w[b + ".mlp.up_proj"].t()
            cur = (g / (1.0 + torch.exp(-g))) * u                              # ggml_silu_f32: x / (1 + expf(-x))
            x = cur @ self.w[b + ".mlp.down_proj"].t() + res2
        x = self.rms(x, self.w["norm"])
        self.pos += n
        return x[-1] @ self.w["lm_head"].t()

    def greedy(self, prompt, steps: int):
        self.reset()
        toks, logits = [], []
        cur = list(prompt)
        for _ in range(steps):
            lg = self.step(cur).numpy()
            logits.append(lg)
            t = int(np.argmax(lg))                                              # first maximum wins, like sampler::max
            toks.append(t)
            cur = [t]
        return np.array(toks, np.int32), np.stack(logits)
