"""oracle/dia_port.py -- TEST INFRASTRUCTURE.  CPU restatement of the reference Dia encoder pass + CFG-paired decode loop.

The checker for a future CUDA path (never imported by the product): what dia_runner::decode computes (reference
src/models/dia/model.cpp:324-637,705-737) and the token loop of generate_from_batch with check_stopping (model.cpp:806-864) under the
greedy sampler, in CPU torch fp32 with explicit caches.  Pinned against oracle/_ref/dia_ref by tests/golden/make_golden.py +
tests/test_oracle_port.py.

Reference semantics restated, quirks included:
  * two sequences throughout: the conditional prompt (byte tokens padded with 0 to the encoder context C) and an all-zero unconditional one;
    logits = cond + 3 * (cond - uncond) (cfg_scale, src/util.cpp:175-200 -- its "mask above max_output" branch is overwritten and has no effect);
  * encoder: RMSNorm (eps 1e-5) * weight, NeoX RoPE (theta base 1e4) on q and k, softmax scale 1.0 (no 1/sqrt(d)), one mask for both sequences
    (positions < prompt length see each other, padded positions see each other), SwiGLU;
  * cross-attention: keys only for the first `prompt length` encoder positions (RoPE'd), the rest of the C-long key cache stays zero and is
    attended to without a mask; values from ALL C encoder positions; the decoder's cross query is RoPE'd with the decode position;
  * decoder self-attention: GQA with k, v repeat-interleaved to the query heads before caching, no mask, scale 1.0;
  * head i is fed BOS until the decode position exceeds i; nine heads of 1028 logits, per-head argmax (first maximum wins).
"""
from __future__ import annotations

import numpy as np
import torch

try:
    from .parler_port import quant_mm, unpack_blocks
except ImportError:                               # imported as a top-level module (oracle/ on sys.path)
    from parler_port import quant_mm, unpack_blocks


class DiaPort:
    MAX_DELAY = 15
    DELAY = (0, 8, 9, 10, 11, 12, 13, 14, 15)

    def __init__(self, gguf_path: str, threads: int = 8):
        import gguf
        torch.set_num_threads(threads)
        rd = gguf.GGUFReader(gguf_path)
        self.w = {}
        self.f16 = set()          # F16 matrices: ggml_mul_mat rounds the activations to fp16 before the product
        self.q = {}               # block-quantised matrices: name -> (scales [N, nb], integer values [N, nb, 32]) (see parler_port.unpack_blocks)
        for t in rd.tensors:
            if t.name.startswith("dia."):
                name = t.name[len("dia."):]
                if t.tensor_type.name in ("Q8_0", "Q5_0", "Q4_0"):
                    d, qv = unpack_blocks(np.array(t.data), t.tensor_type.name)
                    self.q[name] = (torch.from_numpy(d), torch.from_numpy(qv))
                    self.w[name] = torch.from_numpy((d[:, :, None] * qv).reshape(d.shape[0], -1))       # dequantize_row: what ggml_get_rows returns
                    continue
                self.w[name] = torch.from_numpy(np.array(t.data).astype(np.float32))
                if t.tensor_type.name == "F16":
                    self.f16.add(name)
        self.kv = {}
        for k, f in rd.fields.items():
            if len(f.data) == 1 and f.types and f.types[0].name in ("UINT32",):
                self.kv[k] = int(f.parts[f.data[0]][0])
        g = self.kv
        self.enc_layers = g["dia.encoder.layers"]; self.dec_layers = g["dia.decoder.layers"]
        self.heads = g["dia.decoder.attn_heads"]; self.qh = g["dia.decoder.query_heads"]; self.enc_heads = g["dia.encoder.attn_heads"]
        self.hd = g["dia.attn_head_size"]; self.C = g["dia.encoder.max_context_length"]; self.n_out = g["dia.decoder.output_heads"]
        self.vocab = g["dia.decoder.output_vocab_size"]; self.bos = g["dia.bos_token_id"]; self.eos = g["dia.eos_token_id"]; self.pad = g["dia.pad_token_id"]
        self.max_gen = g["dia.decoder.max_generation_size"]

    def mm(self, x, name):
        """ggml_mul_mat(weight, x): exact products of fp16-rounded activations with F16 weights, fp32 accumulation; plain fp32 for F32 weights."""
        if name in self.q:
            return quant_mm(x, *self.q[name])
        if name in self.f16:
            x = x.half().float()
        return x @ self.w[name].t()

    @staticmethod
    def rms(x, w):
        ms = (x.double() ** 2).mean(dim=-1, keepdim=True)
        return x * (1.0 / torch.sqrt(ms.float() + 1e-5)) * w

    def rope(self, x, positions):
        """x [..., n, heads, hd] with n == len(positions); NeoX pairs (i, i + hd/2), theta advanced by repeated multiplication like ggml."""
        hd = x.shape[-1]; half = hd // 2
        theta_scale = np.float32(np.power(np.float32(10000.0), np.float32(-2.0) / np.float32(hd)))
        cos = np.empty((len(positions), half), np.float32); sin = np.empty((len(positions), half), np.float32)
        for t, p in enumerate(positions):
            theta = np.float32(p)
            for i in range(half):
                cos[t, i] = np.cos(theta); sin[t, i] = np.sin(theta)
                theta = np.float32(theta * theta_scale)
        c = torch.from_numpy(cos)[:, None, :]; s = torch.from_numpy(sin)[:, None, :]
        x0, x1 = x[..., :half], x[..., half:]
        return torch.cat([x0 * c - x1 * s, x0 * s + x1 * c], dim=-1)

    @staticmethod
    def softmax(s):
        return torch.softmax(s.double(), dim=-1).float()

    def encode(self, prompt):
        C, S = self.C, len(prompt)
        tok = torch.zeros(2, C, dtype=torch.int64)
        tok[0, :S] = torch.from_numpy(np.asarray(prompt).astype(np.int64))
        x = self.w["encoder.embedding"][tok]                                   # [2, C, 1024]
        mask = torch.full((C, C), float("-inf"))
        mask[:S, :S] = 0.0; mask[S:, S:] = 0.0
        pos = list(range(C))
        H, hd = self.enc_heads, self.hd
        for l in range(self.enc_layers):
            b = f"encoder.layers.{l}"
            res = x
            cur = self.rms(x, self.w[b + ".pre_sa_norm"])
            q = self.rope((self.mm(cur, b + ".q_proj")).reshape(2, C, H, hd), pos)
            k = self.rope((self.mm(cur, b + ".k_proj")).reshape(2, C, H, hd), pos)
            v = (self.mm(cur, b + ".v_proj")).reshape(2, C, H, hd)
            p = self.softmax(torch.einsum("bnhd,bthd->bhnt", q, k) + mask[None, None])
            o = torch.einsum("bhnt,bthd->bnhd", p, v).reshape(2, C, H * hd)
            x = self.mm(o, b + ".o_proj") + res
            res = x
            cur = self.rms(x, self.w[b + ".post_sa_norm"])
            gte = self.mm(cur, b + ".gate")
            cur = (gte / (1.0 + torch.exp(-gte))) * (self.mm(cur, b + ".up"))
            x = self.mm(cur, b + ".wo") + res
        enc = self.rms(x, self.w["encoder.norm"])
        self.ck, self.cv = [], []
        H = self.heads
        for l in range(self.dec_layers):
            b = f"decoder.layers.{l}"
            k = torch.zeros(2, C, H, hd)
            k[:, :S] = self.rope((self.mm(enc[:, :S], b + ".cross_k_proj")).reshape(2, S, H, hd), list(range(S)))
            self.ck.append(k)
            self.cv.append((self.mm(enc, b + ".cross_v_proj")).reshape(2, C, H, hd))
        self.k = [None] * self.dec_layers; self.v = [None] * self.dec_layers
        self.pos = 0

    def step(self, audio_tokens) -> np.ndarray:
        """one decode step for the CFG pair; returns the combined logits [n_out, vocab]."""
        x = None
        for i in range(self.n_out):
            e = self.w[f"decoder.embeddings.{i}"][int(audio_tokens[i])]
            x = e if x is None else e + x
        x = x[None, :].repeat(2, 1)                                             # the same audio tokens for both sequences
        H, hd, rep = self.heads, self.hd, self.qh
        for l in range(self.dec_layers):
            b = f"decoder.layers.{l}"
            res = x
            cur = self.rms(x, self.w[b + ".pre_sa_norm"])
            q = self.rope((self.mm(cur, b + ".self_q_proj")).reshape(2, 1, H, hd), [self.pos])
            k = self.rope((self.mm(cur, b + ".self_k_proj")).reshape(2, 1, H // rep, hd), [self.pos]).repeat_interleave(rep, dim=2)
            v = (self.mm(cur, b + ".self_v_proj")).reshape(2, 1, H // rep, hd).repeat_interleave(rep, dim=2)
            self.k[l] = k if self.k[l] is None else torch.cat([self.k[l], k], 1)
            self.v[l] = v if self.v[l] is None else torch.cat([self.v[l], v], 1)
            p = self.softmax(torch.einsum("bnhd,bthd->bhnt", q, self.k[l]))
            o = torch.einsum("bhnt,bthd->bnhd", p, self.v[l]).reshape(2, H * hd)
            x = self.mm(o, b + ".self_o_proj") + res
            res = x
            cur = self.rms(x, self.w[b + ".pre_ca_norm"])
            q = self.rope((self.mm(cur, b + ".cross_q_proj")).reshape(2, 1, H, hd), [self.pos])
            p = self.softmax(torch.einsum("bnhd,bthd->bhnt", q, self.ck[l]))
            o = torch.einsum("bhnt,bthd->bnhd", p, self.cv[l]).reshape(2, H * hd)
            x = self.mm(o, b + ".cross_o_proj") + res
            res = x
            cur = self.rms(x, self.w[b + ".pre_mlp_norm"])
            gte = self.mm(cur, b + ".gate")
            cur = (gte / (1.0 + torch.exp(-gte))) * (self.mm(cur, b + ".up"))
            x = self.mm(cur, b + ".wo") + res
        x = self.rms(x, self.w["decoder.norm"])
        lg = torch.stack([self.mm(x, f"decoder.heads.{i}") for i in range(self.n_out)])    # [n_out, 2, vocab]
        cond, uncond = lg[:, 0], lg[:, 1]
        self.pos += 1
        return (cond + 3.0 * (cond - uncond)).numpy()

    def greedy(self, prompt, steps: int, teacher=None):
        self.encode(prompt)
        audio = [self.bos] * self.n_out
        toks, logits = [], []
        delay_steps = -1
        for _ in range(steps):
            # check_stopping (model.cpp:806-823)
            if delay_steps == -1 and (audio[0] == self.eos or self.pos >= self.max_gen - self.MAX_DELAY):
                delay_steps = self.MAX_DELAY
            if delay_steps > 0:
                after = self.MAX_DELAY - delay_steps
                for i, d in enumerate(self.DELAY):
                    if after == d: audio[i] = self.eos
                    elif after > d: audio[i] = self.pad
                delay_steps -= 1
            if delay_steps == 0:
                break
            lg = self.step(audio)
            last = lg.argmax(axis=1)
            toks.append(last.astype(np.int32)); logits.append(lg)
            if teacher is not None:                  # teacher-forced comparison: feed the given tokens back, keep reporting the produced ones
                last = np.asarray(teacher[len(toks) - 1])
            audio = [int(last[i]) if self.pos > i else self.bos for i in range(self.n_out)]
        return np.stack(toks), np.stack(logits)
