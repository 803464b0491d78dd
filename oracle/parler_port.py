"""oracle/parler_port.py -- TEST INFRASTRUCTURE.  CPU restatement of the reference Parler-TTS decode loop.

The checker for a future CUDA path (never imported by the product): what parler_tts_runner::decode computes per call (reference
src/models/parler/model.cpp:387-470,520-614) and the token loop of generate_from_batch with its delay pattern (model.cpp:762-786) under
the greedy sampler (src/sampler.cpp: per-head argmax, first maximum wins), in CPU torch fp32 with an explicit KV cache.
Pinned against oracle/_ref/parler_ref by tests/golden/make_golden.py + tests/test_oracle_port.py.

Reference semantics restated: learned positional embeddings added to the prompt / summed codebook embeddings; pre-LayerNorm blocks
(ggml_norm eps 1e-5, double-accumulated, weight and bias); bias-free MHA with a causal -inf mask inside the softmax, scale 1/sqrt(head);
cross-attention over K/V computed once from the stored text encoding (prep_cross_key_values, model.cpp:110-173) with an all-zero mask;
fc1 -> GELU (ggml's fp16 table) -> fc2; nine output heads; head i is fed BOS until step i + 1 (the delay pattern).
"""
from __future__ import annotations

import numpy as np
import torch

try:                                              # same ggml numerics (ggml-cpu.c:1816-1830, 7114-7163)
    from .kokoro_port import gelu_f16_lut, ggml_norm
except ImportError:                               # imported as a top-level module (oracle/ on sys.path)
    from kokoro_port import gelu_f16_lut, ggml_norm


def unpack_blocks(raw: np.ndarray, kind: str):
    """GGUF rows of Q8_0 / Q5_0 / Q4_0 blocks (ggml-common.h block_q8_0 / block_q5_0 / block_q4_0) -> (scales [N, nb] fp32, integer values [N, nb, 32] fp32)."""
    N = raw.shape[0]
    bs = {"Q8_0": 34, "Q5_0": 22, "Q4_0": 18}[kind]
    b = raw.reshape(N, -1, bs)
    d = b[:, :, :2].copy().view(np.float16)[:, :, 0].astype(np.float32)
    if kind == "Q8_0":
        q = b[:, :, 2:].copy().view(np.int8).astype(np.float32)
    else:
        qs = b[:, :, bs - 16:].astype(np.int32)
        lo, hi = qs & 0x0F, qs >> 4                                   # elements 0..15 and 16..31 of the block
        if kind == "Q5_0":
            qh = b[:, :, 2:6].copy().view(np.uint32)[:, :, 0].astype(np.int64)
            j = np.arange(16)
            lo = lo | (((qh[:, :, None] >> j) & 1) << 4)
            hi = hi | (((qh[:, :, None] >> (j + 16)) & 1) << 4)
            q = np.concatenate([lo, hi], axis=2).astype(np.float32) - 16.0
        else:
            q = np.concatenate([lo, hi], axis=2).astype(np.float32) - 8.0
    return d, q


def quant_mm(x: torch.Tensor, d: torch.Tensor, qv: torch.Tensor) -> torch.Tensor:
    """ggml_mul_mat with a block-quantised matrix (scales d [N, nb], integer values qv [N, nb, 32]): the activations (any leading shape, K last) are quantised to
    Q8_0 per 32 columns (vec_dot_type), integer dot products per block, scaled by d_w * d_x."""
    lead, K = x.shape[:-1], x.shape[-1]
    xb = x.reshape(-1, K // 32, 32)
    amax = xb.abs().amax(dim=2)
    idv = torch.where(amax > 0, np.float32(127.0) / amax, torch.zeros_like(amax))
    xq = torch.round(xb * idv[:, :, None])                              # AVX2 quantize_row_q8_0: _mm256_round_ps, nearest-even, scale 127 / amax
    dx = (amax / np.float32(127.0)).half().float()                       # the block scale is stored as fp16
    sumi = torch.einsum("nbk,Nbk->nNb", xq.double(), qv.double())        # exact integers
    y = (sumi.float() * (d[None, :, :] * dx[:, None, :])).sum(dim=2)
    return y.reshape(*lead, d.shape[0])


class ParlerPort:
    def __init__(self, gguf_path: str, threads: int = 8):
        import gguf
        torch.set_num_threads(threads)
        rd = gguf.GGUFReader(gguf_path)
        self.w = {}
        self.f16 = set()          # F16 matrices: ggml_mul_mat rounds the activations to fp16 before the product (ggml-cpu.c: vec_dot_type of F16 is F16)
        self.q = {}               # quantised matrices (Q8_0 / Q5_0 / Q4_0): name -> (block scales [N, nb] fp32, integer values [N, nb, 32] fp32)
        for t in rd.tensors:
            if t.name.startswith("decoder."):
                name = t.name[len("decoder."):]
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
        a = "parler-tts.decoder"
        self.layers = self.kv[f"{a}.num_hidden_layers"]; self.heads = self.kv[f"{a}.attention.head_count"]
        self.hidden = self.kv[f"{a}.hidden_size"]; self.hd = self.hidden // self.heads
        self.n_out = self.kv[f"{a}.output_heads"]; self.vocab = self.kv[f"{a}.out_vocab_size"]
        self.bos = self.kv["audio.bos_token_id"]; self.eos = self.kv["audio.eos_token_id"]
        enc = self.w["text_encoding"]
        self.ck = [self.mm(enc, f"layers.{l}.encoder_attn.k_proj.weight") for l in range(self.layers)]
        self.cv = [self.mm(enc, f"layers.{l}.encoder_attn.v_proj.weight") for l in range(self.layers)]
        self.reset()

    def mm(self, x, name):
        """ggml_mul_mat(weight, x): exact products of fp16-rounded activations with F16 weights, fp32 accumulation; plain fp32 for F32 weights."""
        if name in self.q:
            return quant_mm(x, *self.q[name])
        if name in self.f16:
            x = x.half().float()
        return x @ self.w[name].t()

    def set_text_encoding(self, enc):
        """parler_tts_model::prep_cross_key_values with a replacement conditional-prompt encoding [rows, hidden] (update_conditional_prompt, model.cpp:510-518)"""
        enc = torch.from_numpy(np.asarray(enc, np.float32))
        self.ck = [self.mm(enc, f"layers.{l}.encoder_attn.k_proj.weight") for l in range(self.layers)]
        self.cv = [self.mm(enc, f"layers.{l}.encoder_attn.v_proj.weight") for l in range(self.layers)]

    def reset(self):
        self.k = [None] * self.layers; self.v = [None] * self.layers
        self.pos = 0

    def ln(self, x, base):
        return ggml_norm(x, 1e-5) * self.w[base + ".weight"] + self.w[base + ".bias"]

    def attend(self, q, K, V, mask):
        n, T = q.shape[0], K.shape[0]
        qh = q.reshape(n, self.heads, self.hd); Kh = K.reshape(T, self.heads, self.hd); Vh = V.reshape(T, self.heads, self.hd)
        s = torch.einsum("nhd,thd->hnt", qh, Kh) * (1.0 / np.sqrt(np.float32(self.hd)))
        p = torch.softmax((s + mask[None]).double(), dim=-1).float()
        return torch.einsum("hnt,thd->nhd", p, Vh).reshape(n, self.hidden)

    def step(self, x) -> torch.Tensor:
        """x [n, hidden]: input embeddings (positions self.pos .. self.pos + n - 1 already added); returns logits [n_out, n, vocab]."""
        n = x.shape[0]
        for l in range(self.layers):
            b = f"layers.{l}"
            res = x
            cur = self.ln(x, b + ".self_attn_layer_norm")
            q = self.mm(cur, b + ".self_attn.q_proj.weight")
            k = self.mm(cur, b + ".self_attn.k_proj.weight")
            v = self.mm(cur, b + ".self_attn.v_proj.weight")
            self.k[l] = k if self.k[l] is None else torch.cat([self.k[l], k], 0)
            self.v[l] = v if self.v[l] is None else torch.cat([self.v[l], v], 0)
            T = self.k[l].shape[0]
            mask = torch.zeros(n, T)
            for i in range(n):
                mask[i, self.pos + i + 1:] = float("-inf")
            x = self.mm(self.attend(q, self.k[l], self.v[l], mask), b + ".self_attn.out_proj.weight") + res
            res = x
            cur = self.ln(x, b + ".encoder_attn_layer_norm")
            q = self.mm(cur, b + ".encoder_attn.q_proj.weight")
            x = self.mm(self.attend(q, self.ck[l], self.cv[l], torch.zeros(n, self.ck[l].shape[0])), b + ".encoder_attn.out_proj.weight") + res
            res = x
            cur = self.ln(x, b + ".final_layer_norm")
            cur = gelu_f16_lut(self.mm(cur, b + ".fc1.weight"))
            x = self.mm(cur, b + ".fc2.weight") + res
        x = self.ln(x, "layer_norm")
        self.pos += n
        return torch.stack([self.mm(x, f"lm_heads.{i}.weight.head") for i in range(self.n_out)])

    def greedy(self, prompt, steps: int, stop: bool = False, teacher=None):
        """Returns (tokens [steps', n_out], logits [steps', n_out, vocab]) like oracle/_ref/parler_ref.  stop: with the reference's stop rule
        (parler_context::eos_seen feeding + check_stopping, model.cpp:715-732,795-832) instead of a plain step cap.
        teacher [steps, n_out]: tokens fed back instead of the produced ones (teacher-forced comparison; the outputs are still the produced tokens)."""
        self.reset()
        tok = torch.from_numpy(np.asarray(prompt).astype(np.int64))
        x = self.w["embed_prompts"][tok] + self.w["positional_embed"][torch.arange(tok.numel())]
        self.step(x)
        toks, logits = [], []
        last = None
        seen = [False] * self.n_out                  # eos_seen: updated by check_stopping at the top of an iteration, i.e. AFTER the next batch was built
        max_gen = self.kv["parler-tts.decoder.max_generation"]
        for s in range(steps):                       # the audio batch built after decode number s has current_step == s
            ids = [(self.eos if seen[i] else int(last[i])) if s > i else self.bos for i in range(self.n_out)]
            if stop and s >= 1:                      # check_stopping before this decode
                if self.pos >= max_gen:
                    break
                seen = [seen[i] or int(last[i]) == self.eos for i in range(self.n_out)]
                if all(seen):
                    break
            x = None
            for i in range(self.n_out):              # embds[0][id0], then embds[i][id_i] + accumulated (parler_build_inp_embd)
                e = self.w[f"embed_tokens.{i}.weight"][ids[i]]
                x = e if x is None else e + x
            x = (x + self.w["positional_embed"][self.pos])[None, :]
            lg = self.step(x)[:, 0, :].numpy()
            last = lg.argmax(axis=1)                 # numpy argmax returns the first maximum, like sampler::max
            toks.append(last.astype(np.int32)); logits.append(lg)
            if teacher is not None:
                last = np.asarray(teacher[s])
        return np.stack(toks), np.stack(logits)
