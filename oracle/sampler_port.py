"""oracle/sampler_port.py -- TEST INFRASTRUCTURE.  CPU restatement of the reference sampler (reference src/sampler.cpp, src/sampler.h).

The checker for the on-device sampler (never imported by the product).  `nucleus()` restates the deterministic stages of sampler::sample in the order it
runs them (max -> [softmax over the vocabulary when top_p < 1] -> topk -> [softmax over the picks when top_p >= 1] -> topp; sampler.cpp:3-42,76-151,153-185)
and is pinned to oracle/_ref/sampler_ref stage by stage; `draw()` restates the final loop (sampler.cpp:47-69) for an EXPLICIT uniform per head -- the reference
seeds a fresh std::minstd_rand from std::random_device on every call, so its draws can only be pinned in distribution (tests/test_oracle_port.py does that
against a histogram of the reference's own draws).

Restated as they are, quirks included: the repetition penalty divides the logit of the last sampled token by penalty ** count (also inside the top-k
comparator and the max used to stabilise the softmax); the softmax denominators are sequential fp32 sums in pick order; after top-p trimming the probabilities
are NOT renormalised -- the uniform is scaled by min(prob_sum, top_p) instead; without top-k or top-p the reference's stop test reads picks[i] of an empty
vector (undefined behaviour): here the last vocabulary entry is the fallback."""
from __future__ import annotations

import numpy as np

f32 = np.float32


class SamplerPort:
    def __init__(self, heads: int, vocab: int, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0, repetition_penalty: float = 1.0):
        self.H, self.V = heads, vocab
        self.temperature, self.top_k, self.top_p, self.rp = f32(temperature), int(top_k), f32(top_p), f32(repetition_penalty)
        self.reset()

    def reset(self):
        self.last = np.full(self.H, -1, np.int64)
        self.counts = np.zeros(self.H, np.int64)

    def _eff(self, logits_row: np.ndarray, i: int) -> np.ndarray:
        """logits with the repetition penalty applied to the last sampled token: float(v / pow(penalty, count)) in double like the reference"""
        v = logits_row.astype(np.float32).copy()
        if self.rp != f32(1.0) and 0 <= self.last[i] < self.V:
            j = int(self.last[i])
            v[j] = f32(np.float64(v[j]) / np.power(np.float64(self.rp), np.float64(self.counts[i])))
        return v

    def _softmax(self, eff: np.ndarray, idx: np.ndarray, max_val: np.float32) -> np.ndarray:
        """probabilities of eff[idx] as sampler::softmax computes them: v/temperature, expf(v - max), sequential fp32 denominator"""
        v = eff[idx]
        if self.temperature != f32(1.0):
            v = (v / self.temperature).astype(np.float32)
            max_val = f32(max_val / self.temperature)
        e = np.exp((v - max_val).astype(np.float32)).astype(np.float32)
        s = f32(0.0)
        for x in e:
            s = f32(s + x)
        return (e / s).astype(np.float32)

    def nucleus(self, logits: np.ndarray):
        """-> per head (picks, probs of the picks, max_head_prob)"""
        out = []
        for i in range(self.H):
            eff = self._eff(logits[i], i)
            mx = int(np.argmax(eff))                        # first maximum
            allv = np.arange(self.V)
            if self.top_p < f32(1.0):
                probs_all = self._softmax(eff, allv, eff[mx])
                if 0 < self.top_k < self.V:
                    picks = np.argsort(-probs_all, kind="stable")[: self.top_k]
                else:
                    picks = np.argsort(-probs_all, kind="stable")
                ps, trim = f32(0.0), -1
                for n, j in enumerate(picks):
                    ps = f32(ps + probs_all[j])
                    if ps >= self.top_p:
                        trim = n + 1
                        break
                if trim > 0:
                    picks = picks[:trim]
                out.append((picks, probs_all[picks], f32(min(ps, self.top_p))))
            else:
                picks = np.argsort(-eff, kind="stable")[: self.top_k] if 0 < self.top_k < self.V else allv
                out.append((picks, self._softmax(eff, picks, eff[mx]), f32(1.0)))
        return out

    def draw(self, logits: np.ndarray, u: np.ndarray) -> np.ndarray:
        """tokens for explicit uniforms u[H] in [0, 1); updates the repetition state like sampler::sample"""
        toks = np.zeros(self.H, np.int64)
        for i, (picks, probs, mh) in enumerate(self.nucleus(logits)):
            a = f32(f32(u[i]) * mh) if self.top_p < f32(1.0) else f32(u[i])
            c, tok = f32(0.0), int(picks[-1])
            for n, j in enumerate(picks):
                c = f32(c + probs[n])
                if a <= c or n >= len(picks) - 1:
                    tok = int(j)
                    break
            if self.rp != f32(1.0):
                if self.last[i] != tok:
                    self.counts[i] = 0
                self.last[i] = tok
                self.counts[i] += 1
            toks[i] = tok
        return toks


def uniform_from_counter(seed: int, row: int, step: int) -> np.float32:
    """The uniform the device sampler derives for (seed, row, step) (tts_cpp_b200/csrc/sampler.cu: a splitmix64 finaliser over a mixed counter, 24 bits)."""
    M = (1 << 64) - 1
    x = (seed + 0x9E3779B97F4A7C15 * (row + 1) + 0xD1B54A32D192ED03 * (step + 1)) & M
    x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & M
    x ^= x >> 27; x = (x * 0x94D049BB133111EB) & M
    x ^= x >> 31
    return np.float32(x >> 40) * np.float32(1.0 / 16777216.0)
