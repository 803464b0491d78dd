"""ctypes binding of libb2tts.so -- the host-side mirror (for tests / bench) of the reference interface.

Names follow the reference (src/models/loaders.h:19-20, include/common.h:68-94, src/models/kokoro/model.h:430-468):
`runner_from_file(path)` returns a `KokoroRunner` whose `run(tokens)` is kokoro_runner::run (token-id entry, below the
phonemizer) and `run_batch` its batched form.  There is NO CPU fallback: a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2tts.so")
_lib = None


class B2TTSError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2TTSError(f"{LIB_PATH} is missing: run `python -m tts_cpp_b200.build` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.b2tts_last_error.restype = C.c_char_p
        _lib.b2tts_launch_count.restype = C.c_uint64
        _lib.b2tts_stream.restype = C.c_void_p
        _lib.b2tts_kokoro_voice_name.restype = C.c_char_p
        _lib.b2tts_kokoro_weight_bytes.restype = C.c_size_t
    return _lib


def _chk(rc: int):
    if rc != 0:
        raise B2TTSError(lib().b2tts_last_error().decode())


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Context:
    def __init__(self, device: int = 0):
        self.h = C.c_void_p()
        _chk(lib().b2tts_ctx_create(C.c_int(device), C.byref(self.h)))

    def launches(self) -> int:
        return int(lib().b2tts_launch_count(self.h))

    def gemm_launches(self) -> tuple[int, int]:
        """(tcgen05+TMA kernel launches, mma.sync fallback launches) since the context was created."""
        f = lib().b2tts_gemm_launches
        f.restype = C.c_uint64
        return int(f(self.h, 0)), int(f(self.h, 1))

    def stream(self) -> int:
        return int(lib().b2tts_stream(self.h) or 0)

    def close(self):
        if self.h:
            lib().b2tts_ctx_destroy(self.h)
            self.h = C.c_void_p()

    # ---- the patched ggml ops (ggml layout x[c][l]) -------------------------------------------------
    def conv_transpose_1d(self, kernel, x, stride, pad, out_pad=0, groups=1):
        k = np.ascontiguousarray(kernel, np.float32)      # numpy [Cin, Cout/g, K]
        xx = np.ascontiguousarray(x, np.float32)          # [Cin, L]
        cin, coutg, K = k.shape
        L = xx.shape[1]
        lout = (L - 1) * stride - 2 * pad + (K - 1) + out_pad + 1
        y = np.empty((coutg * groups, lout), np.float32)
        _chk(lib().b2tts_op_conv_transpose_1d(self.h, _fp(k), K, coutg, cin, _fp(xx), L, stride, pad, out_pad, groups, _fp(y)))
        return y

    def conv_1d(self, kernel, x, stride=1, pad=0, dil=1):
        k = np.ascontiguousarray(kernel, np.float32)      # numpy [Cout, Cin, K]
        xx = np.ascontiguousarray(x, np.float32)
        cout, cin, K = k.shape
        L = xx.shape[1]
        lout = (L + 2 * pad - dil * (K - 1) - 1) // stride + 1
        y = np.empty((cout, lout), np.float32)
        _chk(lib().b2tts_op_conv_1d(self.h, _fp(k), K, cin, cout, _fp(xx), L, stride, pad, dil, 1, _fp(y)))
        return y

    def vad_trim(self, utterances, sample_rate=44100.0, ms_per_frame=10, frame_threshold=20, normalized_energy_threshold=0.01, trailing_silent_frames=5,
                 early_cutoff_seconds_threshold=3, early_cutoff_energy_threshold=0.1):
        """apply_energy_voice_inactivity_detection (examples/cli/vad.cpp:11-68) for a batch -> (n_outputs[B] int64, [energies per utterance])"""
        utts = [np.ascontiguousarray(u, np.float32).ravel() for u in utterances]
        n = np.asarray([u.size for u in utts], np.int64)
        pcm = np.concatenate(utts) if utts else np.zeros(0, np.float32)
        spf = int(np.float32(ms_per_frame) * np.float32(sample_rate) / np.float32(1000.0)) if ms_per_frame > 0 else 0
        nf = (n // spf) if spf > 0 else np.zeros_like(n)
        en = np.zeros(max(int(nf.sum()), 1), np.float32)
        out = np.zeros(max(len(utts), 1), np.int64)
        L = lib()
        L.b2tts_op_vad_trim.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        _chk(L.b2tts_op_vad_trim(self.h, pcm.ctypes.data, n.ctypes.data, len(utts), sample_rate, ms_per_frame, frame_threshold, normalized_energy_threshold,
                                 trailing_silent_frames, early_cutoff_seconds_threshold, early_cutoff_energy_threshold, out.ctypes.data, en.ctypes.data))
        cuts = np.concatenate([[0], np.cumsum(nf)]).astype(np.int64)
        return out[:len(utts)], [en[cuts[b]:cuts[b + 1]].copy() for b in range(len(utts))]

    def cumsum(self, x):
        xx = np.ascontiguousarray(x, np.float32); y = np.empty_like(xx)
        _chk(lib().b2tts_op_cumsum(self.h, _fp(xx), xx.shape[-1], int(xx.size // xx.shape[-1]), _fp(y)))
        return y

    def _unary(self, fn, x, *args):
        xx = np.ascontiguousarray(x, np.float32); y = np.empty_like(xx)
        _chk(fn(self.h, _fp(xx), C.c_int64(xx.size), *args, _fp(y)))
        return y

    def mod(self, x, v):
        return self._unary(lib().b2tts_op_mod, x, C.c_float(v))

    def round(self, x):
        return self._unary(lib().b2tts_op_round, x)

    def reciprocal(self, x):
        return self._unary(lib().b2tts_op_reciprocal, x)

    def upscale_linear(self, x, factor):
        xx = np.ascontiguousarray(x, np.float32)
        y = np.empty(xx.shape[:-1] + (xx.shape[-1] * factor,), np.float32)
        _chk(lib().b2tts_op_upscale_linear(self.h, _fp(xx), xx.shape[-1], int(xx.size // xx.shape[-1]), factor, _fp(y)))
        return y

    def snake(self, alpha, x):
        a = np.ascontiguousarray(alpha, np.float32).ravel(); xx = np.ascontiguousarray(x, np.float32); y = np.empty_like(xx)
        _chk(lib().b2tts_op_snake(self.h, _fp(a), xx.shape[0], _fp(xx), xx.shape[1], _fp(y)))
        return y

    def stft(self, x, n_fft=20, hop=5):
        xx = np.ascontiguousarray(x, np.float32)
        fr = xx.shape[0] // hop + 1
        mag = np.empty((fr, n_fft // 2 + 1), np.float32); ph = np.empty_like(mag)
        _chk(lib().b2tts_op_stft(self.h, _fp(xx), xx.shape[0], n_fft, hop, _fp(mag), _fp(ph)))
        return mag, ph

    def istft(self, mag, ph, n_fft=20, hop=5):
        m = np.ascontiguousarray(mag, np.float32); p = np.ascontiguousarray(ph, np.float32)
        y = np.empty((m.shape[0] - 1) * hop, np.float32)
        _chk(lib().b2tts_op_istft(self.h, _fp(m), _fp(p), m.shape[0], n_fft, hop, _fp(y)))
        return y

    def uniform(self, count, skip=0):
        y = np.empty(count, np.float32)
        _chk(lib().b2tts_op_uniform(self.h, C.c_uint64(skip), C.c_int64(count), _fp(y)))
        return y

    def bilstm(self, w_ih, w_hh, b_ih, b_hh, x, lens):
        w_ih = np.ascontiguousarray(w_ih, np.float32); w_hh = np.ascontiguousarray(w_hh, np.float32)
        b_ih = np.ascontiguousarray(b_ih, np.float32); b_hh = np.ascontiguousarray(b_hh, np.float32)
        xx = np.ascontiguousarray(x, np.float32); ln = np.ascontiguousarray(lens, np.int32)
        B, Lmax, In = xx.shape
        H = w_hh.shape[-1]
        y = np.empty((B, Lmax, 2 * H), np.float32)
        _chk(lib().b2tts_op_bilstm(self.h, _fp(w_ih), _fp(w_hh), _fp(b_ih), _fp(b_hh), In, H, _fp(xx), B, Lmax,
                                   ln.ctypes.data_as(C.POINTER(C.c_int32)), _fp(y)))
        return y


class KokoroRunner:
    """kokoro_runner (src/models/kokoro/model.h:430-468) on a B200."""
    sampling_rate = 24000.0
    supports_voices = True

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle

    def list_voices(self):
        n = lib().b2tts_kokoro_n_voices(self.h)
        return [lib().b2tts_kokoro_voice_name(self.h, i).decode() for i in range(n)]

    def weight_bytes(self) -> int:
        return int(lib().b2tts_kokoro_weight_bytes(self.h))

    def run_batch(self, utterances, voice: str | None = None, noise_skip=None, copy: bool = True):
        """utterances: list of token-id lists (BOS/EOS included).  Returns (list of pcm arrays, list of duration arrays).
        copy=False returns views of the runner's pinned host buffers, valid until its next call -- the lifetime the reference
        gives tts_response::data (include/common.h), and what a server would hand to its encoder."""
        B = len(utterances)
        ntok = np.array([len(u) for u in utterances], np.int32)
        toks = np.ascontiguousarray(np.concatenate([np.asarray(u, np.uint32) for u in utterances]))
        pcm = (C.POINTER(C.c_float) * B)()
        ns = (C.c_int64 * B)()
        dur = C.POINTER(C.c_float)()
        skip = None
        if noise_skip is not None:
            skip = np.ascontiguousarray(noise_skip, np.uint64).ctypes.data_as(C.POINTER(C.c_uint64))
        _chk(lib().b2tts_kokoro_run_batch(self.h, B, toks.ctypes.data_as(C.POINTER(C.c_uint32)), ntok.ctypes.data_as(C.POINTER(C.c_int32)),
                                          voice.encode() if voice else None, skip, pcm, ns, C.byref(dur)))
        outs, durs, off = [], [], 0
        for b in range(B):
            a = np.ctypeslib.as_array(pcm[b], shape=(int(ns[b]),)) if ns[b] else np.zeros(0, np.float32)
            outs.append(a.copy() if copy else a)
            durs.append(np.ctypeslib.as_array(dur, shape=(int(ntok.sum()),))[off:off + int(ntok[b])].copy())
            off += int(ntok[b])
        return outs, durs

    def run(self, tokens, voice: str | None = None, noise_skip: int = 0):
        p, d = self.run_batch([tokens], voice, [noise_skip])
        return p[0], d[0]

    def run_batch_device(self, utterances, voice: str | None = None):
        """the forward with the PCM left on the device (b2tts_kokoro_run_batch_device) -> (device pointer, row stride in floats, n_samples per utterance);
        valid until the runner's next call.  For the multi-GPU gather (tts_cpp_b200/sharding.py)."""
        B = len(utterances)
        ntok = np.array([len(u) for u in utterances], np.int32)
        toks = np.ascontiguousarray(np.concatenate([np.asarray(u, np.uint32) for u in utterances]))
        ptr, stride, ns = C.POINTER(C.c_float)(), C.c_int64(), (C.c_int64 * B)()
        _chk(lib().b2tts_kokoro_run_batch_device(self.h, B, toks.ctypes.data_as(C.POINTER(C.c_uint32)), ntok.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 voice.encode() if voice else None, None, C.byref(ptr), C.byref(stride), ns))
        return C.cast(ptr, C.c_void_p).value, int(stride.value), [int(v) for v in ns]

    def timings(self):
        ms = (C.c_float * 3)()
        lib().b2tts_kokoro_last_timings(self.h, ms)
        return {"duration_ms": ms[0], "generation_ms": ms[1], "total_ms": ms[2]}

    # ---- test taps
    def set_taps(self, on=True):
        _chk(lib().b2tts_kokoro_set_taps(self.h, int(on)))

    def tap(self, name: str) -> np.ndarray:
        r, c, p = C.c_int64(), C.c_int64(), C.c_int64()
        _chk(lib().b2tts_kokoro_tap_info(self.h, name.encode(), C.byref(r), C.byref(c), C.byref(p)))
        a = np.empty((r.value, c.value), np.float32)
        _chk(lib().b2tts_kokoro_tap_read(self.h, name.encode(), _fp(a), C.c_size_t(a.size)))
        if p.value > 0 and p.value != c.value and r.value % p.value == 0:
            return a.reshape(r.value // p.value, p.value, c.value)   # [utterance][padded time][channels]
        return a                                                      # [utterance][samples]

    def override(self, name: str, arr=None):
        if arr is None:
            _chk(lib().b2tts_kokoro_override(self.h, name.encode(), None, C.c_size_t(0)))
        else:
            a = np.ascontiguousarray(arr, np.float32)
            _chk(lib().b2tts_kokoro_override(self.h, name.encode(), _fp(a), C.c_size_t(a.size)))

    def close(self):
        if self.h:
            lib().b2tts_kokoro_free(self.h)
            self.h = None


def runner_from_file(path: str, device: int = 0, ctx: Context | None = None) -> KokoroRunner:
    """runner_from_file (src/models/loaders.cpp:34-95) for general.architecture == "kokoro"."""
    ctx = ctx or Context(device)
    h = C.c_void_p()
    _chk(lib().b2tts_kokoro_load_gguf(ctx.h, path.encode(), C.byref(h)))
    return KokoroRunner(ctx, h)


class DacRunner:
    """dac_runner (reference src/decoder/dac_model.h:78-97): codebook indices -> PCM, batched over independent utterances."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle
        nh, up, cb = C.c_int(), C.c_int(), C.c_int()
        _chk(lib().b2tts_dac_info(self.h, C.byref(nh), C.byref(up), C.byref(cb)))
        self.n_heads, self.up_sampling_factor, self.codebook_size = nh.value, up.value, cb.value

    def run_batch(self, codes, copy: bool = True):
        """codes: list of [frames, n_heads] integer arrays (frame-major, the layout dac_runner::run takes).  Returns list of PCM arrays
        (copy=False: views of the runner's pinned host buffer, valid until its next call)."""
        B = len(codes)
        arrs = [np.ascontiguousarray(np.asarray(c, np.uint32).reshape(-1, self.n_heads)) for c in codes]
        frames = np.array([a.shape[0] for a in arrs], np.int32)
        ptrs = (C.POINTER(C.c_uint32) * B)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
        pcm = (C.POINTER(C.c_float) * B)()
        ns = (C.c_int64 * B)()
        _chk(lib().b2tts_dac_decode_batch(self.h, B, ptrs, frames.ctypes.data_as(C.POINTER(C.c_int32)), pcm, ns))
        outs = []
        for b in range(B):
            a = np.ctypeslib.as_array(pcm[b], shape=(int(ns[b]),))
            outs.append(a.copy() if copy else a)
        return outs

    def run(self, codes):
        return self.run_batch([codes])[0]

    def last_ms(self) -> float:
        lib().b2tts_dac_last_ms.restype = C.c_float
        return float(lib().b2tts_dac_last_ms(self.h))

    def close(self):
        if self.h:
            lib().b2tts_dac_free(self.h)
            self.h = None


def dac_runner_from_file(path: str, device: int = 0, ctx: Context | None = None) -> DacRunner:
    """The audio decoder half of the reference's Parler / Dia loaders (src/models/parler/loader.cpp:12-20): reads the "audio_encoder.*"
    tensors of a GGUF."""
    ctx = ctx or Context(device)
    h = C.c_void_p()
    _chk(lib().b2tts_dac_load_gguf(ctx.h, path.encode(), C.byref(h)))
    return DacRunner(ctx, h)


class SnacRunner:
    """snac_runner (reference src/decoder/snac_model.h:65-86): three code streams -> PCM, batched over independent utterances."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle
        up, cb = C.c_int(), C.c_int()
        _chk(lib().b2tts_snac_info(self.h, C.byref(up), C.byref(cb)))
        self.up_sampling_factor, self.codebook_size = up.value, cb.value

    def run_batch(self, codes, copy: bool = True):
        """codes: list of [coarse L/4, medium L/2, fine L] index arrays per utterance (what snac_runner::run takes)."""
        B = len(codes)
        arrs = [np.ascontiguousarray(np.concatenate([np.asarray(s, np.uint32) for s in c])) for c in codes]
        fine = np.array([len(c[2]) for c in codes], np.int32)
        ptrs = (C.POINTER(C.c_uint32) * B)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
        pcm = (C.POINTER(C.c_float) * B)()
        ns = (C.c_int64 * B)()
        _chk(lib().b2tts_snac_decode_batch(self.h, B, ptrs, fine.ctypes.data_as(C.POINTER(C.c_int32)), pcm, ns))
        outs = []
        for b in range(B):
            a = np.ctypeslib.as_array(pcm[b], shape=(int(ns[b]),))
            outs.append(a.copy() if copy else a)
        return outs

    def reset_noise(self):
        _chk(lib().b2tts_snac_reset_noise(self.h))

    def last_ms(self) -> float:
        lib().b2tts_snac_last_ms.restype = C.c_float
        return float(lib().b2tts_snac_last_ms(self.h))

    def close(self):
        if self.h:
            lib().b2tts_snac_free(self.h)
            self.h = None


def snac_runner_from_file(path: str, device: int = 0, ctx: Context | None = None) -> SnacRunner:
    ctx = ctx or Context(device)
    h = C.c_void_p()
    _chk(lib().b2tts_snac_load_gguf(ctx.h, path.encode(), C.byref(h)))
    return SnacRunner(ctx, h)


class Sampling(C.Structure):
    """b2tts_sampling: generation_configuration's sample / top_k / top_p / temperature / repetition_penalty (reference include/common.h:45-66) + a seed."""
    _fields_ = [("do_sample", C.c_int32), ("top_k", C.c_int32), ("top_p", C.c_float), ("temperature", C.c_float), ("repetition_penalty", C.c_float), ("seed", C.c_uint64)]

    def __init__(self, do_sample=True, top_k=50, top_p=1.0, temperature=1.0, repetition_penalty=1.0, seed=0):
        super().__init__(int(bool(do_sample)), int(top_k), float(top_p), float(temperature), float(repetition_penalty), int(seed))


def sample_uniform(seed: int, row: int, step: int) -> float:
    """the uniform head `row` draws at step `step` (b2tts_sample_uniform)"""
    lib().b2tts_sample_uniform.restype = C.c_float
    lib().b2tts_sample_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    return float(lib().b2tts_sample_uniform(seed, row, step))


def _prompt_args(prompts):
    B = len(prompts)
    arrs = [np.ascontiguousarray(np.asarray(p, np.uint32)) for p in prompts]
    npr = np.array([a.size for a in arrs], np.int32)
    ptrs = (C.POINTER(C.c_uint32) * B)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
    return B, arrs, npr, ptrs


class OrpheusRunner:
    """The token loop of orpheus_runner (reference src/models/orpheus/model.cpp:389-398) below the tokenizer, batched, greedy."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle
        v, l, hd = C.c_int(), C.c_int(), C.c_int()
        _chk(lib().b2tts_orpheus_info(self.h, C.byref(v), C.byref(l), C.byref(hd)))
        self.vocab_size, self.n_layers, self.hidden_size = v.value, l.value, hd.value

    def generate_greedy(self, prompts, n_steps: int, want_logits: bool = False):
        B = len(prompts)
        arrs = [np.ascontiguousarray(np.asarray(p, np.uint32)) for p in prompts]
        npr = np.array([a.size for a in arrs], np.int32)
        ptrs = (C.POINTER(C.c_uint32) * B)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
        toks = np.empty((B, n_steps), np.int32)
        logits = np.empty((B, n_steps, self.vocab_size), np.float32) if want_logits else None
        _chk(lib().b2tts_orpheus_generate_greedy(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(n_steps),
                                                 toks.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 logits.ctypes.data_as(C.POINTER(C.c_float)) if want_logits else None))
        return (toks, logits) if want_logits else toks

    def generate(self, prompts, n_steps: int, sampling: "Sampling | None" = None):
        """-> tokens [B][n_steps] under the reference sampler's settings (None: greedy)"""
        B, arrs, npr, ptrs = _prompt_args(prompts)
        toks = np.empty((B, n_steps), np.int32)
        _chk(lib().b2tts_orpheus_generate(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(n_steps), C.byref(sampling) if sampling is not None else None,
                                          toks.ctypes.data_as(C.POINTER(C.c_int32)), None))
        return toks

    def generate_until_stop(self, prompts, max_steps: int, sampling: "Sampling | None" = None):
        """-> (tokens [B][max_steps], n_generated [B]): the reference's loop with its stop condition (stopping token or max_steps)"""
        B, arrs, npr, ptrs = _prompt_args(prompts)
        toks = np.empty((B, max_steps), np.int32)
        ngen = np.empty(B, np.int32)
        _chk(lib().b2tts_orpheus_generate_until_stop(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(max_steps), C.byref(sampling) if sampling is not None else None,
                                                     toks.ctypes.data_as(C.POINTER(C.c_int32)), ngen.ctypes.data_as(C.POINTER(C.c_int32))))
        return toks, ngen

    def set_stopping_token(self, token_id: int):
        _chk(lib().b2tts_orpheus_set_stopping_token(self.h, int(token_id)))

    def step_weight_bytes(self) -> int:
        lib().b2tts_orpheus_step_weight_bytes.restype = C.c_size_t
        return int(lib().b2tts_orpheus_step_weight_bytes(self.h))

    def pdk_stats(self):
        """-> (launches of the persistent decode kernel, decode steps they covered)"""
        a, b = C.c_uint64(), C.c_uint64()
        lib().b2tts_orpheus_pdk_stats(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def last_ms(self) -> float:
        lib().b2tts_orpheus_last_ms.restype = C.c_float
        return float(lib().b2tts_orpheus_last_ms(self.h))

    def weight_bytes(self) -> int:
        lib().b2tts_orpheus_weight_bytes.restype = C.c_size_t
        return int(lib().b2tts_orpheus_weight_bytes(self.h))

    def close(self):
        if self.h:
            lib().b2tts_orpheus_free(self.h)
            self.h = None


def orpheus_runner_from_file(path: str, device: int = 0, ctx: Context | None = None) -> OrpheusRunner:
    ctx = ctx or Context(device)
    h = C.c_void_p()
    _chk(lib().b2tts_orpheus_load_gguf(ctx.h, path.encode(), C.byref(h)))
    return OrpheusRunner(ctx, h)


class ParlerRunner:
    """The decode loop of parler_tts_runner (reference src/models/parler/model.cpp:762-786) below the tokenizer, batched, greedy."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle
        nh, v, l, hd = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _chk(lib().b2tts_parler_info(self.h, C.byref(nh), C.byref(v), C.byref(l), C.byref(hd)))
        self.n_heads, self.out_vocab, self.n_layers, self.hidden_size = nh.value, v.value, l.value, hd.value

    def set_text_encoding(self, encoding):
        """update_conditional_prompt's second half (prep_cross_key_values with a new [rows, hidden] encoding, e.g. T5Runner.run's output)"""
        e = np.ascontiguousarray(encoding, np.float32)
        assert e.ndim == 2 and e.shape[1] == self.hidden_size, (e.shape, self.hidden_size)
        _chk(lib().b2tts_parler_set_text_encoding(self.h, e.ctypes.data_as(C.POINTER(C.c_float)), int(e.shape[0])))

    def generate_greedy(self, prompts, n_steps: int, want_logits: bool = False):
        """-> tokens [B][n_steps][n_heads] (and logits [B][n_steps][n_heads][out_vocab])"""
        B = len(prompts)
        arrs = [np.ascontiguousarray(np.asarray(p, np.uint32)) for p in prompts]
        npr = np.array([a.size for a in arrs], np.int32)
        ptrs = (C.POINTER(C.c_uint32) * B)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
        toks = np.empty((B, n_steps, self.n_heads), np.int32)
        logits = np.empty((B, n_steps, self.n_heads, self.out_vocab), np.float32) if want_logits else None
        _chk(lib().b2tts_parler_generate_greedy(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(n_steps),
                                                toks.ctypes.data_as(C.POINTER(C.c_int32)),
                                                logits.ctypes.data_as(C.POINTER(C.c_float)) if want_logits else None))
        return (toks, logits) if want_logits else toks

    def generate(self, prompts, n_steps: int, sampling: "Sampling | None" = None):
        """-> (tokens [B][n_steps][n_heads], n_generated [B]) under the reference sampler's settings (None: greedy) and its stop rule"""
        B, arrs, npr, ptrs = _prompt_args(prompts)
        toks = np.empty((B, n_steps, self.n_heads), np.int32)
        ngen = np.empty(B, np.int32)
        _chk(lib().b2tts_parler_generate(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(n_steps), C.byref(sampling) if sampling is not None else None,
                                         toks.ctypes.data_as(C.POINTER(C.c_int32)), None, ngen.ctypes.data_as(C.POINTER(C.c_int32))))
        return toks, ngen

    def generate_teacher_forced(self, prompts, teacher):
        """teacher [B][n_steps][n_heads]: the tokens fed back instead of the produced ones -> (produced tokens, logits [B][n_steps][n_heads][out_vocab])"""
        B, arrs, npr, ptrs = _prompt_args(prompts)
        teacher = np.ascontiguousarray(np.asarray(teacher, np.int32))
        n_steps = teacher.shape[1]
        toks = np.empty((B, n_steps, self.n_heads), np.int32)
        logits = np.empty((B, n_steps, self.n_heads, self.out_vocab), np.float32)
        _chk(lib().b2tts_parler_generate_teacher_forced(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(n_steps), teacher.ctypes.data_as(C.POINTER(C.c_int32)),
                                                        toks.ctypes.data_as(C.POINTER(C.c_int32)), logits.ctypes.data_as(C.POINTER(C.c_float))))
        return toks, logits

    def last_ms(self) -> float:
        lib().b2tts_parler_last_ms.restype = C.c_float
        return float(lib().b2tts_parler_last_ms(self.h))

    def weight_bytes(self) -> int:
        lib().b2tts_parler_weight_bytes.restype = C.c_size_t
        return int(lib().b2tts_parler_weight_bytes(self.h))

    def step_weight_bytes(self) -> int:
        """W_step: bytes of the weight tensors one decode step touches, each once"""
        lib().b2tts_parler_step_weight_bytes.restype = C.c_size_t
        return int(lib().b2tts_parler_step_weight_bytes(self.h))

    def pdk_stats(self):
        """-> (launches of the persistent decode kernel, decode steps they covered)"""
        a, b = C.c_uint64(), C.c_uint64()
        lib().b2tts_parler_pdk_stats(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def close(self):
        if self.h:
            lib().b2tts_parler_free(self.h)
            self.h = None


class T5Runner:
    """t5_runner::run (reference src/models/parler/t5/model.cpp:336-363) below the tokenizer, for a ragged batch of prompts."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle
        v = [C.c_int() for _ in range(6)]
        _chk(lib().b2tts_t5_info(self.h, *[C.byref(x) for x in v]))
        self.n_layers, self.hidden_size, self.output_size, self.vocab_size, self.context_length, self.eos_token_id = [x.value for x in v]

    def run(self, prompts):
        """prompts: token-id lists (EOS appended by the caller, like t5_runner::generate) -> list of [n_tokens, output_size] encodings"""
        B = len(prompts)
        arrs = [np.ascontiguousarray(np.asarray(p, np.uint32)) for p in prompts]
        npr = np.array([a.size for a in arrs], np.int32)
        ptrs = (C.POINTER(C.c_uint32) * B)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
        out = np.empty((int(npr.sum()), self.output_size), np.float32)
        _chk(lib().b2tts_t5_encode(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data_as(C.POINTER(C.c_float))))
        cuts = np.concatenate([[0], np.cumsum(npr)])
        return [out[cuts[b]:cuts[b + 1]].copy() for b in range(B)]

    def last_ms(self) -> float:
        lib().b2tts_t5_last_ms.restype = C.c_float
        return float(lib().b2tts_t5_last_ms(self.h))

    def close(self):
        if self.h:
            lib().b2tts_t5_free(self.h)
            self.h = C.c_void_p()


def t5_runner_from_file(path: str, device: int = 0, ctx: Context | None = None) -> T5Runner:
    """text_encoder_from_file (reference src/models/parler/t5/model.cpp:373-400) below the tokenizer"""
    ctx = ctx or Context(device)
    h = C.c_void_p()
    _chk(lib().b2tts_t5_load_gguf(ctx.h, path.encode(), C.byref(h)))
    return T5Runner(ctx, h)


def parler_runner_from_file(path: str, device: int = 0, ctx: Context | None = None) -> ParlerRunner:
    ctx = ctx or Context(device)
    h = C.c_void_p()
    _chk(lib().b2tts_parler_load_gguf(ctx.h, path.encode(), C.byref(h)))
    return ParlerRunner(ctx, h)


class DiaRunner:
    """dia_runner::decode inside generate_from_batch's loop (reference src/models/dia/model.cpp:705-737,806-864) below the tokenizer, batched, greedy."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle
        nh, v, c, mg = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _chk(lib().b2tts_dia_info(self.h, C.byref(nh), C.byref(v), C.byref(c), C.byref(mg)))
        self.n_heads, self.out_vocab, self.encoder_context, self.max_generation = nh.value, v.value, c.value, mg.value

    def generate_greedy(self, prompts, n_steps: int, want_logits: bool = False):
        """-> tokens [B][n_steps][n_heads], n_generated [B] (and the CFG-combined logits [B][n_steps][n_heads][out_vocab])"""
        B = len(prompts)
        arrs = [np.ascontiguousarray(np.asarray(p, np.uint32)) for p in prompts]
        npr = np.array([a.size for a in arrs], np.int32)
        ptrs = (C.POINTER(C.c_uint32) * B)(*[a.ctypes.data_as(C.POINTER(C.c_uint32)) for a in arrs])
        toks = np.empty((B, n_steps, self.n_heads), np.int32)
        ngen = np.empty(B, np.int32)
        logits = np.empty((B, n_steps, self.n_heads, self.out_vocab), np.float32) if want_logits else None
        _chk(lib().b2tts_dia_generate_greedy(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(n_steps),
                                             toks.ctypes.data_as(C.POINTER(C.c_int32)),
                                             logits.ctypes.data_as(C.POINTER(C.c_float)) if want_logits else None,
                                             ngen.ctypes.data_as(C.POINTER(C.c_int32))))
        return (toks, ngen, logits) if want_logits else (toks, ngen)

    def generate(self, prompts, n_steps: int, sampling: "Sampling | None" = None):
        """-> (tokens [B][n_steps][n_heads], n_generated [B]) under the reference sampler's settings (None: greedy)"""
        B, arrs, npr, ptrs = _prompt_args(prompts)
        toks = np.empty((B, n_steps, self.n_heads), np.int32)
        ngen = np.empty(B, np.int32)
        _chk(lib().b2tts_dia_generate(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(n_steps), C.byref(sampling) if sampling is not None else None,
                                      toks.ctypes.data_as(C.POINTER(C.c_int32)), None, ngen.ctypes.data_as(C.POINTER(C.c_int32))))
        return toks, ngen

    def generate_teacher_forced(self, prompts, teacher):
        """teacher [B][n_steps][n_heads]: the tokens fed back instead of the produced ones -> (produced tokens, CFG-combined logits)"""
        B, arrs, npr, ptrs = _prompt_args(prompts)
        teacher = np.ascontiguousarray(np.asarray(teacher, np.int32))
        n_steps = teacher.shape[1]
        toks = np.empty((B, n_steps, self.n_heads), np.int32)
        logits = np.empty((B, n_steps, self.n_heads, self.out_vocab), np.float32)
        _chk(lib().b2tts_dia_generate_teacher_forced(self.h, B, ptrs, npr.ctypes.data_as(C.POINTER(C.c_int32)), int(n_steps), teacher.ctypes.data_as(C.POINTER(C.c_int32)),
                                                     toks.ctypes.data_as(C.POINTER(C.c_int32)), logits.ctypes.data_as(C.POINTER(C.c_float))))
        return toks, logits

    def last_ms(self) -> float:
        lib().b2tts_dia_last_ms.restype = C.c_float
        return float(lib().b2tts_dia_last_ms(self.h))

    def pdk_stats(self):
        """-> (launches of the persistent decode kernel, decode steps they covered)"""
        a, b = C.c_uint64(), C.c_uint64()
        lib().b2tts_dia_pdk_stats(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def weight_bytes(self) -> int:
        lib().b2tts_dia_weight_bytes.restype = C.c_size_t
        return int(lib().b2tts_dia_weight_bytes(self.h))

    def close(self):
        if self.h:
            lib().b2tts_dia_free(self.h)
            self.h = None


def dia_runner_from_file(path: str, device: int = 0, ctx: Context | None = None) -> DiaRunner:
    ctx = ctx or Context(device)
    h = C.c_void_p()
    _chk(lib().b2tts_dia_load_gguf(ctx.h, path.encode(), C.byref(h)))
    return DiaRunner(ctx, h)
