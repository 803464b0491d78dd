"""Synthetic Kokoro-82M-shaped GGUF writer (no real weights exist offline).

Writes a GGUF file in the exact tensor-name / metadata schema the reference loader
consumes (reference py-gguf/tts_encoders/kokoro_gguf_encoder.py:146-487 defines the
schema; src/models/kokoro/model.cpp:413-773 routes the names, :841-930 reads the KV),
with random fp16-representable weights.  Used by the parity tests, bench.py and
__graft_entry__.smoke(); both the reference (oracle/_ref) and the CUDA path load
the same file, so parity is on identical weights.

dtype policy "f16" mirrors `quantize --quantized-type F16 --convert-non-quantized-to-f16`
(reference examples/quantize/quantize_impl.cpp:14-18,264-278): every tensor whose name
has no voice_tensors/bias/gamma/beta/alpha and does not end in embd/norm is stored F16 --
EXCEPT the ConvTranspose1d kernels (ups.N.weight, pool_weight), which stay F32 because the
reference's F16 ConvTranspose1d CPU kernel is mis-indexed (ggml-cpu.c:10091 vs :10189).
"""
from __future__ import annotations

import os
import numpy as np

STYLE = 128
HID = 512
N_VOICE_ROWS = 510


def _f16_ok(name: str) -> bool:
    if any(s in name for s in ("voice_tensors", "bias", "gamma", "beta", "alpha")):
        return False
    if name.endswith("embd") or name.endswith("norm"):
        return False
    if name.endswith("pool_weight") or (".ups." in name and name.endswith(".weight")):
        return False  # ConvTranspose1d kernels: keep F32 (reference F16 kernel bug)
    return True


class _Spec:
    """Collects (name, array) pairs; all randomness flows from one seeded generator."""

    def __init__(self, seed: int):
        self.rng = np.random.default_rng(seed)
        self.items: list[tuple[str, np.ndarray]] = []

    def rand(self, name, shape, scale=None):
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        s = (1.0 / np.sqrt(max(fan_in, 1))) if scale is None else scale
        a = (self.rng.standard_normal(shape).astype(np.float32) * np.float32(s))
        a = a.astype(np.float16).astype(np.float32)  # fp16-representable in every storage dtype
        self.items.append(("kokoro." + name, a))

    def const(self, name, shape, v):
        self.items.append(("kokoro." + name, np.full(shape, v, dtype=np.float32)))

    # -- composite blocks -------------------------------------------------
    def lstm(self, base, n_in, hid=256):
        for part in ("weights", "reverse_weights"):
            for g in range(4):
                self.rand(f"{base}.0.{part}.{2 * g}", (hid, n_in))
                self.rand(f"{base}.0.{part}.{2 * g + 1}", (hid, hid))
        for part in ("biases", "reverse_biases"):
            for i in range(8):
                self.rand(f"{base}.0.{part}.{i}", (hid,), 0.05)

    def adain_block(self, base, cin, cout, up=False):
        self.rand(f"{base}.conv1_weight", (cout, cin, 3))
        self.rand(f"{base}.conv1_bias", (cout,), 0.02)
        self.rand(f"{base}.conv2_weight", (cout, cout, 3))
        self.rand(f"{base}.conv2_bias", (cout,), 0.02)
        for nm, c in (("norm1", cin), ("norm2", cout)):
            for gb in ("gamma", "beta"):
                self.rand(f"{base}.{nm}_{gb}_weight", (c, STYLE))
                self.rand(f"{base}.{nm}_{gb}_bias", (c,), 0.02)
        if up:
            self.rand(f"{base}.pool_weight", (cin, 1, 3))
            self.rand(f"{base}.pool_bias", (cin,), 0.02)
        if cin != cout:
            self.rand(f"{base}.conv1x1_weight", (cout, cin, 1))

    def gen_resblock(self, base, ch, k):
        for i in range(3):
            for j in ("1", "2"):
                self.rand(f"{base}.{i}.gamma{j}_weight", (ch, STYLE))
                self.rand(f"{base}.{i}.gamma{j}_bias", (ch,), 0.02)
                self.rand(f"{base}.{i}.beta{j}_weight", (ch, STYLE))
                self.rand(f"{base}.{i}.beta{j}_bias", (ch,), 0.02)
                self.rand(f"{base}.{i}.convs{j}_weight", (ch, ch, k))
                self.rand(f"{base}.{i}.convs{j}_bias", (ch,), 0.02)
                alpha = (1.0 + 0.25 * self.rng.standard_normal((1, ch, 1))).clip(0.3, 2.0)
                self.items.append((f"kokoro.{base}.{i}.alpha{j}", alpha.astype(np.float16).astype(np.float32)))


def kokoro_tensors(seed: int = 0, dur_sigma: float = 0.01, dur_bias: float = -2.75, f0_bias: float = 120.0):
    """All Kokoro tensors as float32 numpy arrays (numpy shape = reversed ggml ne)."""
    s = _Spec(seed)
    # ALBERT (duration predictor front-end)
    s.rand("albert.token_embd", (178, 128), 0.5)
    s.rand("albert.position_embd", (512, 128), 0.1)
    s.rand("albert.token_type_embd", (128,), 0.1)
    s.const("albert.norm", (128,), 1.0)
    s.const("albert.norm_bias", (128,), 0.0)
    s.rand("albert.embd", (768, 128))
    s.rand("albert.embd_bias", (768,), 0.02)
    for nm in "qkvo":
        s.rand(f"albert.layer.0.{nm}", (768, 768))
        s.rand(f"albert.layer.0.{nm}_bias", (768,), 0.02)
    s.rand("albert.layer.0.ffn", (2048, 768))
    s.rand("albert.layer.0.ffn_bias", (2048,), 0.02)
    s.rand("albert.layer.0.ffn_out", (768, 2048))
    s.rand("albert.layer.0.ffn_out_bias", (768,), 0.02)
    for nm in ("attn_norm", "ffn_norm"):
        s.items.append((f"kokoro.albert.layer.0.{nm}", (1.0 + 0.1 * s.rng.standard_normal(768)).astype(np.float32)))
        s.rand(f"albert.layer.0.{nm}_bias", (768,), 0.05)
    # duration / prosody predictor
    s.rand("duration_predictor.encode", (HID, 768))
    s.rand("duration_predictor.encode_bias", (HID,), 0.02)
    for i in range(3):
        s.lstm(f"duration_predictor.layers.{2 * i}.lstm", HID + STYLE)
        for gb in ("gamma", "beta"):
            s.rand(f"duration_predictor.layers.{2 * i + 1}.{gb}_weight", (HID, STYLE))
            s.rand(f"duration_predictor.layers.{2 * i + 1}.{gb}_bias", (HID,), 0.02)
    s.lstm("duration_predictor.duration_lstm", HID + STYLE)
    s.lstm("duration_predictor.shared_lstm", HID + STYLE)
    s.rand("duration_predictor.duration_proj", (50, HID), dur_sigma)
    s.const("duration_predictor.duration_proj_bias", (50,), dur_bias)
    for br in ("f0", "n"):
        s.adain_block(f"duration_predictor.{br}_blocks.0", 512, 512)
        s.adain_block(f"duration_predictor.{br}_blocks.1", 512, 256, up=True)
        s.adain_block(f"duration_predictor.{br}_blocks.2", 256, 256)
        s.rand(f"duration_predictor.{br}_proj_kernel", (1, 256, 1))
        s.const(f"duration_predictor.{br}_proj_bias", (1,), f0_bias if br == "f0" else 0.0)
    # text encoder
    s.rand("text_encoder.embedding_weight", (178, 512), 0.5)
    for i in range(3):
        s.rand(f"text_encoder.layers.{i}.weight", (512, 512, 5))
        s.rand(f"text_encoder.layers.{i}.bias", (512,), 0.02)
        s.items.append((f"kokoro.text_encoder.layers.{i}.gamma", (1.0 + 0.1 * s.rng.standard_normal(512)).astype(np.float32)))
        s.rand(f"text_encoder.layers.{i}.beta", (512,), 0.05)
    s.lstm("text_encoder.lstm", 512)
    # decoder
    for br in ("f0", "n"):
        s.rand(f"decoder.{br}_conv_weight", (1, 1, 3))
        s.rand(f"decoder.{br}_conv_bias", (1,), 0.02)
    s.rand("decoder.asr_conv_weight", (64, 512, 1))
    s.rand("decoder.asr_conv_bias", (64,), 0.02)
    s.adain_block("decoder.encoder_block", 514, 1024)
    for i in range(3):
        s.adain_block(f"decoder.decoder_blocks.{i}", 1090, 1024)
    s.adain_block("decoder.decoder_blocks.3", 1090, 512, up=True)
    g = "decoder.generator"
    s.rand(f"{g}.m_source_weight", (1, 9))
    s.const(f"{g}.m_source_bias", (1,), 0.0)
    s.rand(f"{g}.ups.0.weight", (512, 256, 20))
    s.rand(f"{g}.ups.0.bias", (256,), 0.02)
    s.rand(f"{g}.ups.1.weight", (256, 128, 12))
    s.rand(f"{g}.ups.1.bias", (128,), 0.02)
    s.rand(f"{g}.noise_blocks.0.conv_weight", (256, 22, 12))
    s.rand(f"{g}.noise_blocks.0.conv_bias", (256,), 0.02)
    s.rand(f"{g}.noise_blocks.1.conv_weight", (128, 22, 1))
    s.rand(f"{g}.noise_blocks.1.conv_bias", (128,), 0.02)
    s.gen_resblock(f"{g}.noise_blocks.0.resblock", 256, 7)
    s.gen_resblock(f"{g}.noise_blocks.1.resblock", 128, 11)
    for i, (ch, k) in enumerate([(256, 3), (256, 7), (256, 11), (128, 3), (128, 7), (128, 11)]):
        s.gen_resblock(f"{g}.resblocks.{i}", ch, k)
    s.rand(f"{g}.conv_post_weight", (22, 128, 7), 0.01)
    s.const(f"{g}.conv_post_bias", (22,), 0.0)
    s.rand("voice_tensors.af_heart", (N_VOICE_ROWS, 256), 0.1)
    return s.items


def kokoro_metadata(ctx_len: int = 512):
    """(key, value) uint32 metadata the reference reads (model.cpp:841-930,264-298)."""
    a = "kokoro.duration_predictor.albert"
    kv = [(f"{a}.context_length", ctx_len), (f"{a}.layers", 1), (f"{a}.attn_heads", 12), (f"{a}.hidden_size", 768),
          (f"{a}.recurrence", 12), ("kokoro.duration_predictor.hidden_size", 512), ("kokoro.duration_predictor.layers", 3),
          ("kokoro.duration_predictor.f0_n_blocks", 3), ("kokoro.text_encoder.layers", 3)]
    G = "kokoro.decoder.generator"
    for k, v in dict(up_sampling_factor=600, kernels=3, upsamples=2, layers=4, padding=3, n_fft=20, hop=5).items():
        kv.append((f"{G}.{k}", v))
    for i, k in enumerate((7, 11)):
        for ii, d in enumerate((1, 3, 5)):
            kv.append((f"{G}.noise_blocks.{i}.res_block.{ii}.padding", (k * d - d) // 2))
            kv.append((f"{G}.noise_blocks.{i}.res_block.{ii}.dilation", d))
    kv += [(f"{G}.noise_blocks.0.stride", 6), (f"{G}.noise_blocks.0.padding", 3),
           (f"{G}.noise_blocks.1.stride", 1), (f"{G}.noise_blocks.1.padding", 0)]
    for i, k in enumerate((3, 7, 11, 3, 7, 11)):
        for ii, d in enumerate((1, 3, 5)):
            kv.append((f"{G}.res_blocks.{i}.{ii}.padding", (k * d - d) // 2))
            kv.append((f"{G}.res_blocks.{i}.{ii}.dilation", d))
    kv += [(f"{G}.up_convs.0.padding", 5), (f"{G}.up_convs.0.stride", 10),
           (f"{G}.up_convs.1.padding", 3), (f"{G}.up_convs.1.stride", 6)]
    return kv


def write_kokoro_gguf(path: str, seed: int = 0, dtype: str = "f16", ctx_len: int = 512, text_vocab: bool = False, **kw) -> dict:
    """Write the synthetic model; returns {"tensors": n, "params": n, "bytes": n}.
    text_vocab: a GGUF the reference's own TEXT front end can drive (examples/cli, generate("some text")): the built-in rule phonemizer (phonemizer.type 0) with one
    rule per letter (a -> a ...) and a token vocabulary of the printable ASCII characters, so that plain lower-case text phonemizes to itself and tokenizes 1:1."""
    import gguf

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    w = gguf.GGUFWriter(path, arch="kokoro")
    n_params = 0
    items = kokoro_tensors(seed=seed, **kw)
    for name, arr in items:
        n_params += arr.size
        if dtype == "f16" and _f16_ok(name):
            w.add_tensor(name, arr.astype(np.float16))
        else:
            w.add_tensor(name, arr.astype(np.float32))
    for k, v in kokoro_metadata(ctx_len):
        w.add_uint32(k, v)
    # the reference loader aborts without a (possibly trivial) rule phonemizer and tokenizer vocabulary
    w.add_uint32("phonemizer.type", 0)
    w.add_uint32("phonemizer.phoneme_type", 1)
    if text_vocab:
        letters = [chr(c) for c in range(ord("a"), ord("z") + 1)]
        w.add_array("phonemizer.graphemes", letters)
        w.add_array("phonemizer.rules.keys", letters)
        w.add_array("phonemizer.rules.phonemes", letters)
        w.add_array("phonemizer.dictionary.keys", ["a"])
        w.add_array("phonemizer.dictionary.values", ["a"])
        ascii_tokens = [chr(c) for c in range(32, 127)]
        w.add_array("tokenizer.ggml.tokens", [""] + ascii_tokens + [chr(0x100 + i) for i in range(177 - len(ascii_tokens))])
    else:
        w.add_array("phonemizer.graphemes", ["a", "b"])
        w.add_array("phonemizer.rules.keys", ["a"])
        w.add_array("phonemizer.rules.phonemes", ["a"])
        w.add_array("phonemizer.dictionary.keys", ["a"])
        w.add_array("phonemizer.dictionary.values", ["a"])
        w.add_array("tokenizer.ggml.tokens", [""] + [chr(0x100 + i) for i in range(177)])
    w.add_array("kokoro.voices", ["af_heart"])
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return {"tensors": len(items), "params": int(n_params), "bytes": os.path.getsize(path)}


def cached_gguf(dtype: str = "f16", ctx_len: int = 128, seed: int = 0, cache_dir: str | None = None, **kw) -> str:
    """Synthetic Kokoro GGUF cached on disk (deterministic in its arguments); safe under concurrent callers."""
    cache_dir = cache_dir or os.environ.get("B2TTS_CACHE", "/tmp/b2tts_cache")
    os.makedirs(cache_dir, exist_ok=True)
    tag = "_".join(f"{k}{v}" for k, v in sorted(kw.items()))
    path = os.path.join(cache_dir, f"kokoro_{dtype}_c{ctx_len}_s{seed}_{tag}.gguf")
    if not os.path.exists(path):
        tmp = f"{path}.{os.getpid()}.tmp"
        write_kokoro_gguf(tmp, seed=seed, dtype=dtype, ctx_len=ctx_len, **kw)      # (kw may carry text_vocab=True)
        os.replace(tmp, path)
    return path


# ------------------------------------------------------------------------------------------ DAC codec decoder
DAC_RATES = (8, 8, 4, 2)          # descript-audio-codec 44 kHz decoder: 1536 -> 768 -> 384 -> 192 -> 96 channels, 512 samples per frame


def dac_tensors(seed: int = 0, d_model: int = 1536, n_heads: int = 9, codebook: int = 1024, cb_dim: int = 8, latent: int = 1024,
                rates=DAC_RATES):
    """Synthetic weights in the reference's DAC schema (py-gguf/tts_encoders/dac_gguf_encoder.py:7-100, names after the
    "audio_encoder." prefix as dac_model::assign_weight sees them, src/decoder/dac_model.cpp:60-98).  Shapes are PyTorch order."""
    rng = np.random.default_rng(seed)
    items: list[tuple[str, np.ndarray]] = []

    def rand(name, shape, fan_in, scale=None):
        s = (1.0 / np.sqrt(max(fan_in, 1))) if scale is None else scale
        a = (rng.standard_normal(shape).astype(np.float32) * np.float32(s)).astype(np.float16).astype(np.float32)
        items.append(("audio_encoder." + name, a))

    def alpha(name, c):
        a = rng.uniform(0.5, 1.5, size=(1, c, 1)).astype(np.float32).astype(np.float16).astype(np.float32)
        items.append(("audio_encoder." + name, a))

    for i in range(n_heads):
        rand(f"quantizers.{i}.codebook.weight", (codebook, cb_dim), 1, 1.0)
        rand(f"quantizers.{i}.out_proj.weight", (latent, cb_dim, 1), cb_dim * n_heads)
        rand(f"quantizers.{i}.out_proj.bias", (latent,), 1, 0.02)
    rand("initial.weight", (d_model, latent, 7), latent * 7)
    rand("initial.bias", (d_model,), 1, 0.02)
    c = d_model
    for l, s in enumerate(rates, start=1):
        co = c // 2
        alpha(f"decoder_block.{l}.final.alpha", c)
        rand(f"decoder_block.{l}.final.weight", (c, co, 2 * s), c * 2)        # ConvTranspose1d kernel [Cin][Cout][K]; 2 taps reach an output
        rand(f"decoder_block.{l}.final.bias", (co,), 1, 0.02)
        for i in range(3):
            b = f"decoder_block.{l}.residual_unit.{i}.res"
            alpha(f"{b}.initial.alpha", co)
            rand(f"{b}.initial.weight", (co, co, 7), co * 7 * 4)                # small residual branch keeps the stack well-conditioned
            rand(f"{b}.initial.bias", (co,), 1, 0.02)
            alpha(f"{b}.final.alpha", co)
            rand(f"{b}.final.weight", (co, co, 1), co * 4)
            rand(f"{b}.final.bias", (co,), 1, 0.02)
        c = co
    alpha("final.alpha", c)
    rand("final.weight", (1, c, 7), c * 7 * 400)                                # keeps the tanh output un-saturated (std ~ 0.3)
    rand("final.bias", (1,), 1, 0.02)
    return items


def write_dac_gguf(path: str, seed: int = 0, max_frames: int = 128, **kw) -> dict:
    """Synthetic DAC decoder GGUF, all tensors F32 (the reference's quantizer leaves audio_encoder.* in F32 unless
    --convert-dac-to-f16 is given, examples/quantize/quantize_impl.cpp:44,265)."""
    import gguf

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    w = gguf.GGUFWriter(path, arch="dac")
    items = dac_tensors(seed=seed, **kw)
    n_params = 0
    for name, arr in items:
        n_params += arr.size
        w.add_tensor(name, arr.astype(np.float32))
    rates = kw.get("rates", DAC_RATES)
    for i, s in enumerate(rates):
        w.add_uint32(f"dac.dac_layer_stride_{i}", int(s))
        w.add_uint32(f"dac.dac_layer_padding_{i}", int((s + 1) // 2))           # DAC: padding = ceil(stride / 2)
    w.add_uint32("dac.up_sampling_factor", int(np.prod(rates)))
    w.add_uint32("output_heads", int(kw.get("n_heads", 9)))
    w.add_uint32("max_generation", int(max_frames))
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return {"tensors": len(items), "params": int(n_params), "bytes": os.path.getsize(path)}


def cached_dac_gguf(seed: int = 0, max_frames: int = 128, cache_dir: str | None = None) -> str:
    cache_dir = cache_dir or os.environ.get("B2TTS_CACHE", "/tmp/b2tts_cache")
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"dac_f32_m{max_frames}_s{seed}.gguf")
    if not os.path.exists(path):
        tmp = f"{path}.{os.getpid()}.tmp"
        write_dac_gguf(tmp, seed=seed, max_frames=max_frames)
        os.replace(tmp, path)
    return path


def synthetic_codes(batch: int, frames: int, n_heads: int = 9, codebook: int = 1024, seed0: int = 4321) -> list[np.ndarray]:
    """Utterance i: frames x n_heads codebook indices, frame-major (the layout dac_runner::run takes), from default_rng(seed0 + i)."""
    return [np.random.default_rng(seed0 + i).integers(0, codebook, size=(frames, n_heads)).astype(np.uint32) for i in range(batch)]


# ------------------------------------------------------------------------------------------ SNAC codec decoder
SNAC_RATES = (8, 8, 4, 2)         # snac_24khz decoder: 768 latent -> 1024 -> 512 -> 256 -> 128 -> 64 channels, 512 samples per fine frame


def snac_tensors(seed: int = 0, latent: int = 768, d_model: int = 1024, codebook: int = 4096, cb_dim: int = 8, rates=SNAC_RATES):
    """Synthetic weights in the reference's SNAC schema (py-gguf/tts_encoders/orpheus_gguf_encoder.py:89-142; names after the "snac."
    prefix as snac_model::assign_weight sees them, src/decoder/snac_model.cpp:52-84).  Shapes are PyTorch order."""
    rng = np.random.default_rng(seed)
    items: list[tuple[str, np.ndarray]] = []

    def rand(name, shape, fan_in, scale=None):
        s = (1.0 / np.sqrt(max(fan_in, 1))) if scale is None else scale
        a = (rng.standard_normal(shape).astype(np.float32) * np.float32(s)).astype(np.float16).astype(np.float32)
        items.append(("snac." + name, a))

    def alpha(name, c):
        a = rng.uniform(0.5, 1.5, size=(1, c, 1)).astype(np.float32).astype(np.float16).astype(np.float32)
        items.append(("snac." + name, a))

    for i in range(3):
        rand(f"quantizers.{i}.codebook.weight", (codebook, cb_dim), 1, 1.0)
        rand(f"quantizers.{i}.out_proj.weight", (latent, cb_dim, 1), cb_dim * 3)
        rand(f"quantizers.{i}.out_proj.bias", (latent,), 1, 0.02)
    rand("in.weight", (latent, 1, 7), 7)                       # depthwise
    rand("in.bias", (latent,), 1, 0.02)
    rand("up.weight", (d_model, latent, 1), latent)
    rand("up.bias", (d_model,), 1, 0.02)
    c = d_model
    for l, s in enumerate(rates):
        co = c // 2
        alpha(f"layers.{l}.alpha", c)
        rand(f"layers.{l}.weight", (c, co, 2 * s), c * 2)       # ConvTranspose1d [Cin][Cout][K]
        rand(f"layers.{l}.bias", (co,), 1, 0.02)
        rand(f"layers.{l}.noise_weight", (co, co, 1), co * 16)  # NoiseBlock linear (no bias)
        for i in range(3):
            b = f"layers.{l}.residual_unit.{i}.res"
            alpha(f"{b}.initial.alpha", co)
            rand(f"{b}.initial.weight", (co, 1, 7), 7 * 4)       # depthwise k7, dilation 3^i
            rand(f"{b}.initial.bias", (co,), 1, 0.02)
            alpha(f"{b}.final.alpha", co)
            rand(f"{b}.final.weight", (co, co, 1), co * 4)
            rand(f"{b}.final.bias", (co,), 1, 0.02)
        c = co
    alpha("alpha_out", c)
    rand("final.weight", (1, c, 7), c * 7 * 400)
    rand("final.bias", (1,), 1, 0.02)
    return items


def write_snac_gguf(path: str, seed: int = 0, max_frames: int = 64, **kw) -> dict:
    """Synthetic SNAC decoder GGUF, all tensors F32 (the reference's quantizer never converts snac.* : it only handles orpheus.* for Orpheus)."""
    import gguf

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    w = gguf.GGUFWriter(path, arch="snac")
    items = snac_tensors(seed=seed, **kw)
    n_params = 0
    for name, arr in items:
        n_params += arr.size
        w.add_tensor(name, arr.astype(np.float32))
    rates = kw.get("rates", SNAC_RATES)
    c = kw.get("d_model", 1024)
    for i, s in enumerate(rates):
        c //= 2
        w.add_uint32(f"snac.snac_layer_stride_{i}", int(s))
        w.add_uint32(f"snac.snac_layer_padding_{i}", int((s + 1) // 2))
        w.add_uint32(f"snac.snac_layer_grouping_{i}", int(c))      # depthwise residual units: groups == channels
    w.add_uint32("snac.audio_token_channels", 3)
    w.add_uint32("snac.up_sampling_factor", int(np.prod(rates)))
    w.add_uint32("snac.max_generation_size", int(max_frames))
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return {"tensors": len(items), "params": int(n_params), "bytes": os.path.getsize(path)}


def cached_snac_gguf(seed: int = 0, max_frames: int = 64, cache_dir: str | None = None) -> str:
    cache_dir = cache_dir or os.environ.get("B2TTS_CACHE", "/tmp/b2tts_cache")
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"snac_f32_m{max_frames}_s{seed}.gguf")
    if not os.path.exists(path):
        tmp = f"{path}.{os.getpid()}.tmp"
        write_snac_gguf(tmp, seed=seed, max_frames=max_frames)
        os.replace(tmp, path)
    return path


def synthetic_snac_codes(batch: int, fine_frames: int, codebook: int = 4096, seed0: int = 8765) -> list[list[np.ndarray]]:
    """Utterance i: three streams of L/4, L/2 and L indices (what snac_runner::run takes), from default_rng(seed0 + i); L % 4 == 0."""
    assert fine_frames % 4 == 0
    out = []
    for i in range(batch):
        r = np.random.default_rng(seed0 + i)
        out.append([r.integers(0, codebook, size=fine_frames // d).astype(np.uint32) for d in (4, 2, 1)])
    return out


# ------------------------------------------------------------------------------------------ Orpheus AR decoder
def orpheus_tensors(seed: int = 0, layers: int = 2, heads: int = 6, kv_heads: int = 2, head_dim: int = 64, ffn: int = 1024, vocab: int = 2048):
    """Synthetic llama-3-style weights in the reference's Orpheus schema (py-gguf/tts_encoders/orpheus_gguf_encoder.py:118-122, names after
    the "orpheus." prefix as orpheus_model::assign_weight sees them, src/models/orpheus/model.cpp:11-61).  heads == 3 * kv_heads: the
    reference hard-codes the GQA repeat of 3 (model.cpp:251)."""
    assert heads == 3 * kv_heads
    rng = np.random.default_rng(seed)
    hidden, kvh = heads * head_dim, kv_heads * head_dim
    items: list[tuple[str, np.ndarray]] = []

    def rand(name, shape, fan_in, scale=None):
        s = (1.0 / np.sqrt(max(fan_in, 1))) if scale is None else scale
        a = (rng.standard_normal(shape).astype(np.float32) * np.float32(s)).astype(np.float16).astype(np.float32)
        items.append(("orpheus." + name, a))

    def norm(name, c):
        items.append(("orpheus." + name, (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32).astype(np.float16).astype(np.float32)))

    rand("embed_tokens", (vocab, hidden), 1, 1.0)
    for l in range(layers):
        b = f"layers.{l}"
        norm(b + ".input_layernorm", hidden)
        rand(b + ".self_attn.q_proj", (hidden, hidden), hidden)
        rand(b + ".self_attn.k_proj", (kvh, hidden), hidden)
        rand(b + ".self_attn.v_proj", (kvh, hidden), hidden)
        rand(b + ".self_attn.o_proj", (hidden, hidden), hidden)
        norm(b + ".post_attention_layernorm", hidden)
        rand(b + ".mlp.gate_proj", (ffn, hidden), hidden)
        rand(b + ".mlp.up_proj", (ffn, hidden), hidden)
        rand(b + ".mlp.down_proj", (hidden, ffn), ffn)
    norm("norm", hidden)
    rand("lm_head", (vocab, hidden), hidden, 4.0 / np.sqrt(hidden))          # spread logits: greedy decoding far from ties
    # llama-3 rope frequency factors (ggml_rope_ext's `c` operand): 1 for the high frequencies, growing towards 8 for the low ones
    ff = np.ones(head_dim // 2, np.float32)
    ff[head_dim // 4:] = np.linspace(1.0, 8.0, head_dim // 2 - head_dim // 4).astype(np.float32)
    items.append(("orpheus.rope_frequencies", ff))
    return items


def write_orpheus_gguf(path: str, seed: int = 0, layers: int = 2, heads: int = 6, kv_heads: int = 2, head_dim: int = 64, ffn: int = 1024,
                       vocab: int = 2048, quant: str | None = None, f16: bool = False) -> dict:
    """Small synthetic Orpheus GGUF (all F32: the only dtype the reference supports for Orpheus, README.md:25), with the SNAC decoder
    tensors the reference's loader also needs."""
    import gguf

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    w = gguf.GGUFWriter(path, arch="orpheus")
    items = orpheus_tensors(seed, layers, heads, kv_heads, head_dim, ffn, vocab) + snac_tensors(seed=seed)
    n_params = 0
    for name, arr in items:
        n_params += arr.size
        is_matrix = name.startswith("orpheus.") and arr.ndim == 2 and arr.shape[1] % 32 == 0      # the decoder matrices, the embedding table and the head (our own
        if quant and is_matrix:                                                                # Q8_0 / F16 Orpheus writer: the reference's quantize tool refuses Orpheus)
            qt = getattr(gguf.GGMLQuantizationType, quant)
            w.add_tensor(name, gguf.quants.quantize(arr.astype(np.float32), qt), raw_dtype=qt)
        else:
            w.add_tensor(name, arr.astype(np.float16 if (f16 and is_matrix) else np.float32))
    for k, v in (("orpheus.vocab_size", vocab), ("orpheus.attn_heads", heads), ("orpheus.kv_attn_heads", kv_heads), ("orpheus.head_dim", head_dim),
                 ("orpheus.layers", layers), ("orpheus.hidden_size", heads * head_dim), ("orpheus.kv_hidden_size", kv_heads * head_dim),
                 ("orpheus.stopping_token_id", vocab - 1), ("tokenizer.ggml.eos_token_id", vocab - 2), ("tokenizer.ggml.bos_token_id", 1)):
        w.add_uint32(k, int(v))
    c = 1024
    for i, s in enumerate(SNAC_RATES):
        c //= 2
        w.add_uint32(f"snac.snac_layer_stride_{i}", int(s))
        w.add_uint32(f"snac.snac_layer_padding_{i}", int((s + 1) // 2))
        w.add_uint32(f"snac.snac_layer_grouping_{i}", int(c))
    w.add_uint32("snac.audio_token_channels", 3)
    w.add_uint32("snac.max_generation_size", 64)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return {"tensors": len(items), "params": int(n_params), "bytes": os.path.getsize(path)}


def cached_orpheus_gguf(seed: int = 0, cache_dir: str | None = None, **kw) -> str:
    cache_dir = cache_dir or os.environ.get("B2TTS_CACHE", "/tmp/b2tts_cache")
    os.makedirs(cache_dir, exist_ok=True)
    tag = "_".join(f"{k}{v}" for k, v in sorted(kw.items()))
    path = os.path.join(cache_dir, f"orpheus_f32_s{seed}_{tag}.gguf")
    if not os.path.exists(path):
        tmp = f"{path}.{os.getpid()}.tmp"
        write_orpheus_gguf(tmp, seed=seed, **kw)
        os.replace(tmp, path)
    return path


# ------------------------------------------------------------------------------------------ Parler-TTS decoder
def parler_tensors(seed: int = 0, layers: int = 8, heads: int = 32, head_dim: int = 8, ffn: int = 1024, out_vocab: int = 1088, n_heads: int = 9,
                   prompt_vocab: int = 512, n_enc: int = 12, ctx: int = 4096, eos_boost: float = 1.0):
    """Synthetic weights in the reference's Parler schema (py-gguf/tts_encoders/parler_tts_gguf_encoder.py:85-131; names after the
    "decoder." prefix as assign_to_decoder sees them, src/models/parler/model.cpp:3-28,271-318)."""
    rng = np.random.default_rng(seed)
    hidden = heads * head_dim
    items: list[tuple[str, np.ndarray]] = []

    def rand(name, shape, fan_in, scale=None):
        s = (1.0 / np.sqrt(max(fan_in, 1))) if scale is None else scale
        items.append(("decoder." + name, (rng.standard_normal(shape).astype(np.float32) * np.float32(s)).astype(np.float16).astype(np.float32)))

    def norm(base, c):
        items.append(("decoder." + base + ".weight", (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32).astype(np.float16).astype(np.float32)))
        items.append(("decoder." + base + ".bias", (0.05 * rng.standard_normal(c)).astype(np.float32).astype(np.float16).astype(np.float32)))

    rand("embed_prompts", (prompt_vocab, hidden), 1, 1.0)
    rand("text_encoding", (n_enc, hidden), 1, 1.0)
    rand("positional_embed", (ctx, hidden), 1, 0.5)
    for i in range(n_heads):
        rand(f"embed_tokens.{i}.weight", (out_vocab + 1, hidden), 1, 0.5)
    for l in range(layers):
        b = f"layers.{l}"
        norm(b + ".self_attn_layer_norm", hidden)
        for part in ("q_proj", "k_proj", "v_proj", "out_proj"):
            rand(f"{b}.self_attn.{part}.weight", (hidden, hidden), hidden)
        norm(b + ".encoder_attn_layer_norm", hidden)
        for part in ("q_proj", "k_proj", "v_proj", "out_proj"):
            rand(f"{b}.encoder_attn.{part}.weight", (hidden, hidden), hidden)
        norm(b + ".final_layer_norm", hidden)
        rand(f"{b}.fc1.weight", (ffn, hidden), hidden)
        rand(f"{b}.fc2.weight", (hidden, ffn), ffn)
    norm("layer_norm", hidden)
    for i in range(n_heads):
        rand(f"lm_heads.{i}.weight.head", (out_vocab, hidden), hidden, 4.0 / np.sqrt(hidden))
        if eos_boost != 1.0:            # make the EOS logit (row 1024) win now and then, so that a greedy run exercises eos_seen / check_stopping
            name, arr = items[-1]
            arr = arr.copy(); arr[1024] = (arr[1024] * np.float32(eos_boost)).astype(np.float16).astype(np.float32)
            items[-1] = (name, arr)
    return items


def parler_f16_tensor(name: str) -> bool:
    """Which tensors `quantize --quantized-type F16` turns into F16 with its default flags (reference examples/quantize/quantize_impl.cpp:51-67:
    parler_is_quanitizable): every decoder matrix and the codebook tables; norms, positional_embed, text_encoding, embed_prompts, the output heads
    and the cross-attention k / v projections stay F32, and so does the DAC."""
    if name.startswith("audio_encoder") or name.endswith(("norm.weight", "norm.bias", "text_encoding", "positional_embed", "weight.head", "embed_prompts",
                                                           "encoder_attn.k_proj.weight", "encoder_attn.v_proj.weight")):
        return False
    return True


def write_parler_gguf(path: str, seed: int = 0, layers: int = 8, heads: int = 32, head_dim: int = 8, ffn: int = 1024, n_enc: int = 12, f16: bool = False,
                      max_generation: int = 64, eos_boost: float = 1.0, quant: str | None = None) -> dict:
    """Small synthetic Parler-TTS GGUF (F32) with a matching small DAC decoder (the reference's loader needs both).  32 heads x 8 layers is
    the smallest shape the reference loads: prep_cross_key_values sizes its metadata pool from n_attn_heads * 2 * n_layers tensors but
    allocates a 4096-node graph in it (src/models/parler/model.cpp:117-129)."""
    import gguf

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    w = gguf.GGUFWriter(path, arch="parler-tts")
    dac_rates = (2, 2, 2, 2)
    items = parler_tensors(seed, layers, heads, head_dim, ffn, n_enc=n_enc, eos_boost=eos_boost) + dac_tensors(seed=seed, d_model=64, latent=32, rates=dac_rates)
    n_params = 0
    for name, arr in items:
        n_params += arr.size
        if quant and parler_f16_tensor(name) and arr.ndim == 2 and arr.shape[1] % 32 == 0:     # the tensors `quantize --quantized-type <quant>` converts (same rule as for F16)
            qt = getattr(gguf.GGMLQuantizationType, quant)
            w.add_tensor(name, gguf.quants.quantize(arr.astype(np.float32), qt), raw_dtype=qt)
        else:
            w.add_tensor(name, arr.astype(np.float16 if f16 and parler_f16_tensor(name) else np.float32))
    a = "parler-tts.decoder"
    for k, v in ((f"{a}.encode_length", n_enc), (f"{a}.hidden_size", heads * head_dim), (f"{a}.output_heads", 9), (f"{a}.context_length", 4096),
                 (f"{a}.attention.head_count", heads), (f"{a}.max_generation", max_generation), (f"{a}.out_vocab_size", 1088), (f"{a}.audio_vocab_size", 1024),
                 (f"{a}.num_hidden_layers", layers), ("audio.bos_token_id", 1025), ("audio.eos_token_id", 1024)):
        w.add_uint32(k, int(v))
    for i, s in enumerate(dac_rates):
        w.add_uint32(f"dac.dac_layer_stride_{i}", int(s))
        w.add_uint32(f"dac.dac_layer_padding_{i}", int((s + 1) // 2))
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return {"tensors": len(items), "params": int(n_params), "bytes": os.path.getsize(path)}


PARLER_MINI_SHAPE = dict(layers=24, heads=16, head_dim=64, ffn=4096, n_enc=32, max_generation=1024)     # parler-tts-mini-v1's decoder: hidden 1024, 24 layers, 9 codebooks


def cached_parler_gguf(seed: int = 0, cache_dir: str | None = None, f16: bool = False, **shape) -> str:
    cache_dir = cache_dir or os.environ.get("B2TTS_CACHE", "/tmp/b2tts_cache")
    os.makedirs(cache_dir, exist_ok=True)
    tag = "".join(f"_{k}{v}" for k, v in sorted(shape.items()) if v is not None)
    path = os.path.join(cache_dir, f"parler_{'f16' if f16 else 'f32'}_s{seed}{tag}.gguf")
    if not os.path.exists(path):
        tmp = f"{path}.{os.getpid()}.tmp"
        write_parler_gguf(tmp, seed=seed, f16=f16, **shape)
        os.replace(tmp, path)
    return path


# ------------------------------------------------------------------------------------------ Dia encoder + decoder
def dia_tensors(seed: int = 0, enc_layers: int = 2, dec_layers: int = 2, head_dim: int = 32, heads: int = 4, query_heads: int = 2, ffn: int = 256,
                vocab: int = 1028, n_heads: int = 9):
    """Synthetic weights in the reference's Dia schema (names exactly as dia_model::assign_weight splits them, src/models/dia/model.cpp:3-132).
    The encoder width is the reference's hard-coded 1024 (src/models/dia/model.h:69: no GGUF key overrides it); encoder heads * head_dim ==
    decoder heads * head_dim == decoder width (the reference reshapes attention outputs to decoder_hidden_size in both stacks)."""
    rng = np.random.default_rng(seed)
    EH, D, KVD = 1024, heads * head_dim, (heads // query_heads) * head_dim
    items: list[tuple[str, np.ndarray]] = []

    def rand(name, shape, fan_in, scale=None):
        s = (1.0 / np.sqrt(max(fan_in, 1))) if scale is None else scale
        items.append(("dia." + name, (rng.standard_normal(shape).astype(np.float32) * np.float32(s)).astype(np.float16).astype(np.float32)))

    def norm(name, c):
        items.append(("dia." + name, (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32).astype(np.float16).astype(np.float32)))

    rand("encoder.embedding", (256, EH), 1, 1.0)
    for l in range(enc_layers):
        b = f"encoder.layers.{l}"
        norm(b + ".pre_sa_norm", EH)
        rand(b + ".q_proj", (D, EH), EH, 2.0 / np.sqrt(EH)); rand(b + ".k_proj", (D, EH), EH, 2.0 / np.sqrt(EH)); rand(b + ".v_proj", (D, EH), EH)
        rand(b + ".o_proj", (EH, D), D)
        norm(b + ".post_sa_norm", EH)
        rand(b + ".gate", (ffn, EH), EH); rand(b + ".up", (ffn, EH), EH); rand(b + ".wo", (EH, ffn), ffn)
    norm("encoder.norm", EH)
    for i in range(n_heads):
        rand(f"decoder.embeddings.{i}", (vocab, D), 1, 0.5)
    for l in range(dec_layers):
        b = f"decoder.layers.{l}"
        norm(b + ".pre_sa_norm", D)
        rand(b + ".self_q_proj", (D, D), D, 2.0 / np.sqrt(D)); rand(b + ".self_k_proj", (KVD, D), D, 2.0 / np.sqrt(D)); rand(b + ".self_v_proj", (KVD, D), D)
        rand(b + ".self_o_proj", (D, D), D)
        norm(b + ".pre_ca_norm", D)
        rand(b + ".cross_q_proj", (D, D), D, 2.0 / np.sqrt(D)); rand(b + ".cross_k_proj", (D, EH), EH, 2.0 / np.sqrt(EH)); rand(b + ".cross_v_proj", (D, EH), EH)
        rand(b + ".cross_o_proj", (D, D), D)
        norm(b + ".pre_mlp_norm", D)
        rand(b + ".gate", (ffn, D), D); rand(b + ".up", (ffn, D), D); rand(b + ".wo", (D, ffn), ffn)
    norm("decoder.norm", D)
    for i in range(n_heads):
        rand(f"decoder.heads.{i}", (vocab, D), D, 4.0 / np.sqrt(D))
    return items


def dia_f16_tensor(name: str) -> bool:
    """Which tensors `quantize --quantized-type F16` turns into F16 for Dia with its default flags (reference examples/quantize/quantize_impl.cpp:42-49:
    dia_is_quantizable): everything but the DAC, the norms and the output heads."""
    return not (name.startswith("audio_encoder") or name.endswith("norm") or name.startswith("dia.decoder.heads"))


def write_dia_gguf(path: str, seed: int = 0, enc_layers: int = 2, dec_layers: int = 2, head_dim: int = 32, heads: int = 4, query_heads: int = 2,
                   ffn: int = 256, max_ctx: int = 32, f16: bool = False, quant: str | None = None) -> dict:
    """Small synthetic Dia GGUF (F32) with a matching small DAC decoder."""
    import gguf

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    w = gguf.GGUFWriter(path, arch="dia")
    dac_rates = (2, 2, 2, 2)
    items = dia_tensors(seed, enc_layers, dec_layers, head_dim, heads, query_heads, ffn) + dac_tensors(seed=seed, d_model=64, latent=32, rates=dac_rates)
    n_params = 0
    for name, arr in items:
        n_params += arr.size
        if quant and dia_f16_tensor(name) and arr.ndim == 2 and arr.shape[1] % 32 == 0:        # the tensors `quantize --quantized-type <quant>` converts
            qt = getattr(gguf.GGMLQuantizationType, quant)
            w.add_tensor(name, gguf.quants.quantize(arr.astype(np.float32), qt), raw_dtype=qt)
        else:
            w.add_tensor(name, arr.astype(np.float16 if f16 and dia_f16_tensor(name) else np.float32))
    for k, v in (("dia.decoder.output_heads", 9), ("dia.decoder.layers", dec_layers), ("dia.encoder.layers", enc_layers), ("dia.decoder.hidden_size", heads * head_dim),
                 ("dia.decoder.attn_heads", heads), ("dia.decoder.query_heads", query_heads), ("dia.encoder.attn_heads", heads), ("dia.attn_head_size", head_dim),
                 ("dia.eos_token_id", 1024), ("dia.bos_token_id", 1026), ("dia.pad_token_id", 1025), ("dia.encoder.max_context_length", max_ctx),
                 ("dia.decoder.output_vocab_size", 1028), ("dia.decoder.audio_vocab_size", 1024), ("dia.decoder.max_generation_size", 64), ("dia.max_delay", 15)):
        w.add_uint32(k, int(v))
    for i, s in enumerate(dac_rates):
        w.add_uint32(f"dac.dac_layer_stride_{i}", int(s))
        w.add_uint32(f"dac.dac_layer_padding_{i}", int((s + 1) // 2))
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return {"tensors": len(items), "params": int(n_params), "bytes": os.path.getsize(path)}


def cached_dia_gguf(seed: int = 0, cache_dir: str | None = None, f16: bool = False, quant: str | None = None, head_dim: int = 32) -> str:
    """head_dim 64: decoder width 256 -- the smallest shape the persistent decode kernel accepts (whole 256-column k-slices, head size 64 / 128)"""
    cache_dir = cache_dir or os.environ.get("B2TTS_CACHE", "/tmp/b2tts_cache")
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"dia_{quant.lower() if quant else ('f16' if f16 else 'f32')}{'' if head_dim == 32 else f'_hd{head_dim}'}_s{seed}.gguf")
    if not os.path.exists(path):
        tmp = f"{path}.{os.getpid()}.tmp"
        write_dia_gguf(tmp, seed=seed, f16=f16, quant=quant, head_dim=head_dim)
        os.replace(tmp, path)
    return path


def synthetic_prompts(batch: int, n_phonemes: int = 64, seed0: int = 1234) -> list[list[int]]:
    """Utterance i = BOS(0) + n_phonemes ids ~ U[1,177] from default_rng(seed0+i) + EOS(0)  (SURVEY 8d config 2)."""
    out = []
    for i in range(batch):
        r = np.random.default_rng(seed0 + i)
        out.append([0] + [int(v) for v in r.integers(1, 178, size=n_phonemes)] + [0])
    return out


if __name__ == "__main__":
    import sys
    print(write_kokoro_gguf(sys.argv[1], dtype=sys.argv[2] if len(sys.argv) > 2 else "f16",
                            ctx_len=int(sys.argv[3]) if len(sys.argv) > 3 else 512))


# ------------------------------------------------------------------------------------------ full-size synthetic models without a GGUF file
# BASELINE configs 4 and 5 name Dia-1.6B and Orpheus-3B.  Writing (and re-reading) a 3-15 GB synthetic GGUF on every fresh benchmark box costs minutes; the weights are
# random anyway, so bench.py hands them to the library tensor by tensor through the same C-ABI runner_from_file drives (b2tts_<model>_create / _assign_weight /
# _prepare, reference src/models/loaders.cpp:79-89) -- the layers share one set of random values (throughput does not depend on them; every layer still owns its HBM copy).
ORPHEUS_3B_SHAPE = dict(layers=28, heads=24, kv_heads=8, head_dim=128, ffn=8192, vocab=156940)       # reference src/models/orpheus/model.h:30-46
DIA_1B6_SHAPE = dict(enc_layers=12, dec_layers=18, head_dim=128, heads=16, query_heads=4, ffn=8192, enc_ffn=4096, vocab=1028, enc_ctx=1024)   # reference src/models/dia/model.h:63-85


def _q8_0_blocks(rng, n: int, scale: float) -> np.ndarray:
    """n weights ~ uniform int8 * scale as raw ggml Q8_0 blocks (fp16 scale + 32 int8: 34 bytes per 32 weights)"""
    nb = n // 32
    blk = np.zeros(nb, dtype=np.dtype([("d", "<f2"), ("q", "i1", (32,))]))
    blk["d"] = np.float16(scale / 64.0)
    blk["q"] = rng.integers(-127, 128, size=(nb, 32), dtype=np.int8)
    return blk.view(np.uint8).reshape(-1)


def build_orpheus_direct(ctx, dtype: str = "q8_0", seed: int = 0, **shape):
    """-> OrpheusRunner over random weights of the given shape; dtype "q8_0" (config 5), "f16" or "f32" (the reference's)."""
    import ctypes as C
    from .binding import OrpheusRunner, _chk, lib
    sh = dict(ORPHEUS_3B_SHAPE, **shape)
    L, heads, kvh, hd, F, V = sh["layers"], sh["heads"], sh["kv_heads"], sh["head_dim"], sh["ffn"], sh["vocab"]
    H, KV = heads * hd, kvh * hd
    kv = {"orpheus.vocab_size": V, "orpheus.attn_heads": heads, "orpheus.kv_attn_heads": kvh, "orpheus.head_dim": hd, "orpheus.layers": L, "orpheus.hidden_size": H,
          "orpheus.kv_hidden_size": KV, "orpheus.stopping_token_id": 128258}
    keys = (C.c_char_p * len(kv))(*[k.encode() for k in kv]); vals = (C.c_uint32 * len(kv))(*kv.values())
    h = C.c_void_p()
    _chk(lib().b2tts_orpheus_create(ctx.h, len(kv), keys, vals, C.byref(h)))
    rng = np.random.default_rng(seed)
    L_ = lib()
    L_.b2tts_orpheus_assign_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_size_t]

    def put(name, arr, ggml_type, shape_np):
        ne = (C.c_int64 * 4)(*(list(reversed(shape_np)) + [1] * (4 - len(shape_np))))
        a = np.ascontiguousarray(arr)
        _chk(L_.b2tts_orpheus_assign_weight(h, ("orpheus." + name).encode(), ggml_type, len(shape_np), ne, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def mat(n_out, n_in, scale=None):
        s = (1.0 / np.sqrt(n_in)) if scale is None else scale
        if dtype == "q8_0":
            return _q8_0_blocks(rng, n_out * n_in, 2.0 * s), 8
        w = (rng.standard_normal((n_out, n_in), dtype=np.float32) * np.float32(s))
        return (w.astype(np.float16), 1) if dtype == "f16" else (w, 0)

    layer = {"self_attn.q_proj": (H, H), "self_attn.k_proj": (KV, H), "self_attn.v_proj": (KV, H), "self_attn.o_proj": (H, H), "mlp.gate_proj": (F, H), "mlp.up_proj": (F, H), "mlp.down_proj": (H, F)}
    shared = {k: mat(*v) for k, v in layer.items()}
    ones = np.ones(H, np.float32)
    for l in range(L):
        for k, (a, t) in shared.items():
            put(f"layers.{l}.{k}", a, t, list(layer[k]))
        put(f"layers.{l}.input_layernorm", ones, 0, [H]); put(f"layers.{l}.post_attention_layernorm", ones, 0, [H])
    emb, t = mat(V, H, 1.0)
    put("embed_tokens", emb, t, [V, H]); del emb
    hd_, t = mat(V, H, 4.0 / np.sqrt(H))
    put("lm_head", hd_, t, [V, H]); del hd_
    put("norm", ones, 0, [H])
    ff = np.ones(hd // 2, np.float32); ff[hd // 4:] = np.linspace(1.0, 8.0, hd // 2 - hd // 4).astype(np.float32)
    put("rope_frequencies", ff, 0, [hd // 2])
    _chk(lib().b2tts_orpheus_prepare(h))
    return OrpheusRunner(ctx, h)


def build_dia_direct(ctx, dtype: str = "f16", seed: int = 0, **shape):
    """-> DiaRunner over random weights of the Dia-1.6B shape; dtype "f16" (config 4) or "f32".  (The output heads and norms stay F32, as the quantize tool leaves them.)"""
    import ctypes as C
    from .binding import DiaRunner, _chk, lib
    sh = dict(DIA_1B6_SHAPE, **shape)
    EL, DL, hd, heads, qh, F, EF, V, C_ = sh["enc_layers"], sh["dec_layers"], sh["head_dim"], sh["heads"], sh["query_heads"], sh["ffn"], sh["enc_ffn"], sh["vocab"], sh["enc_ctx"]
    EH, D, KVD = 1024, heads * hd, (heads // qh) * hd
    kv = {"dia.decoder.output_heads": 9, "dia.decoder.layers": DL, "dia.encoder.layers": EL, "dia.decoder.hidden_size": D, "dia.decoder.attn_heads": heads, "dia.decoder.query_heads": qh,
          "dia.encoder.attn_heads": heads, "dia.attn_head_size": hd, "dia.eos_token_id": 1024, "dia.bos_token_id": 1026, "dia.pad_token_id": 1025, "dia.encoder.max_context_length": C_,
          "dia.decoder.output_vocab_size": V, "dia.decoder.audio_vocab_size": 1024, "dia.decoder.max_generation_size": 3072, "dia.max_delay": 15}
    keys = (C.c_char_p * len(kv))(*[k.encode() for k in kv]); vals = (C.c_uint32 * len(kv))(*kv.values())
    h = C.c_void_p()
    _chk(lib().b2tts_dia_create(ctx.h, len(kv), keys, vals, C.byref(h)))
    rng = np.random.default_rng(seed)
    L_ = lib()
    L_.b2tts_dia_assign_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_size_t]

    def put(name, arr, ggml_type, shape_np):
        ne = (C.c_int64 * 4)(*(list(reversed(shape_np)) + [1] * (4 - len(shape_np))))
        a = np.ascontiguousarray(arr)
        _chk(L_.b2tts_dia_assign_weight(h, ("dia." + name).encode(), ggml_type, len(shape_np), ne, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def mat(n_out, n_in, scale=None, force_f32=False):
        w = rng.standard_normal((n_out, n_in), dtype=np.float32) * np.float32((1.0 / np.sqrt(n_in)) if scale is None else scale)
        return (w.astype(np.float16), 1) if (dtype == "f16" and not force_f32) else (w, 0)

    enc = {"q_proj": (D, EH), "k_proj": (D, EH), "v_proj": (D, EH), "o_proj": (EH, D), "gate": (EF, EH), "up": (EF, EH), "wo": (EH, EF)}
    dec = {"self_q_proj": (D, D), "self_k_proj": (KVD, D), "self_v_proj": (KVD, D), "self_o_proj": (D, D), "cross_q_proj": (D, D), "cross_k_proj": (D, EH), "cross_v_proj": (D, EH),
           "cross_o_proj": (D, D), "gate": (F, D), "up": (F, D), "wo": (D, F)}
    se = {k: mat(*v) for k, v in enc.items()}; sd = {k: mat(*v) for k, v in dec.items()}
    put("encoder.embedding", *mat(256, EH, 1.0), [256, EH])
    for l in range(EL):
        for k, (a, t) in se.items():
            put(f"encoder.layers.{l}.{k}", a, t, list(enc[k]))
        put(f"encoder.layers.{l}.pre_sa_norm", np.ones(EH, np.float32), 0, [EH]); put(f"encoder.layers.{l}.post_sa_norm", np.ones(EH, np.float32), 0, [EH])
    put("encoder.norm", np.ones(EH, np.float32), 0, [EH])
    tab = mat(V, D, 0.5)
    for i in range(9):
        put(f"decoder.embeddings.{i}", tab[0], tab[1], [V, D])
    for l in range(DL):
        for k, (a, t) in sd.items():
            put(f"decoder.layers.{l}.{k}", a, t, list(dec[k]))
        for nm in ("pre_sa_norm", "pre_ca_norm", "pre_mlp_norm"):
            put(f"decoder.layers.{l}.{nm}", np.ones(D, np.float32), 0, [D])
    put("decoder.norm", np.ones(D, np.float32), 0, [D])
    hw = mat(V, D, 4.0 / np.sqrt(D), force_f32=True)
    hw[0][1024:] = 0.0                                        # EOS / PAD / BOS never win the argmax: a benchmark generation runs its full length (check_stopping would end it at a random step)
    for i in range(9):
        put(f"decoder.heads.{i}", hw[0], hw[1], [V, D])
    _chk(lib().b2tts_dia_prepare(h))
    return DiaRunner(ctx, h)


# ---------------------------------------------------------------------------------------------------------------------------------------------------
# T5 conditional-prompt encoder: the text encoder Parler-TTS conditions on (reference src/models/parler/t5/model.cpp)
def t5_tensors(seed: int = 0, layers: int = 3, heads: int = 2, ffn: int = 192, vocab: int = 96, out_size: int = 256, down_proj: bool = True):
    """Synthetic weights in the reference's T5-encoder schema (T5_TENSOR_GGUF_LOOKUP, src/models/parler/t5/model.cpp:3-19; hyper-parameter keys
    :118-158).  The head size is fixed at 64 upstream (model.h:46), so hidden = 64 * heads."""
    rng = np.random.default_rng(seed + 7700)
    hidden = 64 * heads
    items: list[tuple[str, np.ndarray]] = []

    def rand(name, shape, scale):
        items.append((name, (rng.standard_normal(shape).astype(np.float32) * np.float32(scale)).astype(np.float16).astype(np.float32)))

    def norm(name, c):
        items.append((name, (1.0 + 0.1 * rng.standard_normal(c)).astype(np.float32).astype(np.float16).astype(np.float32)))

    rand("t5encoder.token_embd", (vocab, hidden), 1.0)
    for l in range(layers):
        b = f"t5encoder.enc.blk.{l}"        # parse_layer_count(name, 2) takes the layer index from the fourth dot-separated field (src/util.cpp)
        norm(b + ".attn_norm", hidden)
        # T5 attends WITHOUT 1/sqrt(d) (softmax scale 1.0, model.cpp:262): keep the scores O(1)
        rand(b + ".attn_q", (hidden, hidden), 0.6 / np.sqrt(hidden)); rand(b + ".attn_k", (hidden, hidden), 0.6 / np.sqrt(hidden))
        rand(b + ".attn_v", (hidden, hidden), 1.0 / np.sqrt(hidden)); rand(b + ".attn_o", (hidden, hidden), 1.0 / np.sqrt(hidden))
        if l == 0:
            rand(b + ".attn_rel_b", (32, heads), 0.5)         # [relative_attn_buckets][heads]: rows are fetched by bucket (ggml_get_rows, model.cpp:196)
        norm(b + ".ffn_norm", hidden)
        rand(b + ".ffn_gate", (ffn, hidden), 1.0 / np.sqrt(hidden)); rand(b + ".ffn_up", (ffn, hidden), 1.0 / np.sqrt(hidden))
        rand(b + ".ffn_down", (hidden, ffn), 1.0 / np.sqrt(ffn))
    norm("t5encoder.enc.final_layer_norm", hidden)
    if down_proj:
        rand("t5encoder.down_proj", (out_size, hidden), 1.0 / np.sqrt(hidden))
        rand("t5encoder.down_proj_bias", (out_size,), 0.1)
    return items


def write_t5_gguf(path: str, seed: int = 0, layers: int = 3, heads: int = 2, ffn: int = 192, vocab: int = 96, out_size: int = 256, down_proj: bool = True,
                  context_length: int = 64, f16: bool = False, quant: str | None = None) -> dict:
    """Small synthetic T5-encoder GGUF.  f16: the layer matrices as F16; quant ("Q8_0" / "Q5_0" / "Q4_0"): the layer matrices as ggml blocks (the reference's
    quantize tool does not handle text encoders, but its graph takes whatever type the file holds: ggml_mul_mat)."""
    import gguf

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    w = gguf.GGUFWriter(path, arch="t5encoder")
    items = t5_tensors(seed, layers, heads, ffn, vocab, out_size, down_proj)
    n_params = 0
    for name, arr in items:
        n_params += arr.size
        mat = arr.ndim == 2 and ".attn_rel_b" not in name and "token_embd" not in name and name != "t5encoder.down_proj"
        if quant and mat and arr.shape[1] % 32 == 0:
            qt = getattr(gguf.GGMLQuantizationType, quant)
            w.add_tensor(name, gguf.quants.quantize(arr.astype(np.float32), qt), raw_dtype=qt)
        else:
            w.add_tensor(name, arr.astype(np.float16 if f16 and mat else np.float32))
    for k, v in (("t5encoder.block_count", layers), ("t5encoder.embedding_length", 64 * heads), ("t5encoder.attention.head_count", heads),
                 ("t5encoder.context_length", context_length), ("t5encoder.vocab_size", vocab), ("t5encoder.output_size", out_size if down_proj else 64 * heads),
                 ("tokenizer.ggml.bos_token_id", 0), ("tokenizer.ggml.eos_token_id", 1)):
        w.add_uint32(k, int(v))
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return {"tensors": len(items), "params": int(n_params), "bytes": os.path.getsize(path)}


def cached_t5_gguf(seed: int = 0, cache_dir: str | None = None, f16: bool = False, **shape) -> str:
    cache_dir = cache_dir or os.environ.get("B2TTS_CACHE", "/tmp/b2tts_cache")
    os.makedirs(cache_dir, exist_ok=True)
    tag = "_".join(f"{k}{v}" for k, v in sorted(shape.items()) if v is not None)
    path = os.path.join(cache_dir, f"t5_{'f16' if f16 else 'f32'}_s{seed}{('_' + tag) if tag else ''}.gguf")
    if not os.path.exists(path):
        tmp = path + f".tmp{os.getpid()}"
        write_t5_gguf(tmp, seed=seed, f16=f16, **shape)
        os.replace(tmp, path)
    return path
