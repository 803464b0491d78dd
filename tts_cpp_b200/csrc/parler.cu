// parler.cu -- Parler-TTS autoregressive decode, first correct CUDA path.  See parler.h for what it replaces.
#include "parler.h"
#include "ar_kernels.cuh"
#include "pdk.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace b2 {

int Parler::assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {
    if (prepared) { set_error("parler: assign_weight after prepare"); return 1; }
    std::string nm(name);
    if (nm.rfind("decoder.", 0) == 0) nm = nm.substr(8);
    HostTensor t;
    if (host_tensor_from_gguf(t, name, type, n_dims, ne, data, nbytes, true)) return 1;
    host[nm] = std::move(t);
    return 0;
}

namespace {
struct PFwd : ArLaunch {
    Parler * m; bool fail = false;
    PFwd(Parler * m_, Ctx * c, cudaStream_t s) : m(m_) { ctx = c; st = s; }
    template <class T> T * al(size_t n) { T * p = (T *) m->arena.alloc(n * sizeof(T)); if (!p) fail = true; return p; }
    int ln(const float * x, const float * w, const float * b, int H, int R, float * y) {
        layernorm_kernel<<<cdiv(R, 8), 256, 0, st>>>(x, w, b, H, R, 1e-5f, y);
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
};
}  // namespace

int Parler::prepare() {
    if (prepared) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    const std::string a = "parler-tts.decoder.";
    auto kvreq = [&](const std::string & k, int & out) { auto it = kv.find(k); if (it == kv.end()) { set_error("the '%s' key must be specified in the GGUF file.", k.c_str()); return 1; } out = (int) it->second; return 0; };
    if (kvreq(a + "num_hidden_layers", n_layers) || kvreq(a + "attention.head_count", heads) || kvreq(a + "hidden_size", hidden) || kvreq(a + "output_heads", n_out) ||
        kvreq(a + "out_vocab_size", vocab) || kvreq(a + "encode_length", n_enc) || kvreq(a + "context_length", max_ctx) || kvreq(a + "max_generation", max_generation)) return 1;
    { auto it = kv.find("audio.bos_token_id"); if (it != kv.end()) bos = (int) it->second; it = kv.find("audio.eos_token_id"); if (it != kv.end()) eos = (int) it->second; }
    if (heads <= 0 || hidden % heads || hidden % 4 || (hidden / heads) % 4) { set_error("parler: inconsistent head configuration (head size must be a multiple of 4)"); return 1; }
    head_dim = hidden / heads;
    bool ok = true;
    auto find = [&](const std::string & n, int64_t expect) -> const HostTensor * {
        auto it = host.find(n);
        if (it == host.end()) { set_error("missing tensor decoder.%s", n.c_str()); ok = false; return nullptr; }
        if (expect && (int64_t) it->second.v.size() != expect) { set_error("tensor decoder.%s has %zu elements, expected %lld", n.c_str(), it->second.v.size(), (long long) expect); ok = false; return nullptr; }
        return &it->second;
    };
    auto dev = [&](const float * src, size_t n) -> float * {
        void * d = nullptr;
        if (cudaMalloc(&d, n * 4) != cudaSuccess) { cudaGetLastError(); set_error("parler: cudaMalloc of %zu bytes failed", n * 4); ok = false; return nullptr; }
        if (src) cudaMemcpy(d, src, n * 4, cudaMemcpyHostToDevice);
        dev_allocs.push_back(d); weight_bytes += n * 4;
        return (float *) d;
    };
    auto up = [&](const std::string & n, int64_t expect) -> float * { const HostTensor * t = find(n, expect); return t ? dev(t->v.data(), t->v.size()) : nullptr; };
    auto dev_mat = [&](const float * src, size_t n, bool f16) -> ArW {      // F16 tensors go to HBM as fp16 (their fp32 host copies are exact widenings)
        ArW w; w.f16 = f16;
        if (!f16) { w.p = dev(src, n); return w; }
        std::vector<__half> h(n);
        for (size_t i = 0; i < n; i++) h[i] = __float2half_rn(src[i]);
        void * d = nullptr;
        if (cudaMalloc(&d, n * 2) != cudaSuccess) { cudaGetLastError(); set_error("parler: cudaMalloc of %zu bytes failed", n * 2); ok = false; return w; }
        cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice);
        dev_allocs.push_back(d); weight_bytes += n * 2;
        w.p = d;
        return w;
    };
    auto upw = [&](const std::string & n, int64_t expect) -> ArW {
        const HostTensor * t = find(n, expect);
        if (!t) return ArW();
        if (!t->qtype) return dev_mat(t->v.data(), t->v.size(), t->f16);
        ArW w;                                                      // block-quantised matrix: value / scale / fifth-bit planes (ar_kernels.cuh)
        if (!upload_quant_planes(*t, w, dev_allocs, weight_bytes)) ok = false;
        return w;
    };

    embed_prompts = up("embed_prompts", 0);
    { const HostTensor * t = find("embed_prompts", 0); if (t) prompt_vocab = (int) (t->v.size() / (size_t) hidden); }
    pos_embed = up("positional_embed", 0);
    { const HostTensor * t = find("positional_embed", 0); if (t) max_ctx = std::min<int>(max_ctx, (int) (t->v.size() / (size_t) hidden)); }
    ln_w = up("layer_norm.weight", hidden); ln_b = up("layer_norm.bias", hidden);
    {   // the n_out codebook tables and output heads, each family in one buffer
        std::vector<float> tab, hw;
        bool heads_f16 = true, heads_any_f16 = false;
        for (int i = 0; i < n_out && ok; i++) {
            const HostTensor * t = find("embed_tokens." + std::to_string(i) + ".weight", 0);
            const HostTensor * h = find("lm_heads." + std::to_string(i) + ".weight.head", (int64_t) vocab * hidden);
            if (!t || !h) break;
            const int rows = (int) (t->v.size() / (size_t) hidden);
            if (i == 0) tab_rows = rows;
            if (rows != tab_rows || t->v.size() % (size_t) hidden) { set_error("parler: codebook table %d has %d rows, table 0 has %d", i, rows, tab_rows); ok = false; break; }
            tab.insert(tab.end(), t->v.begin(), t->v.end());
            hw.insert(hw.end(), h->v.begin(), h->v.end());
            heads_f16 = heads_f16 && h->f16; heads_any_f16 = heads_any_f16 || h->f16;
            if (h->qtype) { set_error("parler: block-quantised output heads (quantize --quantize-output-heads) are not supported"); ok = false; break; }
        }
        if (ok && heads_any_f16 != heads_f16) { set_error("parler: the output heads mix F16 and F32 tensors"); ok = false; }
        if (ok) { tables = dev(tab.data(), tab.size()); heads_w = dev_mat(hw.data(), hw.size(), heads_f16); }   // tables: ggml_get_rows widens F16 rows to fp32 exactly
        if (ok && !heads_f16) {   // F32 heads (the quantize tool leaves them F32): fp16 (hi, 2^11-scaled lo) planes for the persistent decode kernel's split tensor-core product
            std::vector<__half> hi(hw.size()), lo(hw.size());
            for (size_t i = 0; i < hw.size(); i++) { hi[i] = __float2half_rn(hw[i]); lo[i] = __float2half_rn((hw[i] - __half2float(hi[i])) * GM_LO_SCALE); }
            for (int pl = 0; pl < 2 && ok; pl++) {
                void * d = nullptr;
                if (cudaMalloc(&d, hw.size() * 2) != cudaSuccess) { cudaGetLastError(); set_error("parler: cudaMalloc of %zu bytes failed", hw.size() * 2); ok = false; break; }
                cudaMemcpy(d, pl ? lo.data() : hi.data(), hw.size() * 2, cudaMemcpyHostToDevice);
                dev_allocs.push_back(d); weight_bytes += hw.size() * 2;
                (pl ? heads_lo : heads_hi) = (__half *) d;
            }
        }
    }
    {
        const HostTensor * t = find("layers.0.fc1.weight", 0);
        if (t) { ffn = (int) t->shape[0]; if (ffn % 4) { set_error("parler: ffn size %d must be a multiple of 4", ffn); return 1; } }
    }
    const HostTensor * enc = find("text_encoding", (int64_t) n_enc * hidden);
    float * d_enc = enc ? dev(enc->v.data(), enc->v.size()) : nullptr;
    layers.resize((size_t) n_layers);
    PFwd Fw{this, ctx, ctx->stream};
    for (int l = 0; l < n_layers && ok; l++) {
        const std::string b = "layers." + std::to_string(l);
        ParlerLayer & L = layers[(size_t) l];
        const int64_t HH = (int64_t) hidden * hidden;
        L.ln1_w = up(b + ".self_attn_layer_norm.weight", hidden);    L.ln1_b = up(b + ".self_attn_layer_norm.bias", hidden);
        L.wq = upw(b + ".self_attn.q_proj.weight", HH); L.wk = upw(b + ".self_attn.k_proj.weight", HH); L.wv = upw(b + ".self_attn.v_proj.weight", HH); L.wo = upw(b + ".self_attn.out_proj.weight", HH);
        L.ln2_w = up(b + ".encoder_attn_layer_norm.weight", hidden); L.ln2_b = up(b + ".encoder_attn_layer_norm.bias", hidden);
        L.cq = upw(b + ".encoder_attn.q_proj.weight", HH); L.co = upw(b + ".encoder_attn.out_proj.weight", HH);
        L.ln3_w = up(b + ".final_layer_norm.weight", hidden);        L.ln3_b = up(b + ".final_layer_norm.bias", hidden);
        L.fc1 = upw(b + ".fc1.weight", (int64_t) ffn * hidden);      L.fc2 = upw(b + ".fc2.weight", (int64_t) hidden * ffn);
        // prep_cross_key_values (model.cpp:110-173): K and V of the stored text encoding, once per model
        L.wck = upw(b + ".encoder_attn.k_proj.weight", HH); L.wcv = upw(b + ".encoder_attn.v_proj.weight", HH);
        L.cross_k = dev(nullptr, (size_t) n_enc * hidden); L.cross_v = dev(nullptr, (size_t) n_enc * hidden);
        if (!ok) break;
        B2_CUDA(cudaDeviceSynchronize());    // the blocking uploads above went through the legacy stream; ctx->stream is non-blocking and does not wait for it by itself
        if (Fw.gemv(d_enc, hidden, L.wck, hidden, hidden, n_enc, nullptr, L.cross_k, hidden)) return 1;
        if (Fw.gemv(d_enc, hidden, L.wcv, hidden, hidden, n_enc, nullptr, L.cross_v, hidden)) return 1;
    }
    if (!ok) return 1;
    B2_CUDA(cudaDeviceSynchronize());      // the uploads above are blocking copies on the legacy stream, the kernels run on ctx->stream (non-blocking): order them once
#ifndef B2EMU
    B2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
#endif
    for (int i = 0; i < 2; i++) B2_CUDA(cudaEventCreate(&ev[i]));
    host.clear();
    prepared = true;
    return 0;
}

int Parler::set_text_encoding(const float * enc, int n_rows) {
    if (!prepared) { set_error("parler: model not prepared"); return 1; }
    if (!enc || n_rows <= 0) { set_error("parler: empty text encoding"); return 1; }
    B2_CUDA(cudaSetDevice(ctx->device));
    float * d_enc = nullptr;
    B2_CUDA(cudaMalloc(&d_enc, (size_t) n_rows * hidden * 4));
    B2_CUDA(cudaMemcpy(d_enc, enc, (size_t) n_rows * hidden * 4, cudaMemcpyHostToDevice));
    B2_CUDA(cudaDeviceSynchronize());        // legacy-stream upload before kernels on the non-blocking ctx->stream
    PFwd Fw{this, ctx, ctx->stream};
    int rc = 0;
    for (int l = 0; l < n_layers && !rc; l++) {
        ParlerLayer & L = layers[(size_t) l];
        float * nk = nullptr, * nv = nullptr;
        if (cudaMalloc(&nk, (size_t) n_rows * hidden * 4) != cudaSuccess || cudaMalloc(&nv, (size_t) n_rows * hidden * 4) != cudaSuccess) { cudaGetLastError(); set_error("parler: cudaMalloc failed for the cross K / V"); rc = 1; break; }
        dev_allocs.push_back(nk); dev_allocs.push_back(nv);          // the previous stores stay allocated until free_all (a few MB per call)
        rc = Fw.gemv(d_enc, hidden, L.wck, hidden, hidden, n_rows, nullptr, nk, hidden) || Fw.gemv(d_enc, hidden, L.wcv, hidden, hidden, n_rows, nullptr, nv, hidden);
        L.cross_k = nk; L.cross_v = nv;
    }
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error("parler: recomputing the cross K / V failed"); rc = 1; }
    cudaFree(d_enc);
    if (!rc) n_enc = n_rows;
    return rc;
}

void Parler::free_all() {
    for (void * p : dev_allocs) cudaFree(p);
    dev_allocs.clear();
    arena.release();
    for (int i = 0; i < 2; i++) if (ev[i]) cudaEventDestroy(ev[i]);
}

int Parler::generate(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const ArSampling * sampling, int32_t * out_tokens, float * out_logits,
                     int32_t * n_generated, const int32_t * teacher) {
    const ArSampling samp = sampling ? *sampling : ArSampling();
    if (!prepared) { set_error("parler: model not prepared"); return 1; }
    if (B <= 0 || n_steps <= 0) return 0;
    // more than 16 sequences, greedy / teacher-forced, F16 matrices: groups of 16, each inside the persistent decode kernel (rows are independent: the same tokens as
    // one large batch on the launch-per-op path, at the persistent kernel's step time)
    if (B > 16 && !samp.do_sample && heads_w.f16 + (heads_hi != nullptr) > 0 && !layers.empty() && layers[0].wq.f16 && !(getenv("B2TTS_AR_PDK") && getenv("B2TTS_AR_PDK")[0] == '0')) {
        const int NV_ = n_out * vocab;
        float total_ms = 0.f;
        for (int b0 = 0; b0 < B; b0 += 16) {
            const int nb = std::min(16, B - b0);
            if (generate(nb, prompts + b0, n_prompt + b0, n_steps, sampling, out_tokens + (size_t) b0 * n_steps * n_out, out_logits ? out_logits + (size_t) b0 * n_steps * NV_ : nullptr,
                         n_generated ? n_generated + b0 : nullptr, teacher ? teacher + (size_t) b0 * n_steps * n_out : nullptr)) return 1;
            total_ms += timing_ms;
        }
        timing_ms = total_ms;
        return 0;
    }
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int R0 = 0, Pmax = 0;
    for (int b = 0; b < B; b++) {
        if (n_prompt[b] <= 0) { set_error("parler: prompt %d is empty", b); return 1; }
        for (int i = 0; i < n_prompt[b]; i++) if (prompts[b][i] >= (uint32_t) prompt_vocab) { set_error("parler: prompt %d token %u >= prompt vocabulary %d", b, prompts[b][i], prompt_vocab); return 1; }
        R0 += n_prompt[b]; Pmax = std::max(Pmax, (int) n_prompt[b]);
    }
    const int Tmax = Pmax + n_steps, Rmax = std::max(R0, B), H = hidden, F = ffn, NV = n_out * vocab;
    if (Tmax > max_ctx) { set_error("parler: %d positions exceed the model's context of %d", Tmax, max_ctx); return 1; }
    // ---- the persistent decode kernel (pdk.cuh) takes the decode steps when the model and the request fit it: greedy / teacher-forced, <= 16 sequences, F16 decoder
    // matrices (the BASELINE config; F32 and block-quantised GGUFs stay on the launch-per-op path below), shapes in whole 256-column k-slices.  B2TTS_AR_PDK=0 turns it off.
    static const bool pdk_env = [] { const char * e = getenv("B2TTS_AR_PDK"); return !(e && e[0] == '0'); }();
    static const bool kv_f32 = [] { const char * e = getenv("B2TTS_KV"); return e && (e[0] == 'f' || e[0] == 'F') && e[1] == '3'; }();     // B2TTS_KV=f32: fp32 pages (default fp16)
#ifdef B2EMU
    const int pk_grid = [] { const char * e = getenv("B2TTS_PDK_GRID"); const int v = e ? atoi(e) : 3; return v > 0 ? v : 3; }();
#else
    const int pk_grid = [&] { const char * e = getenv("B2TTS_PDK_GRID"); const int v = e ? atoi(e) : 0; return v > 0 && v <= sm_count ? v : sm_count; }();
#endif
    const int pk_ak = 2048;                                         // activation chunk (columns): LayerNorm weight + bias of a 2 048-wide row fill one ring stage
    bool use_pdk = pdk_env && !samp.do_sample && B <= 16 && H % 256 == 0 && F % 256 == 0 && (head_dim == 8 || head_dim == 64 || head_dim == 128) && (heads_w.f16 || heads_hi) && !heads_w.qtype &&
                   pk_grid > 0 && (F <= pk_ak || H / 8 <= 3 * pk_grid) && (!out_logits || (size_t) n_steps * B * NV * 4 <= ((size_t) 1 << 30));
    for (const ParlerLayer & L : layers) for (const ArW * w : {&L.wq, &L.wk, &L.wv, &L.wo, &L.cq, &L.co, &L.fc1, &L.fc2}) use_pdk = use_pdk && w->f16 && !w->qtype;
    const int Tst = use_pdk ? Pmax : Tmax;                          // positions per sequence in the contiguous fp32 cache: the persistent path keeps only the prompt pass there
    const int pk_max_pages = cdiv(Tmax, PK_PAGE);
    int pk_pages = 0;
    for (int b = 0; b < B; b++) pk_pages += cdiv(n_prompt[b] + n_steps, PK_PAGE);
    const size_t pk_layer_bytes = (size_t) pk_pages * 2 * H * PK_PAGE * (kv_f32 ? 4 : 2);
    const size_t pk_need = use_pdk ? (size_t) n_layers * pk_layer_bytes + (size_t) B * pk_max_pages * 4 + (size_t) (8 * n_layers + 8) * sizeof(PkOp) + (size_t) R0 * 4 + 4096 + (size_t) PK_REP * 16 * ((size_t) 10 * H + 2 * F) + 4096 +
                                     (out_logits ? (size_t) n_steps * B * NV * 4 : 0) : 0;
    const size_t cache = (size_t) n_layers * B * Tst * H * 4;
    const size_t need = 2 * cache + (size_t) Rmax * ((size_t) 6 * H + F) * 4 + (size_t) B * NV * 4 + (size_t) n_steps * B * n_out * 4 + (size_t) B * n_out * 4 +
                        (size_t) Rmax * 32 + (size_t) B * 16 + (32 << 20) + (size_t) B * n_out * 12 + (size_t) B * 4 + (teacher ? (size_t) n_steps * B * n_out * 4 : 0) + (sampling_needs_scratch(samp, vocab) ? (size_t) B * NV * 4 : 0) + pk_need;
    if (arena.reserve(need)) return 1;
    PFwd Fw(this, ctx, st);
    float * Kc = Fw.al<float>((size_t) n_layers * B * Tst * H), * Vc = Fw.al<float>((size_t) n_layers * B * Tst * H);
    float * x = Fw.al<float>((size_t) Rmax * H), * xn = Fw.al<float>((size_t) Rmax * H), * q = Fw.al<float>((size_t) Rmax * H), * att = Fw.al<float>((size_t) Rmax * H);
    float * kbuf = Fw.al<float>((size_t) Rmax * H), * vbuf = Fw.al<float>((size_t) Rmax * H), * g = Fw.al<float>((size_t) Rmax * F);
    float * logits = Fw.al<float>((size_t) B * NV);
    int * row_tok = Fw.al<int>((size_t) Rmax), * row_pos = Fw.al<int>((size_t) Rmax), * row_base = Fw.al<int>((size_t) Rmax), * row_len = Fw.al<int>((size_t) Rmax), * row_dst = Fw.al<int>((size_t) Rmax);
    int * cross_base = Fw.al<int>((size_t) Rmax), * cross_len = Fw.al<int>((size_t) Rmax);
    int * d_np = Fw.al<int>((size_t) B), * ids = Fw.al<int>((size_t) B * n_out), * d_out = Fw.al<int>((size_t) n_steps * B * n_out), * d_step = Fw.al<int>(1);
    int * s_last = Fw.al<int>((size_t) B * n_out), * s_cnt = Fw.al<int>((size_t) B * n_out);
    int * seen = n_generated ? Fw.al<int>((size_t) B * n_out) : nullptr, * stopped = n_generated ? Fw.al<int>((size_t) B) : nullptr;
    int * d_teacher = teacher ? Fw.al<int>((size_t) n_steps * B * n_out) : nullptr;
    float * s_scratch = sampling_needs_scratch(samp, vocab) ? Fw.al<float>((size_t) B * NV) : nullptr;
    if (Fw.fail) return 1;
    B2_CUDA(cudaMemsetAsync(s_last, 0xff, (size_t) B * n_out * 4, st));     // sampler::reset: last_token_ids = -1, repetition_counts = 0
    B2_CUDA(cudaMemsetAsync(s_cnt, 0, (size_t) B * n_out * 4, st));
    if (n_generated) { B2_CUDA(cudaMemsetAsync(seen, 0, (size_t) B * n_out * 4, st)); B2_CUDA(cudaMemsetAsync(stopped, 0xff, (size_t) B * 4, st)); }

    std::vector<int> ht((size_t) R0), hp((size_t) R0), hb((size_t) R0), hl((size_t) R0), hd((size_t) R0), hcb((size_t) Rmax, 0), hcl((size_t) Rmax, n_enc), hnp((size_t) B);
    {
        int r = 0;
        for (int b = 0; b < B; b++) {
            hnp[(size_t) b] = n_prompt[b];
            for (int i = 0; i < n_prompt[b]; i++, r++) { ht[(size_t) r] = (int) prompts[b][i]; hp[(size_t) r] = i; hb[(size_t) r] = b * Tst; hl[(size_t) r] = i + 1; hd[(size_t) r] = b * Tst + i; }
        }
    }
    B2_CUDA(cudaEventRecord(ev[0], st));
    B2_CUDA(cudaMemcpyAsync(row_tok, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(row_pos, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(row_base, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(row_len, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(row_dst, hd.data(), hd.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(cross_base, hcb.data(), hcb.size() * 4, cudaMemcpyHostToDevice, st));   // every row attends to the whole stored encoding (all-zero cross mask)
    B2_CUDA(cudaMemcpyAsync(cross_len, hcl.data(), hcl.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(d_np, hnp.data(), hnp.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemsetAsync(d_step, 0, 4, st));
    std::vector<int> hteach;
    if (teacher) {                                                  // [B][n_steps][n_out] -> the device's [n_steps][B][n_out]
        hteach.resize((size_t) n_steps * B * n_out);
        for (int b = 0; b < B; b++) for (int s2 = 0; s2 < n_steps; s2++) for (int i = 0; i < n_out; i++) hteach[((size_t) s2 * B + b) * n_out + i] = teacher[((size_t) b * n_steps + s2) * n_out + i];
        B2_CUDA(cudaMemcpyAsync(d_teacher, hteach.data(), hteach.size() * 4, cudaMemcpyHostToDevice, st));
    }
    B2_CUDA(cudaStreamSynchronize(st));   // the host vectors above are stack-owned

    const float scale = 1.0f / sqrtf((float) head_dim);
    const int Tcap = std::max(Tmax, n_enc);

    // one pass over the layers for R rows whose inputs are already in x; leaves the result in x
    const bool fuse = ar_fuse_enabled();
    auto run_layers = [&](int R) -> int {
        for (int l = 0; l < n_layers; l++) {
            const ParlerLayer & L = layers[(size_t) l];
            float * Kl = Kc + (size_t) l * B * Tst * H, * Vl = Vc + (size_t) l * B * Tst * H;
            if (Fw.ln(x, L.ln1_w, L.ln1_b, H, R, xn)) return 1;
            if (fuse) {                                                                        // q, k, v in one launch; the k / v rows go straight to their cache slots
                const ArW * W3[3] = {&L.wq, &L.wk, &L.wv}; const int N3[3] = {H, H, H};
                const GemvOut o3[3] = {GemvOut{nullptr, q, nullptr, H, 0}, GemvOut{nullptr, Kl, row_dst, H, 0}, GemvOut{nullptr, Vl, row_dst, H, 0}};
                if (Fw.gemv_group(xn, H, H, R, W3, N3, o3, 3)) return 1;
            } else {
                if (Fw.gemv(xn, H, L.wq, H, H, R, nullptr, q, H)) return 1;
                if (Fw.gemv(xn, H, L.wk, H, H, R, nullptr, kbuf, H)) return 1;
                if (Fw.gemv(xn, H, L.wv, H, H, R, nullptr, vbuf, H)) return 1;
                store_kv_kernel<<<R, 256, 0, st>>>(kbuf, vbuf, row_dst, H, Kl, Vl); B2_LAUNCH_CHECK(ctx);
            }
            if (Fw.attend(q, Kl, Vl, row_base, row_len, R, heads, heads, head_dim, Tcap, scale, att)) return 1;
            if (Fw.gemv(att, H, L.wo, H, H, R, x, xn, H)) return 1;                            // xn = self-attention + residual(x)
            if (Fw.ln(xn, L.ln2_w, L.ln2_b, H, R, x)) return 1;
            if (Fw.gemv(x, H, L.cq, H, H, R, nullptr, q, H)) return 1;
            if (Fw.attend(q, L.cross_k, L.cross_v, cross_base, cross_len, R, heads, heads, head_dim, Tcap, scale, att)) return 1;
            if (Fw.gemv(att, H, L.co, H, H, R, xn, x, H)) return 1;                            // x = cross-attention + residual(xn)
            if (Fw.ln(x, L.ln3_w, L.ln3_b, H, R, xn)) return 1;
            if (fuse) {                                                                        // fc1 with ggml_gelu in its epilogue
                const ArW * W1[1] = {&L.fc1}; const int N1[1] = {F}; const GemvOut o1[1] = {GemvOut{nullptr, g, nullptr, F, 1}};
                if (Fw.gemv_group(xn, H, H, R, W1, N1, o1, 1)) return 1;
            } else {
                if (Fw.gemv(xn, H, L.fc1, H, F, R, nullptr, g, F)) return 1;
                { const size_t n = (size_t) R * F; gelu_f16lut_kernel<<<cdiv((int64_t) n, 256), 256, 0, st>>>(g, n); B2_LAUNCH_CHECK(ctx); }
            }
            if (Fw.gemv(g, F, L.fc2, F, H, R, x, x, H)) return 1;                              // x = mlp + residual(x), in place
        }
        return 0;
    };
    // the prompt pass: its logits are never read (generate_from_batch only samples after audio decodes)
    embed_pos_kernel<<<R0, 256, 0, st>>>(row_tok, row_pos, embed_prompts, pos_embed, H, x); B2_LAUNCH_CHECK(ctx);
    if (run_layers(R0)) return 1;
    const int exit_every = [] { const char * e = getenv("B2TTS_AR_EXIT_EVERY"); const int v = e ? atoi(e) : 32; return v > 0 ? v : 32; }();
    const bool track_stop = n_generated != nullptr;
    std::vector<int32_t> hflags((size_t) B);
    auto all_stopped = [&]() -> int {          // 1 all stopped, 0 not yet, -1 error
        if (cudaMemcpyAsync(hflags.data(), stopped, (size_t) B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { set_error("parler: reading the stop flags failed"); return -1; }
        for (int b = 0; b < B; b++) if (hflags[(size_t) b] < 0) return 0;
        return 1;
    };
    float * logits_all = nullptr;
    if (use_pdk) {
        // ---- paged KV cache: a page table per sequence over one pool per layer; pages are handed out in sequence order for the positions this call can reach
        // (ragged prompts take what they need, not Pmax + n_steps each)
        unsigned char * pool = (unsigned char *) arena.alloc((size_t) n_layers * pk_layer_bytes);
        int * page_table = Fw.al<int>((size_t) B * pk_max_pages), * row_seq = Fw.al<int>((size_t) R0);
        PkOp * d_ops = (PkOp *) arena.alloc((size_t) (8 * n_layers + 8) * sizeof(PkOp));
        unsigned * d_bar = (unsigned *) arena.alloc(256);
        if (out_logits) logits_all = Fw.al<float>((size_t) n_steps * B * NV);
        // PK_REP copies of every buffer all CTAs read at a phase start (pdk.cuh): the residual stream x / xn (fp32) and the fp16 hand-offs attention -> o-projection, GELU(fc1) -> fc2
        const size_t xrep = (size_t) 16 * H, grep = (size_t) 16 * F;
        float * px = Fw.al<float>(PK_REP * xrep), * pxn = Fw.al<float>(PK_REP * xrep);
        __half * att16 = Fw.al<__half>(PK_REP * xrep), * g16 = Fw.al<__half>(PK_REP * grep);
        if (!pool || !d_ops || !d_bar || Fw.fail) return 1;
        std::vector<int> hpt((size_t) B * pk_max_pages, 0), hrs((size_t) R0);
        { int next = 0, r = 0; for (int b = 0; b < B; b++) { const int np = cdiv(n_prompt[b] + n_steps, PK_PAGE); for (int i = 0; i < np; i++) hpt[(size_t) b * pk_max_pages + i] = next++; for (int i = 0; i < n_prompt[b]; i++) hrs[(size_t) r++] = b; } }
        std::vector<PkOp> ops;
        auto seg = [&](const __half * W, const __half * Wl, int N, int epi, float * Y, const float * res, int ldy, int kvsel) { PkSeg sg; memset(&sg, 0, sizeof sg); sg.W = W; sg.Wl = Wl; sg.N = N; sg.epi = epi; sg.Y = Y; sg.res = res; sg.ldy = ldy; sg.kv = kvsel; return sg; };
        auto gemv_op = [&](int layer, const float * X, int ldx, int K, const float * nw, const float * nb, std::initializer_list<PkSeg> segs) {
            PkOp op; memset(&op, 0, sizeof op);
            op.kind = PK_GEMV; op.layer = layer; op.X = X; op.ldx = ldx; op.K = K; op.norm = nw ? PKN_LAYER : PKN_NONE; op.nw = nw; op.nb = nb; op.eps = 1e-5f;
            int u = 0;
            for (const PkSeg & sg : segs) { op.seg[op.nseg] = sg; op.seg[op.nseg].unit0 = u; op.seg[op.nseg].n_units = cdiv(sg.N, 8); u += cdiv(sg.N, 8); op.nseg++; }
            op.n_units = u;
            ops.push_back(op);
        };
        auto attn_op = [&](int layer, const float * qv, __half * outv, const float * ckp, const float * cvp, int cross_len) {
            PkOp op; memset(&op, 0, sizeof op);
            op.kind = PK_ATTN; op.layer = layer; op.q = qv; op.out16 = outv; op.ck = ckp; op.cv = cvp; op.cross = ckp ? 1 : 0; op.cross_len = cross_len; op.scale = scale;
            ops.push_back(op);
        };
        { PkOp op; memset(&op, 0, sizeof op); op.kind = PK_ROWS; ops.push_back(op); }
        auto rep_in = [&](size_t stride) { ops.back().xrep = stride; };
        auto rep_out = [&](size_t stride) { ops.back().seg[0].yrep = stride; };
        for (int l = 0; l < n_layers; l++) {
            const ParlerLayer & L = layers[(size_t) l];
            gemv_op(l, px, H, H, L.ln1_w, L.ln1_b, {seg((const __half *) L.wq.p, nullptr, H, PKE_STORE, q, nullptr, H, 0), seg((const __half *) L.wk.p, nullptr, H, PKE_KV, nullptr, nullptr, H, 0),
                                                    seg((const __half *) L.wv.p, nullptr, H, PKE_KV, nullptr, nullptr, H, 1)});
            ops.back().kv_prefetch = 1; rep_in(xrep);
            attn_op(l, q, att16, nullptr, nullptr, 0); ops.back().orep = xrep;
            gemv_op(l, nullptr, H, H, nullptr, nullptr, {seg((const __half *) L.wo.p, nullptr, H, PKE_RES, pxn, px, H, 0)});           // xn = self-attention + residual(x)
            ops.back().X16 = att16; rep_in(xrep); rep_out(xrep);
            gemv_op(l, pxn, H, H, L.ln2_w, L.ln2_b, {seg((const __half *) L.cq.p, nullptr, H, PKE_STORE, q, nullptr, H, 0)}); rep_in(xrep);
            attn_op(l, q, att16, L.cross_k, L.cross_v, n_enc); ops.back().orep = xrep;
            gemv_op(l, nullptr, H, H, nullptr, nullptr, {seg((const __half *) L.co.p, nullptr, H, PKE_RES, px, pxn, H, 0)});           // x = cross-attention + residual(xn)
            ops.back().X16 = att16; rep_in(xrep); rep_out(xrep);
            gemv_op(l, px, H, H, L.ln3_w, L.ln3_b, {seg((const __half *) L.fc1.p, nullptr, F, PKE_GELU, nullptr, nullptr, F, 0)});
            ops.back().seg[0].Y16 = g16; rep_in(xrep); rep_out(grep);
            gemv_op(l, nullptr, F, F, nullptr, nullptr, {seg((const __half *) L.fc2.p, nullptr, H, PKE_RES, px, px, H, 0)});           // x = mlp + residual(x), element-wise in place (every copy)
            ops.back().X16 = g16; rep_in(grep); rep_out(xrep);
        }
        gemv_op(0, px, H, H, ln_w, ln_b, {seg(heads_w.f16 ? (const __half *) heads_w.p : heads_hi, heads_w.f16 ? nullptr : heads_lo, NV, PKE_LOGITS, logits, nullptr, NV, 0)}); rep_in(xrep);
        { PkOp op; memset(&op, 0, sizeof op); op.kind = PK_ARGMAX; ops.push_back(op); }
        B2_CUDA(cudaMemcpyAsync(page_table, hpt.data(), hpt.size() * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(row_seq, hrs.data(), hrs.size() * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_ops, ops.data(), ops.size() * sizeof(PkOp), cudaMemcpyHostToDevice, st));
        PkParams Pk; memset(&Pk, 0, sizeof Pk);
        Pk.ops = d_ops; Pk.n_ops = (int) ops.size(); Pk.R = B; Pk.H = H; Pk.heads = heads; Pk.kv_heads = heads; Pk.hd = head_dim; Pk.n_out = n_out; Pk.vocab = vocab;
        Pk.model = PKM_PARLER; Pk.ak = pk_ak; Pk.pos_off = 0; Pk.n_steps_total = n_steps;
        Pk.bar = d_bar; Pk.d_step = d_step;
        Pk.first_pos = d_np; Pk.d_out = d_out; Pk.d_teacher = d_teacher; Pk.bos = bos; Pk.eos = eos; Pk.max_gen = max_generation; Pk.seen = seen; Pk.stopped = stopped; Pk.ids = ids; Pk.row_pos = row_pos;
        Pk.tables = tables; Pk.tab_stride = (size_t) tab_rows * H; Pk.pos_embed = pos_embed; Pk.x0 = px; Pk.x0rep = xrep;
        Pk.kv_pool = pool; Pk.kv_layer_bytes = pk_layer_bytes; Pk.page_table = page_table; Pk.max_pages = pk_max_pages;
        Pk.logits = logits; Pk.logits_all = logits_all;
        // shared memory: activation rows of the widest phase (twice for split matrices) or the attention scratch of two half-CTA groups, the rest is the weight ring
        PkLaunch pkl;
        if (pk_configure(Pk, kv_f32, std::min(std::max(H, F), pk_ak), heads_w.f16 ? 0 : std::min(H, pk_ak), std::max(Tmax, n_enc), pkl)) { set_error("parler: the persistent decode kernel does not fit this shape (%d positions) in shared memory", Tmax); return 1; }
        pk_prof_begin(Pk, ops.size(), pk_grid, st);
        pk_kv_import(kv_f32, Kc, Vc, (size_t) B * Tst * H, row_dst, row_seq, row_pos, Pk, R0, n_layers, st);      // the prompt pass' K / V rows (fp32, contiguous) into the pages
        B2_LAUNCH_CHECK(ctx);
        for (int s0 = 0; s0 < n_steps; s0 += exit_every) {
            if (track_stop && s0 > 0) { const int a = all_stopped(); if (a < 0) return 1; if (a) break; }
            Pk.step_begin = s0; Pk.n_steps = std::min(exit_every, n_steps - s0);
            B2_CUDA(pk_launch(pkl, Pk, pk_grid, st));
            ctx->launches++; pdk_launches++; pdk_steps += (uint64_t) Pk.n_steps;
        }
        pk_prof_end(Pk, ops, pk_grid, st);
    }
    // one audio step; the step number is device-resident (d_step), so the launches are identical for every step
    auto run_step = [&]() -> int {
        delay_rows_kernel<<<cdiv(B, 128), 128, 0, st>>>(d_teacher ? d_teacher : d_out, d_np, B, n_out, d_step, bos, eos, max_generation, Tmax, seen, stopped, ids, row_pos, row_base, row_len, row_dst); B2_LAUNCH_CHECK(ctx);
        codebook_embed_kernel<<<B, 256, 0, st>>>(ids, n_out, tables, (size_t) tab_rows * H, pos_embed, row_pos, H, x); B2_LAUNCH_CHECK(ctx);
        if (run_layers(B)) return 1;
        if (Fw.ln(x, ln_w, ln_b, H, B, xn)) return 1;
        if (Fw.gemv(xn, H, heads_w, H, NV, B, nullptr, logits, NV)) return 1;                  // the n_out heads as one [n_out * vocab][hidden] matrix
        if (samp.do_sample) { if (sample_rows(ctx, make_sample_params(samp, logits, B * n_out, vocab, s_last, s_cnt, s_scratch, d_step, d_out))) return 1; }
        else { argmax_rows_kernel<<<B * n_out, 256, 0, st>>>(logits, vocab, d_out, d_step); B2_LAUNCH_CHECK(ctx); }
        step_advance_kernel<<<1, 32, 0, st>>>(d_step); B2_LAUNCH_CHECK(ctx);
        return 0;
    };
    // (launch-per-op path) every `exit_every` steps the per-sequence stop flags are read back (one small sync): when the reference's loop would have ended for EVERY
    // sequence of the batch the remaining steps are skipped
    // B2TTS_AR_GRAPH=1: capture one step into a CUDA graph and replay it (an audio step is ~15 launches per layer of microsecond kernels: launch-bound
    // otherwise).  Off by default until it has run on hardware; not used when the caller wants every step's logits (a host copy per step).
    const char * ge = getenv("B2TTS_AR_GRAPH");
    const bool use_graph = !(ge && ge[0] == '0') && !out_logits;      // on by default since it reproduced the reference's tokens on a B200 (round 2); B2TTS_AR_GRAPH=0 for A/B runs
    const uint64_t launches_before_step = ctx->launches;
    if (use_pdk) {
        // the steps already ran inside the persistent kernel; every step's logits sit in logits_all [step][b][NV]
        if (out_logits)
            for (int s = 0; s < n_steps; s++)
                for (int b = 0; b < B; b++)
                    B2_CUDA(cudaMemcpyAsync(out_logits + ((size_t) b * n_steps + s) * NV, logits_all + ((size_t) s * B + b) * NV, (size_t) NV * 4, cudaMemcpyDeviceToHost, st));
    } else if (use_graph && n_steps > 1) {
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        if (run_step()) return 1;                               // step 0 runs directly: every kernel instantiation has its attributes set before the capture
        B2_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = run_step();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (!rc) set_error("parler: stream capture failed: %s", cudaGetErrorString(ce)); return 1; }
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); set_error("parler: cudaGraphInstantiate failed"); return 1; }
        cudaError_t le = cudaSuccess;
        for (int s = 1; s < n_steps && le == cudaSuccess; s++) {
            le = cudaGraphLaunch(exec, st);
            if (track_stop && le == cudaSuccess && (s + 1) % exit_every == 0 && s + 1 < n_steps) { const int a = all_stopped(); if (a < 0) { le = cudaErrorUnknown; } else if (a) break; }
        }
        ctx->launches += (uint64_t) (n_steps - 2) * (uint64_t) ((ctx->launches - launches_before_step) / 2);   // direct step + captured step counted so far; the replays
        cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
        if (le != cudaSuccess) { set_error("parler: cudaGraphLaunch failed: %s", cudaGetErrorString(le)); return 1; }
    } else {
        for (int s = 0; s < n_steps; s++) {
            if (track_stop && s > 0 && s % exit_every == 0) { const int a = all_stopped(); if (a < 0) return 1; if (a) break; }
            if (run_step()) return 1;
            if (out_logits)
                for (int b = 0; b < B; b++)
                    B2_CUDA(cudaMemcpyAsync(out_logits + ((size_t) b * n_steps + s) * NV, logits + (size_t) b * NV, (size_t) NV * 4, cudaMemcpyDeviceToHost, st));
        }
    }
    B2_CUDA(cudaEventRecord(ev[1], st));
    std::vector<int32_t> tmp((size_t) n_steps * B * n_out), hstop((size_t) B, -1);
    B2_CUDA(cudaMemcpyAsync(tmp.data(), d_out, tmp.size() * 4, cudaMemcpyDeviceToHost, st));
    if (n_generated) B2_CUDA(cudaMemcpyAsync(hstop.data(), stopped, hstop.size() * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < B; b++) {
        const int n_gen = hstop[(size_t) b] >= 0 ? hstop[(size_t) b] : n_steps;
        if (n_generated) n_generated[b] = n_gen;
        for (int s = 0; s < n_steps; s++) {
            for (int i = 0; i < n_out; i++) out_tokens[((size_t) b * n_steps + s) * n_out + i] = s < n_gen ? tmp[((size_t) s * B + b) * n_out + i] : 0;
            if (out_logits && s >= n_gen) memset(out_logits + ((size_t) b * n_steps + s) * NV, 0, (size_t) NV * 4);
        }
    }
    cudaEventElapsedTime(&timing_ms, ev[0], ev[1]);
    return 0;
}

}  // namespace b2
