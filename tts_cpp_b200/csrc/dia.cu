// dia.cu -- Dia autoregressive decode (encoder pass + CFG-paired decoder loop), first correct CUDA path.  See dia.h for what it replaces.
#include "dia.h"
#include "ar_kernels.cuh"
#include "pdk.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace b2 {

int Dia::assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {
    if (prepared) { set_error("dia: assign_weight after prepare"); return 1; }
    std::string nm(name);
    if (nm.rfind("dia.", 0) == 0) nm = nm.substr(4);
    HostTensor t;
    if (host_tensor_from_gguf(t, name, type, n_dims, ne, data, nbytes, true)) return 1;
    host[nm] = std::move(t);
    return 0;
}

int Dia::prepare() {
    if (prepared) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    auto kvreq = [&](const char * k, int & out) { auto it = kv.find(k); if (it == kv.end()) { set_error("the '%s' key must be specified in the GGUF file.", k); return 1; } out = (int) it->second; return 0; };
    auto kvopt = [&](const char * k, int & out) { auto it = kv.find(k); if (it != kv.end()) out = (int) it->second; };
    if (kvreq("dia.encoder.layers", enc_layers) || kvreq("dia.decoder.layers", dec_layers) || kvreq("dia.decoder.attn_heads", heads) || kvreq("dia.decoder.query_heads", rep) ||
        kvreq("dia.encoder.attn_heads", enc_heads) || kvreq("dia.attn_head_size", head_dim) || kvreq("dia.encoder.max_context_length", enc_ctx) ||
        kvreq("dia.decoder.output_heads", n_out) || kvreq("dia.decoder.output_vocab_size", vocab) || kvreq("dia.decoder.max_generation_size", max_gen)) return 1;
    kvopt("dia.bos_token_id", bos); kvopt("dia.eos_token_id", eos); kvopt("dia.pad_token_id", pad); kvopt("dia.max_delay", max_delay);
    { auto it = kv.find("dia.cfg_scale#f32"); if (it != kv.end()) memcpy(&cfg, &it->second, 4); }
    if (heads <= 0 || rep <= 0 || heads % rep || head_dim % 4 || n_out > 9) { set_error("dia: inconsistent head configuration"); return 1; }
    hidden = heads * head_dim; kv_hidden = heads / rep * head_dim; enc_inner = enc_heads * head_dim;
    bool ok = true;
    auto find = [&](const std::string & n, int64_t expect) -> const HostTensor * {
        auto it = host.find(n);
        if (it == host.end()) { set_error("missing tensor dia.%s", n.c_str()); ok = false; return nullptr; }
        if (expect && (int64_t) it->second.v.size() != expect) { set_error("tensor dia.%s has %zu elements, expected %lld", n.c_str(), it->second.v.size(), (long long) expect); ok = false; return nullptr; }
        return &it->second;
    };
    auto dev = [&](const float * src, size_t n) -> float * {
        void * d = nullptr;
        if (cudaMalloc(&d, n * 4) != cudaSuccess) { cudaGetLastError(); set_error("dia: cudaMalloc of %zu bytes failed", n * 4); ok = false; return nullptr; }
        cudaMemcpy(d, src, n * 4, cudaMemcpyHostToDevice);
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
        if (cudaMalloc(&d, n * 2) != cudaSuccess) { cudaGetLastError(); set_error("dia: cudaMalloc of %zu bytes failed", n * 2); ok = false; return w; }
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

    {   // the encoder width is not in the metadata (the reference hard-codes 1024, model.h:68): take it from the embedding table
        const HostTensor * t = find("encoder.embedding", 0);
        if (!t || t->shape.size() != 2) { if (ok) set_error("dia: encoder.embedding must be 2-D"); return 1; }
        enc_vocab = (int) t->shape[0]; enc_hidden = (int) t->shape[1];
        const HostTensor * g = find("encoder.layers.0.gate", 0), * gd = find("decoder.layers.0.gate", 0);
        if (!g || !gd) return 1;
        enc_ffn = (int) g->shape[0]; ffn = (int) gd->shape[0];
        if (enc_hidden % 4 || hidden % 4 || enc_ffn % 4 || ffn % 4 || enc_inner % 4 || kv_hidden % 4) { set_error("dia: layer widths must be multiples of 4"); return 1; }
    }
    enc_embed = up("encoder.embedding", 0);
    enc_norm = up("encoder.norm", enc_hidden);
    dec_norm = up("decoder.norm", hidden);
    enc.resize((size_t) enc_layers); dec.resize((size_t) dec_layers);
    for (int l = 0; l < enc_layers && ok; l++) {
        const std::string b = "encoder.layers." + std::to_string(l);
        DiaEncLayer & L = enc[(size_t) l];
        L.pre_sa = up(b + ".pre_sa_norm", enc_hidden); L.post_sa = up(b + ".post_sa_norm", enc_hidden);
        L.wq = upw(b + ".q_proj", (int64_t) enc_inner * enc_hidden); L.wk = upw(b + ".k_proj", (int64_t) enc_inner * enc_hidden); L.wv = upw(b + ".v_proj", (int64_t) enc_inner * enc_hidden);
        L.wo = upw(b + ".o_proj", (int64_t) enc_hidden * enc_inner);
        L.gate = upw(b + ".gate", (int64_t) enc_ffn * enc_hidden); L.up = upw(b + ".up", (int64_t) enc_ffn * enc_hidden); L.down = upw(b + ".wo", (int64_t) enc_hidden * enc_ffn);
    }
    for (int l = 0; l < dec_layers && ok; l++) {
        const std::string b = "decoder.layers." + std::to_string(l);
        DiaDecLayer & L = dec[(size_t) l];
        const int64_t DD = (int64_t) hidden * hidden;
        L.pre_sa = up(b + ".pre_sa_norm", hidden); L.pre_ca = up(b + ".pre_ca_norm", hidden); L.pre_mlp = up(b + ".pre_mlp_norm", hidden);
        L.sq = upw(b + ".self_q_proj", DD); L.sk = upw(b + ".self_k_proj", (int64_t) kv_hidden * hidden); L.sv = upw(b + ".self_v_proj", (int64_t) kv_hidden * hidden); L.so = upw(b + ".self_o_proj", DD);
        L.cq = upw(b + ".cross_q_proj", DD); L.ck = upw(b + ".cross_k_proj", (int64_t) hidden * enc_hidden); L.cv = upw(b + ".cross_v_proj", (int64_t) hidden * enc_hidden); L.co = upw(b + ".cross_o_proj", DD);
        L.gate = upw(b + ".gate", (int64_t) ffn * hidden); L.up = upw(b + ".up", (int64_t) ffn * hidden); L.down = upw(b + ".wo", (int64_t) hidden * ffn);
    }
    if (ok) {   // the n_out codebook tables and output heads, each family in one buffer
        std::vector<float> tab, hw;
        bool heads_f16 = true, heads_any_f16 = false;
        for (int i = 0; i < n_out && ok; i++) {
            const HostTensor * t = find("decoder.embeddings." + std::to_string(i), (int64_t) vocab * hidden);
            const HostTensor * h = find("decoder.heads." + std::to_string(i), (int64_t) vocab * hidden);
            if (!t || !h) break;
            tab.insert(tab.end(), t->v.begin(), t->v.end());
            hw.insert(hw.end(), h->v.begin(), h->v.end());
            heads_f16 = heads_f16 && h->f16; heads_any_f16 = heads_any_f16 || h->f16;
            if (h->qtype) { set_error("dia: block-quantised output heads (quantize --quantize-output-heads) are not supported"); ok = false; break; }
        }
        if (ok && heads_any_f16 != heads_f16) { set_error("dia: the output heads mix F16 and F32 tensors"); ok = false; }
        if (ok) { tables = dev(tab.data(), tab.size()); heads_w = dev_mat(hw.data(), hw.size(), heads_f16); }   // tables: ggml_get_rows widens F16 rows to fp32 exactly
        if (ok && !heads_f16) {   // F32 heads (dia_is_quantizable leaves them F32): fp16 (hi, 2^11-scaled lo) planes for the persistent decode kernel's split tensor-core product
            std::vector<__half> hi(hw.size()), lo(hw.size());
            for (size_t i = 0; i < hw.size(); i++) { hi[i] = __float2half_rn(hw[i]); lo[i] = __float2half_rn((hw[i] - __half2float(hi[i])) * GM_LO_SCALE); }
            for (int pl = 0; pl < 2 && ok; pl++) {
                void * d = nullptr;
                if (cudaMalloc(&d, hw.size() * 2) != cudaSuccess) { cudaGetLastError(); set_error("dia: cudaMalloc of %zu bytes failed", hw.size() * 2); ok = false; break; }
                cudaMemcpy(d, pl ? lo.data() : hi.data(), hw.size() * 2, cudaMemcpyHostToDevice);
                dev_allocs.push_back(d); weight_bytes += hw.size() * 2;
                (pl ? heads_lo : heads_hi) = (__half *) d;
            }
        }
    }
    if (!ok) return 1;
    for (int i = 0; i < 2; i++) B2_CUDA(cudaEventCreate(&ev[i]));
    host.clear();
    B2_CUDA(cudaDeviceSynchronize());      // the uploads above are blocking copies on the legacy stream; kernels run on ctx->stream (non-blocking), which does not wait for it by itself
#ifndef B2EMU
    B2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
#endif
    prepared = true;
    return 0;
}

void Dia::free_all() {
    for (void * p : dev_allocs) cudaFree(p);
    dev_allocs.clear();
    arena.release();
    for (int i = 0; i < 2; i++) if (ev[i]) cudaEventDestroy(ev[i]);
}

namespace {

// cross-attention keys exist only for the prompt's positions: the rest of the enc_ctx-long key store stays zero and is attended to unmasked
// (build_dia_cross_kv_store, model.cpp:392-424 stores sentence_length rows of K but all rows of V)
__global__ void zero_key_tail_kernel(float * K, const int * __restrict__ seq_len, int C, int D) {
    const int r = blockIdx.x, seq = r / C, n = r % C;
    if (n < seq_len[seq]) return;
    for (int c = threadIdx.x; c < D; c += blockDim.x) K[(size_t) r * D + c] = 0.f;
}

// the decoder rows of one step.  Both sequences of an utterance are fed the same audio tokens: head i gets BOS until the decode position exceeds i,
// then the token it produced one step earlier (generate_from_batch, model.cpp:843-858), with check_stopping's end-of-stream injection on top
// (model.cpp:806-823: `delay` is dia_context::delay_steps; `stopped` records the step at which the reference's loop would have ended).
__global__ void dia_step_rows_kernel(const int * __restrict__ d_out, int B, int n_out, const int * __restrict__ d_step, int bos, int eos, int pad, int max_gen, int max_delay, int Tmax,
                                     int * delay, int * stopped, int * ids, int * row_pos, int * row_base, int * row_len, int * row_dst) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= B) return;
    const int step = *d_step;                                      // device-resident so that a captured graph of one step can be replayed
    const int * last = d_out + (size_t) (step > 0 ? step - 1 : 0) * B * n_out;
    const int pattern[9] = {0, 8, 9, 10, 11, 12, 13, 14, 15};     // dia_model::delay_pattern (model.h:85)
    int audio[9];
    for (int i = 0; i < n_out; i++) audio[i] = step > i ? last[u * n_out + i] : bos;
    int d = delay[u];
    if (d == -1 && (audio[0] == eos || step >= max_gen - max_delay)) d = max_delay;
    if (d > 0) {
        const int after = max_delay - d;
        for (int i = 0; i < n_out; i++) {
            if (after == pattern[i]) audio[i] = eos;
            else if (after > pattern[i]) audio[i] = pad;
        }
        d -= 1;
    }
    delay[u] = d;
    if (d == 0 && stopped[u] < 0) stopped[u] = step;
    for (int q = 0; q < 2; q++) {
        const int r = 2 * u + q;
        for (int i = 0; i < n_out; i++) ids[r * n_out + i] = audio[i];
        row_pos[r] = step; row_base[r] = r * Tmax; row_len[r] = step + 1; row_dst[r] = r * Tmax + step;
    }
}

// cfg_scale (src/util.cpp:175-200): out = cond + scale * (cond - uncond); the "-inf above max_output" store is overwritten by it and has no effect
__global__ void cfg_combine_kernel(const float * __restrict__ logits2, int NV, float scale, float * __restrict__ out) {
    const int u = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= NV) return;
    const float cr = logits2[(size_t) (2 * u) * NV + j], ur = logits2[(size_t) (2 * u + 1) * NV + j];
    out[(size_t) u * NV + j] = cr + scale * (cr - ur);
}

struct DFwd : ArLaunch {
    Dia * m; bool fail = false;
    DFwd(Dia * m_, Ctx * c, cudaStream_t s) : m(m_) { ctx = c; st = s; }
    template <class T> T * al(size_t n) { T * p = (T *) m->arena.alloc(n * sizeof(T)); if (!p) fail = true; return p; }
    // 2 or 3 matrices against the same rows (q / k / v, gate / up, cross k / v): one grouped launch by default, separate launches with B2TTS_AR_FUSE=0
    int gemv_n(const float * X, int ldx, int K, int R, int n, const ArW * const * W, const int * N, float * const * Y) {
        if (!ar_fuse_enabled()) { for (int i = 0; i < n; i++) if (gemv(X, ldx, *W[i], K, N[i], R, nullptr, Y[i], N[i])) return 1; return 0; }
        GemvOut o[3];
        for (int i = 0; i < n; i++) o[i] = GemvOut{nullptr, Y[i], nullptr, N[i], 0};
        return gemv_group(X, ldx, K, R, W, N, o, n);
    }
    int rms(const float * x, const float * w, int H, int R, float * y) { rmsnorm_kernel<<<cdiv(R, 8), 256, 0, st>>>(x, w, H, R, y); B2_LAUNCH_CHECK(ctx); return 0; }
    int rope(float * x, const int * pos, int R, int nh, int hd, float theta_scale) { dim3 grid(R, nh); rope_rows_kernel<<<grid, 64, 0, st>>>(x, pos, nh, hd, theta_scale); B2_LAUNCH_CHECK(ctx); return 0; }
    int swiglu(float * g, const float * u, size_t n) { silu_mul_kernel<<<cdiv((int64_t) n, 256), 256, 0, st>>>(g, u, n); B2_LAUNCH_CHECK(ctx); return 0; }
};

}  // namespace

int Dia::generate(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const ArSampling * sampling, int32_t * out_tokens, float * out_logits, int32_t * n_generated,
                  const int32_t * teacher) {
    const ArSampling samp = sampling ? *sampling : ArSampling();
    if (!prepared) { set_error("dia: model not prepared"); return 1; }
    if (B <= 0 || n_steps <= 0) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int C = enc_ctx, S2 = 2 * B, RE = S2 * C, EH = enc_hidden, EI = enc_inner, D = hidden, KVD = kv_hidden, NV = n_out * vocab, Tmax = n_steps;
    for (int b = 0; b < B; b++) {
        if (n_prompt[b] <= 0 || n_prompt[b] > C) { set_error("dia: prompt %d has %d tokens (1 .. %d supported)", b, n_prompt[b], C); return 1; }
        for (int i = 0; i < n_prompt[b]; i++) if (prompts[b][i] >= (uint32_t) enc_vocab) { set_error("dia: prompt %d token %u >= encoder vocabulary %d", b, prompts[b][i], enc_vocab); return 1; }
    }
    // ---- the persistent decode kernel (pdk.cuh) runs the decoder loop when the model and the request fit it: greedy / teacher-forced, <= 8 utterances (16 rows with
    // their unconditional twins), F16 decoder matrices (BASELINE config 4), decoder width <= 2 048 in whole 256-column k-slices.  B2TTS_AR_PDK=0 turns it off.
    static const bool pdk_env = [] { const char * e = getenv("B2TTS_AR_PDK"); return !(e && e[0] == '0'); }();
    // fp32 pages by default for Dia (B2TTS_KV=f16: fp16 pages): its decoder attends with softmax scale 1.0 and amplifies the cache's rounding -- fp16 pages moved the
    // logits of the test model by up to 2.5 at a logit std of 13 (fp32 pages: 0.05, the summation-order floor); the cache is a small share of the step's traffic here
    static const bool kv_f32 = [] { const char * e = getenv("B2TTS_KV"); return !(e && (e[0] == 'f' || e[0] == 'F') && e[1] == '1'); }();
#ifdef B2EMU
    const int pk_grid = [] { const char * e = getenv("B2TTS_PDK_GRID"); const int v = e ? atoi(e) : 3; return v > 0 ? v : 3; }();
#else
    const int pk_grid = [&] { const char * e = getenv("B2TTS_PDK_GRID"); const int v = e ? atoi(e) : 0; return v > 0 && v <= sm_count ? v : sm_count; }();
#endif
    const int pk_ak = [&] { const char * e = getenv("B2TTS_PDK_AK"); const int v = e ? atoi(e) : 0; return v >= 256 && v % 256 == 0 && v <= PK_AK_MAX && v >= D ? v : (D <= 2048 ? 2048 : PK_AK_MAX); }();
    bool use_pdk = pdk_env && !samp.do_sample && S2 <= 16 && D % 256 == 0 && ffn % 256 == 0 && D <= pk_ak && heads * head_dim == D && (head_dim == 64 || head_dim == 128) && (heads_w.f16 || heads_hi) && !heads_w.qtype && n_out <= 9 &&
                   pk_grid > 0 && (ffn <= pk_ak || cdiv(D / 8, pk_grid) <= 3) && (!out_logits || (size_t) n_steps * B * NV * 4 <= ((size_t) 1 << 30));
    for (const DiaDecLayer & L : dec) for (const ArW * w : {&L.sq, &L.sk, &L.sv, &L.so, &L.cq, &L.co, &L.gate, &L.up, &L.down}) use_pdk = use_pdk && w->f16 && !w->qtype;
    const int pk_max_pages = cdiv(Tmax, PK_PAGE);
    const size_t pk_layer_bytes = (size_t) S2 * pk_max_pages * 2 * KVD * PK_PAGE * (kv_f32 ? 4 : 2);
    // few (row, head) items and long contexts (the 1 024-position encodings): every attention item is cut into position chunks spread over the idle CTAs, then combined
    const int pk_tsplit = [&] { const char * e = getenv("B2TTS_PDK_TSPLIT"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 8 ? v : std::max(1, std::min(8, 2 * pk_grid / std::max(1, S2 * heads))); }();
    const size_t pk_need = use_pdk ? (size_t) dec_layers * pk_layer_bytes + (size_t) S2 * pk_max_pages * 4 + (size_t) (10 * dec_layers + 8) * sizeof(PkOp) + 8192 + (size_t) S2 * heads * pk_tsplit * (head_dim + 4) * 4 + (size_t) PK_REP * 16 * ((size_t) 8 * D + 2 * ffn) + 4096 +
                                     (size_t) 16 * head_dim * 4 + (out_logits ? (size_t) n_steps * B * NV * 4 : 0) : 0;
    const size_t enc_ws = (size_t) RE * ((size_t) 3 * EH + 4 * EI + 2 * enc_ffn) * 4;
    const size_t cross = (size_t) 2 * dec_layers * RE * D * 4;
    const size_t self_cache = use_pdk ? 0 : (size_t) 2 * dec_layers * S2 * Tmax * KVD * 4;      // the persistent kernel keeps the self-attention cache in its pages
    const size_t dec_ws = (size_t) S2 * ((size_t) 4 * D + 2 * KVD + 2 * ffn + NV) * 4 + (size_t) B * NV * 4;
    const size_t need = pk_need + enc_ws + cross + self_cache + dec_ws + (size_t) RE * 16 + (size_t) S2 * (32 + 4 * n_out) + (size_t) n_steps * B * n_out * 4 + (size_t) B * 16 + (32 << 20) +
                        (size_t) B * n_out * 8 + (teacher ? (size_t) n_steps * B * n_out * 4 : 0) + (sampling_needs_scratch(samp, vocab) ? (size_t) B * NV * 4 : 0);
    if (arena.reserve(need)) return 1;
    DFwd Fw(this, ctx, st);
    // ---- buffers
    float * ck = Fw.al<float>((size_t) dec_layers * RE * D), * cv = Fw.al<float>((size_t) dec_layers * RE * D);
    float * Kc = use_pdk ? nullptr : Fw.al<float>((size_t) dec_layers * S2 * Tmax * KVD), * Vc = use_pdk ? nullptr : Fw.al<float>((size_t) dec_layers * S2 * Tmax * KVD);
    int * e_tok = Fw.al<int>((size_t) RE), * e_pos = Fw.al<int>((size_t) RE), * e_base = Fw.al<int>((size_t) RE), * e_len = Fw.al<int>((size_t) RE);
    int * seq_len = Fw.al<int>((size_t) S2), * cross_base = Fw.al<int>((size_t) S2), * cross_len = Fw.al<int>((size_t) S2);
    int * ids = Fw.al<int>((size_t) S2 * n_out), * row_pos = Fw.al<int>((size_t) S2), * row_base = Fw.al<int>((size_t) S2), * row_len = Fw.al<int>((size_t) S2), * row_dst = Fw.al<int>((size_t) S2);
    int * delay = Fw.al<int>((size_t) B), * stopped = Fw.al<int>((size_t) B), * d_out = Fw.al<int>((size_t) n_steps * B * n_out), * d_step = Fw.al<int>(1);
    int * s_last = Fw.al<int>((size_t) B * n_out), * s_cnt = Fw.al<int>((size_t) B * n_out);
    int * d_teacher = teacher ? Fw.al<int>((size_t) n_steps * B * n_out) : nullptr;
    float * s_scratch = sampling_needs_scratch(samp, vocab) ? Fw.al<float>((size_t) B * NV) : nullptr;
    if (Fw.fail) return 1;
    B2_CUDA(cudaMemsetAsync(s_last, 0xff, (size_t) B * n_out * 4, st));     // sampler::reset: last_token_ids = -1, repetition_counts = 0
    B2_CUDA(cudaMemsetAsync(s_cnt, 0, (size_t) B * n_out * 4, st));

    std::vector<int> htok((size_t) RE, 0), hpos((size_t) RE), hbase((size_t) RE), hlen((size_t) RE), hseq((size_t) S2), hcb((size_t) S2), hcl((size_t) S2, C), hm1((size_t) B, -1);
    for (int s = 0; s < S2; s++) {
        const int u = s >> 1, S = n_prompt[u];
        hseq[(size_t) s] = S;                   // one mask for both sequences of the pair: the conditional prompt's length (model.cpp:352-366)
        hcb[(size_t) s] = s * C;
        for (int n = 0; n < C; n++) {
            const size_t r = (size_t) s * C + n;
            if (!(s & 1) && n < S) htok[r] = (int) prompts[u][n];      // the unconditional sequence is all zeros
            hpos[r] = n;
            hbase[r] = s * C + (n < S ? 0 : S);                       // prompt positions see each other, padded positions see each other
            hlen[r] = n < S ? S : C - S;
        }
    }
    B2_CUDA(cudaEventRecord(ev[0], st));
    B2_CUDA(cudaMemcpyAsync(e_tok, htok.data(), htok.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(e_pos, hpos.data(), hpos.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(e_base, hbase.data(), hbase.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(e_len, hlen.data(), hlen.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(seq_len, hseq.data(), hseq.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(cross_base, hcb.data(), hcb.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(cross_len, hcl.data(), hcl.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(delay, hm1.data(), hm1.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(stopped, hm1.data(), hm1.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemsetAsync(d_step, 0, 4, st));
    std::vector<int> hteach;
    if (teacher) {                                                  // [B][n_steps][n_out] -> the device's [n_steps][B][n_out]
        hteach.resize((size_t) n_steps * B * n_out);
        for (int b = 0; b < B; b++) for (int s2 = 0; s2 < n_steps; s2++) for (int i = 0; i < n_out; i++) hteach[((size_t) s2 * B + b) * n_out + i] = teacher[((size_t) b * n_steps + s2) * n_out + i];
        B2_CUDA(cudaMemcpyAsync(d_teacher, hteach.data(), hteach.size() * 4, cudaMemcpyHostToDevice, st));
    }
    B2_CUDA(cudaStreamSynchronize(st));   // the host vectors above are stack-owned

    const float theta_scale = powf(10000.0f, -2.0f / (float) head_dim);
    const int Tcap = std::max(Tmax, C);

    // ---- encoder pass over both sequences of every utterance (build_dia_encoder, model.cpp:440-500); softmax scale 1.0
    {
        float * x = Fw.al<float>((size_t) RE * EH), * xn = Fw.al<float>((size_t) RE * EH), * enc_out = Fw.al<float>((size_t) RE * EH);
        float * q = Fw.al<float>((size_t) RE * EI), * k = Fw.al<float>((size_t) RE * EI), * v = Fw.al<float>((size_t) RE * EI), * att = Fw.al<float>((size_t) RE * EI);
        float * g = Fw.al<float>((size_t) RE * enc_ffn), * up = Fw.al<float>((size_t) RE * enc_ffn);
        if (Fw.fail) return 1;
        embed_kernel<<<RE, 256, 0, st>>>(e_tok, enc_embed, EH, x); B2_LAUNCH_CHECK(ctx);
        for (int l = 0; l < enc_layers; l++) {
            const DiaEncLayer & L = enc[(size_t) l];
            if (Fw.rms(x, L.pre_sa, EH, RE, xn)) return 1;
            { const ArW * W3[3] = {&L.wq, &L.wk, &L.wv}; const int N3[3] = {EI, EI, EI}; float * Y3[3] = {q, k, v}; if (Fw.gemv_n(xn, EH, EH, RE, 3, W3, N3, Y3)) return 1; }
            if (Fw.rope(q, e_pos, RE, enc_heads, head_dim, theta_scale) || Fw.rope(k, e_pos, RE, enc_heads, head_dim, theta_scale)) return 1;
            if (Fw.attend(q, k, v, e_base, e_len, RE, enc_heads, enc_heads, head_dim, Tcap, 1.0f, att)) return 1;
            if (Fw.gemv(att, EI, L.wo, EI, EH, RE, x, xn, EH)) return 1;                    // xn = attention + residual(x)
            if (Fw.rms(xn, L.post_sa, EH, RE, x)) return 1;
            { const ArW * W2[2] = {&L.gate, &L.up}; const int N2[2] = {enc_ffn, enc_ffn}; float * Y2[2] = {g, up}; if (Fw.gemv_n(x, EH, EH, RE, 2, W2, N2, Y2)) return 1; }
            if (Fw.swiglu(g, up, (size_t) RE * enc_ffn)) return 1;
            if (Fw.gemv(g, enc_ffn, L.down, enc_ffn, EH, RE, xn, x, EH)) return 1;           // x = mlp + residual(xn)
        }
        if (Fw.rms(x, enc_norm, EH, RE, enc_out)) return 1;
        for (int l = 0; l < dec_layers; l++) {                                               // cross K (RoPE'd, prompt positions only) and V (all positions)
            const DiaDecLayer & L = dec[(size_t) l];
            float * ckl = ck + (size_t) l * RE * D, * cvl = cv + (size_t) l * RE * D;
            { const ArW * W2[2] = {&L.ck, &L.cv}; const int N2[2] = {D, D}; float * Y2[2] = {ckl, cvl}; if (Fw.gemv_n(enc_out, EH, EH, RE, 2, W2, N2, Y2)) return 1; }
            if (Fw.rope(ckl, e_pos, RE, heads, head_dim, theta_scale)) return 1;
            zero_key_tail_kernel<<<RE, 128, 0, st>>>(ckl, seq_len, C, D); B2_LAUNCH_CHECK(ctx);
        }
    }
    // ---- decoder loop
    float * x = Fw.al<float>((size_t) S2 * D), * xn = Fw.al<float>((size_t) S2 * D), * q = Fw.al<float>((size_t) S2 * D), * att = Fw.al<float>((size_t) S2 * D);
    float * kbuf = Fw.al<float>((size_t) S2 * KVD), * vbuf = Fw.al<float>((size_t) S2 * KVD), * g = Fw.al<float>((size_t) S2 * ffn), * up = Fw.al<float>((size_t) S2 * ffn);
    float * logits2 = Fw.al<float>((size_t) S2 * NV), * logits = Fw.al<float>((size_t) B * NV);
    if (Fw.fail) return 1;
    const int R = S2;
    auto run_step = [&]() -> int {
        dia_step_rows_kernel<<<cdiv(B, 128), 128, 0, st>>>(d_teacher ? d_teacher : d_out, B, n_out, d_step, bos, eos, pad, max_gen, max_delay, Tmax, delay, stopped, ids, row_pos, row_base, row_len, row_dst);
        B2_LAUNCH_CHECK(ctx);
        codebook_embed_kernel<<<R, 256, 0, st>>>(ids, n_out, tables, (size_t) vocab * D, nullptr, row_pos, D, x); B2_LAUNCH_CHECK(ctx);
        for (int l = 0; l < dec_layers; l++) {
            const DiaDecLayer & L = dec[(size_t) l];
            float * Kl = Kc + (size_t) l * S2 * Tmax * KVD, * Vl = Vc + (size_t) l * S2 * Tmax * KVD;
            const float * ckl = ck + (size_t) l * RE * D, * cvl = cv + (size_t) l * RE * D;
            if (Fw.rms(x, L.pre_sa, D, R, xn)) return 1;
            { const ArW * W3[3] = {&L.sq, &L.sk, &L.sv}; const int N3[3] = {D, KVD, KVD}; float * Y3[3] = {q, kbuf, vbuf}; if (Fw.gemv_n(xn, D, D, R, 3, W3, N3, Y3)) return 1; }
            if (ar_fuse_enabled()) {                                                         // RoPE of q (in place) and of k on its way into the cache, v copied: one launch instead of three
                dim3 grid(R, heads + heads / rep);
                rope_append_kernel<<<grid, 64, 0, st>>>(q, kbuf, vbuf, nullptr, nullptr, row_pos, heads, heads / rep, head_dim, theta_scale, Kl, Vl, Tmax, row_dst); B2_LAUNCH_CHECK(ctx);
            } else {
                if (Fw.rope(q, row_pos, R, heads, head_dim, theta_scale) || Fw.rope(kbuf, row_pos, R, heads / rep, head_dim, theta_scale)) return 1;
                store_kv_kernel<<<R, 256, 0, st>>>(kbuf, vbuf, row_dst, KVD, Kl, Vl); B2_LAUNCH_CHECK(ctx);
            }
            if (Fw.attend(q, Kl, Vl, row_base, row_len, R, heads, heads / rep, head_dim, Tcap, 1.0f, att)) return 1;
            if (Fw.gemv(att, D, L.so, D, D, R, x, xn, D)) return 1;                          // xn = self-attention + residual(x)
            if (Fw.rms(xn, L.pre_ca, D, R, x)) return 1;
            if (Fw.gemv(x, D, L.cq, D, D, R, nullptr, q, D)) return 1;
            if (Fw.rope(q, row_pos, R, heads, head_dim, theta_scale)) return 1;              // the cross query is RoPE'd with the decode position
            if (Fw.attend(q, ckl, cvl, cross_base, cross_len, R, heads, heads, head_dim, Tcap, 1.0f, att)) return 1;
            if (Fw.gemv(att, D, L.co, D, D, R, xn, x, D)) return 1;                          // x = cross-attention + residual(xn)
            if (Fw.rms(x, L.pre_mlp, D, R, xn)) return 1;
            { const ArW * W2[2] = {&L.gate, &L.up}; const int N2[2] = {ffn, ffn}; float * Y2[2] = {g, up}; if (Fw.gemv_n(xn, D, D, R, 2, W2, N2, Y2)) return 1; }
            if (Fw.swiglu(g, up, (size_t) R * ffn)) return 1;
            if (Fw.gemv(g, ffn, L.down, ffn, D, R, x, x, D)) return 1;                       // x = mlp + residual(x), in place
        }
        if (Fw.rms(x, dec_norm, D, R, xn)) return 1;
        if (Fw.gemv(xn, D, heads_w, D, NV, R, nullptr, logits2, NV)) return 1;
        { dim3 grid(cdiv(NV, 256), B); cfg_combine_kernel<<<grid, 256, 0, st>>>(logits2, NV, cfg, logits); B2_LAUNCH_CHECK(ctx); }
        if (samp.do_sample) { if (sample_rows(ctx, make_sample_params(samp, logits, B * n_out, vocab, s_last, s_cnt, s_scratch, d_step, d_out))) return 1; }
        else { argmax_rows_kernel<<<B * n_out, 256, 0, st>>>(logits, vocab, d_out, d_step); B2_LAUNCH_CHECK(ctx); }
        step_advance_kernel<<<1, 32, 0, st>>>(d_step); B2_LAUNCH_CHECK(ctx);
        return 0;
    };
    // every `exit_every` steps the per-sequence stop flags are read back (one small sync): when the reference's loop would have ended for EVERY sequence of the
    // batch the remaining steps are skipped
    const int exit_every = [] { const char * e = getenv("B2TTS_AR_EXIT_EVERY"); const int v = e ? atoi(e) : 32; return v > 0 ? v : 32; }();
    const bool track_stop = true;
    std::vector<int32_t> hflags((size_t) B);
    auto all_stopped = [&]() -> int {          // 1 all stopped, 0 not yet, -1 error
        if (cudaMemcpyAsync(hflags.data(), stopped, (size_t) B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { set_error("dia: reading the stop flags failed"); return -1; }
        for (int b = 0; b < B; b++) if (hflags[(size_t) b] < 0) return 0;
        return 1;
    };
    if (use_pdk) {
        // ---- the whole decoder loop inside the persistent kernel.  Program of a step: rows (delay pattern + end-of-stream injection, codebook embeddings of the row
        // pair, RoPE table) | per layer { [RMSNorm] q|k|v GEMV with NeoX rotation + cache append in the epilogue -> GQA self-attention over the pages -> o GEMV +
        // residual -> [RMSNorm] cross-q GEMV with rotation -> cross-attention over the row's encoding -> o GEMV + residual -> [RMSNorm] gate|up GEMV with SwiGLU ->
        // down GEMV + residual } | [RMSNorm] heads GEMV | cfg_scale + argmax
        const float scale = 1.0f;
        unsigned char * pool = (unsigned char *) arena.alloc((size_t) dec_layers * pk_layer_bytes);
        int * page_table = Fw.al<int>((size_t) S2 * pk_max_pages), * first_pos = Fw.al<int>(16);
        PkOp * d_ops = (PkOp *) arena.alloc((size_t) (10 * dec_layers + 8) * sizeof(PkOp));
        float * att_part = pk_tsplit > 1 ? Fw.al<float>((size_t) S2 * heads * pk_tsplit * (head_dim + 4)) : nullptr;
        unsigned * d_bar = (unsigned *) arena.alloc(256);
        float * logits_all = out_logits ? Fw.al<float>((size_t) n_steps * B * NV) : nullptr;
        const size_t xrep = (size_t) 16 * D, grep = (size_t) 16 * ffn;
        float * px = Fw.al<float>(PK_REP * xrep), * pxn = Fw.al<float>(PK_REP * xrep);
        __half * att16 = Fw.al<__half>(PK_REP * xrep), * g16 = Fw.al<__half>(PK_REP * grep);
        float2 * rope_cs = (float2 *) arena.alloc((size_t) 16 * (head_dim / 2) * sizeof(float2));
        if (!pool || !d_ops || !d_bar || !rope_cs || Fw.fail) return 1;
        std::vector<int> hpt((size_t) S2 * pk_max_pages);
        for (size_t i = 0; i < hpt.size(); i++) hpt[i] = (int) i;                            // every row takes its pages in order
        std::vector<PkOp> ops;
        auto seg = [&](const ArW & W, int N, int epi, int pair, int n_units) { PkSeg sg; memset(&sg, 0, sizeof sg); sg.W = (const __half *) W.p; sg.Wp = sg.W; sg.N = N; sg.epi = epi; sg.ldy = N; sg.pair = pair; sg.n_units = n_units; return sg; };
        auto gemv_op = [&](int layer, const float * X, const __half * X16, size_t xr, int K, const float * nw, std::initializer_list<PkSeg> segs) {
            PkOp op; memset(&op, 0, sizeof op);
            op.kind = PK_GEMV; op.layer = layer; op.X = X; op.X16 = X16; op.xrep = xr; op.ldx = K; op.K = K; op.norm = nw ? PKN_RMS : PKN_NONE; op.nw = nw; op.eps = 1e-5f;
            int u = 0;
            for (const PkSeg & sg : segs) { op.seg[op.nseg] = sg; op.seg[op.nseg].unit0 = u; u += sg.n_units; op.nseg++; }
            op.n_units = u;
            ops.push_back(op);
        };
        auto attn_op = [&](int layer, const float * ckp, const float * cvp) {
            PkOp op; memset(&op, 0, sizeof op);
            op.kind = PK_ATTN; op.layer = layer; op.q = q; op.out16 = att16; op.orep = xrep; op.scale = scale; op.ck = ckp; op.cv = cvp; op.cross = ckp ? 1 : 0; op.cross_len = C; op.cross_row_stride = ckp ? (size_t) C * D : 0;
            op.tsplit = pk_tsplit;
            ops.push_back(op);
            if (pk_tsplit > 1) { op.kind = PK_ATTNC; ops.push_back(op); }
        };
        { PkOp op; memset(&op, 0, sizeof op); op.kind = PK_ROWS; ops.push_back(op); }
        const int kv_heads = heads / rep, rope_units_q = heads * (head_dim / 16), rope_units_k = kv_heads * (head_dim / 16);
        for (int l = 0; l < dec_layers; l++) {
            const DiaDecLayer & L = dec[(size_t) l];
            PkSeg sq = seg(L.sq, D, PKE_ROPE_Q, PKP_ROPE, rope_units_q); sq.Y = q;
            PkSeg sk = seg(L.sk, KVD, PKE_ROPE_K, PKP_ROPE, rope_units_k);
            PkSeg sv = seg(L.sv, KVD, PKE_KV, PKP_NONE, KVD / 8); sv.kv = 1;
            gemv_op(l, px, nullptr, xrep, D, L.pre_sa, {sq, sk, sv}); ops.back().kv_prefetch = 1;
            attn_op(l, nullptr, nullptr);
            PkSeg so = seg(L.so, D, PKE_RES, PKP_NONE, D / 8); so.Y = pxn; so.res = px; so.yrep = xrep;                    // xn = self-attention + residual(x)
            gemv_op(l, nullptr, att16, xrep, D, nullptr, {so});
            PkSeg scq = seg(L.cq, D, PKE_ROPE_Q, PKP_ROPE, rope_units_q); scq.Y = q;                                      // the cross query is RoPE'd with the decode position
            gemv_op(l, pxn, nullptr, xrep, D, L.pre_ca, {scq});
            attn_op(l, ck + (size_t) l * RE * D, cv + (size_t) l * RE * D);
            PkSeg sco = seg(L.co, D, PKE_RES, PKP_NONE, D / 8); sco.Y = px; sco.res = pxn; sco.yrep = xrep;                // x = cross-attention + residual(xn)
            gemv_op(l, nullptr, att16, xrep, D, nullptr, {sco});
            PkSeg sg2 = seg(L.gate, ffn, PKE_SWIGLU, PKP_SWIGLU, ffn / 8); sg2.Wp = (const __half *) L.up.p; sg2.Y16 = g16; sg2.yrep = grep;
            gemv_op(l, px, nullptr, xrep, D, L.pre_mlp, {sg2});
            PkSeg sd = seg(L.down, D, PKE_RES, PKP_NONE, D / 8); sd.Y = px; sd.res = px; sd.yrep = xrep;                   // x = mlp + residual(x), element-wise in place (every copy)
            gemv_op(l, nullptr, g16, grep, ffn, nullptr, {sd});
        }
        {
            PkSeg sh = seg(heads_w, NV, PKE_LOGITS, PKP_NONE, cdiv(NV, 8)); sh.Y = logits2;
            if (!heads_w.f16) { sh.W = heads_hi; sh.Wp = heads_hi; sh.Wl = heads_lo; }
            gemv_op(0, px, nullptr, xrep, D, dec_norm, {sh});
        }
        { PkOp op; memset(&op, 0, sizeof op); op.kind = PK_ARGMAX; ops.push_back(op); }
        B2_CUDA(cudaMemcpyAsync(page_table, hpt.data(), hpt.size() * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_ops, ops.data(), ops.size() * sizeof(PkOp), cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemsetAsync(first_pos, 0, 64, st));
        PkParams Pk; memset(&Pk, 0, sizeof Pk);
        Pk.ops = d_ops; Pk.n_ops = (int) ops.size(); Pk.R = S2; Pk.H = D; Pk.heads = heads; Pk.kv_heads = kv_heads; Pk.hd = head_dim; Pk.n_out = n_out; Pk.vocab = vocab;
        Pk.model = PKM_DIA; Pk.ak = pk_ak; Pk.pos_off = 0; Pk.n_steps_total = n_steps;
        Pk.rope_cs = rope_cs; Pk.theta_scale = theta_scale; Pk.pad = pad; Pk.max_delay = max_delay; Pk.cfg = cfg; Pk.delay = delay; Pk.logits_cfg = logits; Pk.att_part = att_part;
        Pk.bar = d_bar; Pk.d_step = d_step; Pk.first_pos = first_pos; Pk.d_out = d_out; Pk.d_teacher = d_teacher; Pk.bos = bos; Pk.eos = eos; Pk.max_gen = max_gen; Pk.stopped = stopped; Pk.ids = ids; Pk.row_pos = row_pos;
        Pk.tables = tables; Pk.tab_stride = (size_t) vocab * D; Pk.x0 = px; Pk.x0rep = xrep;
        Pk.kv_pool = pool; Pk.kv_layer_bytes = pk_layer_bytes; Pk.page_table = page_table; Pk.max_pages = pk_max_pages;
        Pk.logits = logits2; Pk.logits_all = logits_all;
        PkLaunch pkl;
        if (pk_configure(Pk, kv_f32, std::min(std::max(D, ffn), pk_ak), heads_w.f16 ? 0 : std::min(D, pk_ak), std::max(Tmax, C), pkl)) { set_error("dia: the persistent decode kernel does not fit this shape (%d positions) in shared memory", Tmax); return 1; }
        pk_prof_begin(Pk, ops.size(), pk_grid, st);
        for (int s0 = 0; s0 < n_steps; s0 += exit_every) {
            if (s0 > 0) { const int a = all_stopped(); if (a < 0) return 1; if (a) break; }
            Pk.step_begin = s0; Pk.n_steps = std::min(exit_every, n_steps - s0);
            B2_CUDA(pk_launch(pkl, Pk, pk_grid, st));
            ctx->launches++; pdk_launches++; pdk_steps += (uint64_t) Pk.n_steps;
        }
        pk_prof_end(Pk, ops, pk_grid, st);
        if (out_logits)
            for (int s = 0; s < n_steps; s++)
                for (int b = 0; b < B; b++)
                    B2_CUDA(cudaMemcpyAsync(out_logits + ((size_t) b * n_steps + s) * NV, logits_all + ((size_t) s * B + b) * NV, (size_t) NV * 4, cudaMemcpyDeviceToHost, st));
    }
    // B2TTS_AR_GRAPH=1: capture one decoder step into a CUDA graph and replay it (see parler.cu); not used when every step's logits go to the host
    const char * ge = getenv("B2TTS_AR_GRAPH");
    if (use_pdk) {
    } else if (!(ge && ge[0] == '0') && !out_logits && n_steps > 1) {      // on by default (reproduced the reference's tokens on a B200, round 2); B2TTS_AR_GRAPH=0 for A/B runs
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        if (run_step()) return 1;                               // step 0 runs directly: every kernel instantiation has its attributes set before the capture
        const uint64_t l0 = ctx->launches;
        B2_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = run_step();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (!rc) set_error("dia: stream capture failed: %s", cudaGetErrorString(ce)); return 1; }
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); set_error("dia: cudaGraphInstantiate failed"); return 1; }
        cudaError_t le = cudaSuccess;
        for (int s = 1; s < n_steps && le == cudaSuccess; s++) {
            le = cudaGraphLaunch(exec, st);
            if (track_stop && le == cudaSuccess && (s + 1) % exit_every == 0 && s + 1 < n_steps) { const int a = all_stopped(); if (a < 0) { le = cudaErrorUnknown; } else if (a) break; }
        }
        ctx->launches += (uint64_t) (n_steps - 2) * (ctx->launches - l0);
        cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
        if (le != cudaSuccess) { set_error("dia: cudaGraphLaunch failed: %s", cudaGetErrorString(le)); return 1; }
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
    std::vector<int32_t> tmp((size_t) n_steps * B * n_out), hstop((size_t) B);
    B2_CUDA(cudaMemcpyAsync(tmp.data(), d_out, tmp.size() * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(hstop.data(), stopped, hstop.size() * 4, cudaMemcpyDeviceToHost, st));
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
