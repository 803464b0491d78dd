// orpheus.cu -- Orpheus autoregressive decode (llama-3 style), first correct CUDA path.  See orpheus.h for what it replaces and why it is plain.
#include "orpheus.h"
#include "ar_kernels.cuh"
#include "pdk.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace b2 {

int Orpheus::assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {
    if (prepared) { set_error("orpheus: assign_weight after prepare"); return 1; }
    std::string nm(name);
    if (nm.rfind("orpheus.", 0) == 0) nm = nm.substr(8);
    HostTensor t;
    if (host_tensor_from_gguf(t, name, type, n_dims, ne, data, nbytes, true)) return 1;
    host[nm] = std::move(t);
    return 0;
}

int Orpheus::prepare() {
    if (prepared) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    auto kvreq = [&](const char * k, int & out) { auto it = kv.find(k); if (it == kv.end()) { set_error("the '%s' key must be specified in the GGUF file.", k); return 1; } out = (int) it->second; return 0; };
    if (kvreq("orpheus.layers", n_layers) || kvreq("orpheus.vocab_size", vocab) || kvreq("orpheus.attn_heads", heads) || kvreq("orpheus.kv_attn_heads", kv_heads) ||
        kvreq("orpheus.head_dim", head_dim) || kvreq("orpheus.hidden_size", hidden) || kvreq("orpheus.kv_hidden_size", kv_hidden)) return 1;
    { auto it = kv.find("orpheus.stopping_token_id"); stopping_token = it != kv.end() ? (int) it->second : 128258; }
    if (hidden != heads * head_dim || kv_hidden != kv_heads * head_dim || heads % kv_heads || head_dim % 4 || hidden % 4) { set_error("orpheus: inconsistent head configuration"); return 1; }
    { auto it = kv.find("orpheus.context_length"); max_context = it != kv.end() ? (int) it->second : 0; }
    bool ok = true;
    auto find = [&](const std::string & n, int64_t expect) -> const HostTensor * {
        auto it = host.find(n);
        if (it == host.end()) { set_error("missing tensor orpheus.%s", n.c_str()); ok = false; return nullptr; }
        if (expect && (int64_t) it->second.v.size() != expect) { set_error("tensor orpheus.%s has %zu elements, expected %lld", n.c_str(), it->second.v.size(), (long long) expect); ok = false; return nullptr; }
        return &it->second;
    };
    auto dev = [&](const void * src, size_t bytes) -> void * {
        void * d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaGetLastError(); set_error("orpheus: cudaMalloc of %zu bytes failed", bytes); ok = false; return nullptr; }
        cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
        dev_allocs.push_back(d); weight_bytes += bytes;
        return d;
    };
    auto up = [&](const std::string & n, int64_t expect) -> float * { const HostTensor * t = find(n, expect); return t ? (float *) dev(t->v.data(), t->v.size() * 4) : nullptr; };   // fp32 values (F16 / blocks widened exactly like ggml_get_rows)
    auto upw = [&](const std::string & n, int64_t expect) -> ArW {      // a matrix in its stored form
        ArW w;
        const HostTensor * t = find(n, expect);
        if (!t) return w;
        if (t->qtype) { if (!upload_quant_planes(*t, w, dev_allocs, weight_bytes)) ok = false; return w; }
        if (t->f16) {
            std::vector<__half> h(t->v.size());
            for (size_t i = 0; i < h.size(); i++) h[i] = __float2half_rn(t->v[i]);
            w.f16 = true; w.p = dev(h.data(), h.size() * 2);
            return w;
        }
        w.p = dev(t->v.data(), t->v.size() * 4);
        if (gemv_split_mma_enabled() && t->shape.size() == 2 && w.p) {                   // the split copy of an F32 matrix: W = hi + lo, lo carried scaled by 2^11
            const size_t cnt = t->v.size();
            std::vector<__half> hi(cnt), lo(cnt);
            for (size_t i = 0; i < cnt; i++) { const float x = t->v[i]; hi[i] = __float2half_rn(x); lo[i] = __float2half_rn((x - __half2float(hi[i])) * GM_LO_SCALE); }
            const void * dh = dev(hi.data(), cnt * 2), * dl = dev(lo.data(), cnt * 2);
            if (dh && dl) split[(const float *) w.p] = {dh, dl};
        }
        return w;
    };
    embed = up("embed_tokens", (int64_t) vocab * hidden);
    out_norm = up("norm", hidden);
    head = upw("lm_head", (int64_t) vocab * hidden);
    rope_ff = up("rope_frequencies", head_dim / 2);
    {
        auto it = host.find("layers.0.mlp.gate_proj");
        if (it == host.end()) { set_error("missing tensor orpheus.layers.0.mlp.gate_proj"); return 1; }
        ffn = (int) it->second.shape[0];
        if (ffn % 4) { set_error("orpheus: ffn size %d must be a multiple of 4", ffn); return 1; }
    }
    layers.resize((size_t) n_layers);
    for (int l = 0; l < n_layers && ok; l++) {
        const std::string b = "layers." + std::to_string(l);
        OrpheusLayer & L = layers[(size_t) l];
        L.in_norm = up(b + ".input_layernorm", hidden);  L.post_norm = up(b + ".post_attention_layernorm", hidden);
        L.wq = upw(b + ".self_attn.q_proj", (int64_t) hidden * hidden);    L.wk = upw(b + ".self_attn.k_proj", (int64_t) kv_hidden * hidden);
        L.wv = upw(b + ".self_attn.v_proj", (int64_t) kv_hidden * hidden); L.wo = upw(b + ".self_attn.o_proj", (int64_t) hidden * hidden);
        L.wgate = upw(b + ".mlp.gate_proj", (int64_t) ffn * hidden);       L.wup = upw(b + ".mlp.up_proj", (int64_t) ffn * hidden);
        L.wdown = upw(b + ".mlp.down_proj", (int64_t) hidden * ffn);
        for (auto * w : {&L.wq, &L.wk, &L.wv, &L.wo, &L.wgate, &L.wup, &L.wdown}) { (void) w; }
        host.erase(b + ".self_attn.q_proj"); host.erase(b + ".self_attn.k_proj"); host.erase(b + ".self_attn.v_proj"); host.erase(b + ".self_attn.o_proj");      // free the host copies layer by layer
        host.erase(b + ".mlp.gate_proj"); host.erase(b + ".mlp.up_proj"); host.erase(b + ".mlp.down_proj");
    }
    if (!ok) return 1;
    for (int i = 0; i < 2; i++) B2_CUDA(cudaEventCreate(&ev[i]));
#ifndef B2EMU
    B2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
#endif
    host.clear();
    B2_CUDA(cudaDeviceSynchronize());      // the uploads above are blocking copies on the legacy stream; kernels run on ctx->stream (non-blocking), which does not wait for it by itself
    prepared = true;
    return 0;
}

void Orpheus::free_all() {
    for (void * p : dev_allocs) cudaFree(p);
    dev_allocs.clear();
    arena.release();
    for (int i = 0; i < 2; i++) if (ev[i]) cudaEventDestroy(ev[i]);
}

namespace {

// generate_from_batch's stop rule (model.cpp:389-398): a sequence has ended once it produced the stopping token; `stopped[b]` keeps the number of tokens up to and
// including it.  Also advances the device-resident step counter (one launch instead of two).
__global__ void orpheus_stop_advance_kernel(int * d_step, const int * __restrict__ cur_tok, int * stopped, int B, int stop_token) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int step = *d_step;
    if (stopped && b < B && stopped[b] < 0 && cur_tok[b] == stop_token) stopped[b] = step + 1;
    __syncthreads();
    if (b == 0) *d_step = step + 1;
}

struct OFwd : ArLaunch {
    Orpheus * m; bool fail = false;
    OFwd(Orpheus * m_, Ctx * c, cudaStream_t s) : m(m_) { ctx = c; st = s; }
    template <class T> T * al(size_t n) { T * p = (T *) m->arena.alloc(n * sizeof(T)); if (!p) fail = true; return p; }
    // up to 3 matrices against the same rows in one launch when they share a storage kind (F32: fp32-faithful tensor-core path when every one has its fp16 split)
    int gemv_n(const float * X, int ldx, int K, int R, const ArW * const * W, const int * N, float * const * Y, int n) {
        bool f32 = true;
        for (int i = 0; i < n; i++) f32 = f32 && !W[i]->f16 && !W[i]->qtype;
        if (!f32) {
            GemvOut o[3];
            for (int i = 0; i < n; i++) o[i] = GemvOut{nullptr, Y[i], nullptr, N[i], 0};
            return gemv_group(X, ldx, K, R, W, N, o, n);
        }
        GemvItem it[3];
        bool mma = gemv_split_mma_enabled();
        for (int i = 0; i < n && mma; i++) mma = gemv_mma_ok(K, N[i], 16) && m->split.find((const float *) W[i]->p) != m->split.end();
        for (int i = 0; i < n; i++) {
            it[i] = GemvItem{W[i]->p, nullptr, nullptr, N[i], GemvOut{nullptr, Y[i], nullptr, N[i], 0}};
            if (mma) { const auto & sp = m->split.find((const float *) W[i]->p)->second; it[i].W = sp.first; it[i].W2 = sp.second; }
        }
        return gemv_group_launch(ctx, st, group_smem, mma ? GEMV_SPLIT_MMA : GEMV_F32, 0, X, ldx, K, R, it, n);
    }
    // one matrix in any storage kind (F32: fp32-faithful tensor-core path when its fp16 split exists)
    int gemv_w(const float * X, int ldx, const ArW & W, int K, int N, int R, const float * res, float * Y, int ldy) {
        if (W.f16 || W.qtype) return gemv(X, ldx, W, K, N, R, res, Y, ldy);
        if (gemv_split_mma_enabled() && gemv_mma_ok(K, N, 16)) {
            auto it = m->split.find((const float *) W.p);
            if (it != m->split.end()) return gemv_mma_launch(ctx, st, mma_smem_set, X, ldx, (const __half *) it->second.first, (const __half *) it->second.second, K, N, R, res, Y, ldy);
        }
        gemv_rows_launch(st, X, ldx, W.p, false, K, N, R, res, Y, ldy);
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
};

}  // namespace

int Orpheus::generate(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const ArSampling * sampling, int32_t * out_tokens, float * out_logits,
                      int32_t * n_generated) {
    const ArSampling samp = sampling ? *sampling : ArSampling();
    if (!prepared) { set_error("orpheus: model not prepared"); return 1; }
    if (B <= 0 || n_steps <= 0) return 0;
    // more than 16 sequences, greedy, F16 or Q8_0 matrices: groups of 16, each inside the persistent decode kernel (sequences are independent)
    if (B > 16 && !samp.do_sample && (head.f16 || head.qtype == 8) && !(getenv("B2TTS_AR_PDK") && getenv("B2TTS_AR_PDK")[0] == '0')) {
        float total_ms = 0.f;
        for (int b0 = 0; b0 < B; b0 += 16) {
            const int nb = std::min(16, B - b0);
            if (generate(nb, prompts + b0, n_prompt + b0, n_steps, sampling, out_tokens + (size_t) b0 * n_steps, out_logits ? out_logits + (size_t) b0 * n_steps * vocab : nullptr,
                         n_generated ? n_generated + b0 : nullptr)) return 1;
            total_ms += timing_ms;
        }
        timing_ms = total_ms;
        return 0;
    }
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int R0 = 0, Pmax = 0;
    for (int b = 0; b < B; b++) {
        if (n_prompt[b] <= 0) { set_error("orpheus: prompt %d is empty", b); return 1; }
        for (int i = 0; i < n_prompt[b]; i++) if (prompts[b][i] >= (uint32_t) vocab) { set_error("orpheus: prompt %d token %u >= vocab %d", b, prompts[b][i], vocab); return 1; }
        R0 += n_prompt[b]; Pmax = std::max(Pmax, (int) n_prompt[b]);
    }
    const int Tmax = Pmax + n_steps, Rmax = std::max(R0, B), H = hidden, KV = kv_hidden, F = ffn;
    if (max_context > 0 && Tmax > max_context) { set_error("orpheus: %d positions (longest prompt + n_steps) exceed the model's context of %d", Tmax, max_context); return 1; }
    if (B > 128) { set_error("orpheus: at most 128 sequences per call (%d given)", B); return 1; }
    // ---- the persistent decode kernel (pdk.cuh) takes decode steps 1 .. n_steps - 1 when the model and the request fit it: greedy, <= 16 sequences, every matrix F16
    // (block-quantised and F32 GGUFs stay on the launch-per-op path below), hidden size <= 3 072 in whole 256-column k-slices.  B2TTS_AR_PDK=0 turns it off.
    static const bool pdk_env = [] { const char * e = getenv("B2TTS_AR_PDK"); return !(e && e[0] == '0'); }();
    static const bool kv_f32 = [] { const char * e = getenv("B2TTS_KV"); return e && (e[0] == 'f' || e[0] == 'F') && e[1] == '3'; }();     // B2TTS_KV=f32: fp32 pages (default fp16)
#ifdef B2EMU
    const int pk_grid = [] { const char * e = getenv("B2TTS_PDK_GRID"); const int v = e ? atoi(e) : 3; return v > 0 ? v : 3; }();
#else
    const int pk_grid = [&] { const char * e = getenv("B2TTS_PDK_GRID"); const int v = e ? atoi(e) : 0; return v > 0 && v <= sm_count ? v : sm_count; }();
#endif
    const int pk_ak = [&] { const char * e = getenv("B2TTS_PDK_AK"); const int v = e ? atoi(e) : 0; return v >= 256 && v % 256 == 0 && v <= PK_AK_MAX && v >= H ? v : (H <= 2048 ? 2048 : PK_AK_MAX); }();
    bool use_pdk = pdk_env && !samp.do_sample && B <= 16 && n_steps >= 2 && H % 256 == 0 && F % 256 == 0 && H <= pk_ak && (head_dim == 64 || head_dim == 128) && (head.f16 || head.qtype == 8) && pk_grid > 0 &&
                   (F <= pk_ak || cdiv(H / 8, pk_grid) <= 3) && (!out_logits || (size_t) n_steps * B * vocab * 4 <= ((size_t) 1 << 30));
    // every matrix F16, or every matrix Q8_0 (BASELINE config 5's dtype: int8 MMA over Q8_0-quantised activations, ggml_vec_dot_q8_0_q8_0's arithmetic)
    const bool pk_q8 = head.qtype == 8;
    for (const OrpheusLayer & L : layers) for (const ArW * w : {&L.wq, &L.wk, &L.wv, &L.wo, &L.wgate, &L.wup, &L.wdown}) use_pdk = use_pdk && (pk_q8 ? w->qtype == 8 : (w->f16 && !w->qtype));
    const int Tst = use_pdk ? Pmax : Tmax;                          // positions per sequence in the contiguous fp32 cache: the persistent path keeps only the prompt pass there
    const int pk_max_pages = cdiv(Tmax, PK_PAGE);
    int pk_pages = 0;
    for (int b = 0; b < B; b++) pk_pages += cdiv(n_prompt[b] + n_steps, PK_PAGE);
    const size_t pk_layer_bytes = (size_t) pk_pages * 2 * KV * PK_PAGE * (kv_f32 ? 4 : 2);
    const int pk_amax = std::max(1, std::min(256, pk_grid / B));     // chunks a row's argmax is cut into: about one (row, chunk) item per CTA
    const size_t pk_need = use_pdk ? (size_t) n_layers * pk_layer_bytes + (size_t) B * pk_max_pages * 4 + (size_t) (10 * n_layers + 8) * sizeof(PkOp) + (size_t) PK_REP * 16 * std::max(H, F) * 2 + (size_t) R0 * 8 + 8192 + (size_t) PK_REP * 16 * ((size_t) 8 * H + 2 * F) + 4096 +
                                     (size_t) 16 * head_dim * 4 + (size_t) B * pk_amax * 8 + (out_logits ? (size_t) n_steps * B * vocab * 4 : 0) : 0;
    const size_t cache = (size_t) n_layers * B * Tst * KV * 4;
    const size_t need = pk_need + 2 * cache + (size_t) Rmax * ((size_t) 4 * H + 2 * KV + 2 * F) * 4 + (size_t) B * ((size_t) vocab + H) * 4 + (size_t) B * n_steps * 4 +
                        (size_t) Rmax * 32 + (size_t) B * 16 + (32 << 20) + (size_t) B * 8 + (sampling_needs_scratch(samp, vocab) ? (size_t) B * vocab * 4 : 0);
    if (arena.reserve(need)) return 1;
    OFwd Fw(this, ctx, st);
    float * Kc = Fw.al<float>((size_t) n_layers * B * Tst * KV), * Vc = Fw.al<float>((size_t) n_layers * B * Tst * KV);
    float * x = Fw.al<float>((size_t) Rmax * H), * xn = Fw.al<float>((size_t) Rmax * H), * q = Fw.al<float>((size_t) Rmax * H), * att = Fw.al<float>((size_t) Rmax * H);
    float * kbuf = Fw.al<float>((size_t) Rmax * KV), * vbuf = Fw.al<float>((size_t) Rmax * KV);
    float * g = Fw.al<float>((size_t) Rmax * F), * u = Fw.al<float>((size_t) Rmax * F);
    float * last = Fw.al<float>((size_t) B * H), * logits = Fw.al<float>((size_t) B * vocab);
    int * row_seq = Fw.al<int>((size_t) Rmax), * row_pos = Fw.al<int>((size_t) Rmax), * row_tok = Fw.al<int>((size_t) Rmax);
    int * row_base = Fw.al<int>((size_t) Rmax), * row_len = Fw.al<int>((size_t) Rmax);
    int * d_np = Fw.al<int>((size_t) B), * d_last = Fw.al<int>((size_t) B), * cur_tok = Fw.al<int>((size_t) B), * d_out = Fw.al<int>((size_t) B * n_steps), * d_step = Fw.al<int>(1);
    int * s_last = Fw.al<int>((size_t) B), * s_cnt = Fw.al<int>((size_t) B);
    int * stopped = n_generated ? Fw.al<int>((size_t) B) : nullptr;
    float * s_scratch = sampling_needs_scratch(samp, vocab) ? Fw.al<float>((size_t) B * vocab) : nullptr;
    if (Fw.fail) return 1;
    B2_CUDA(cudaMemsetAsync(s_last, 0xff, (size_t) B * 4, st));     // sampler::reset: last_token_ids = -1, repetition_counts = 0
    B2_CUDA(cudaMemsetAsync(s_cnt, 0, (size_t) B * 4, st));
    if (stopped) B2_CUDA(cudaMemsetAsync(stopped, 0xff, (size_t) B * 4, st));

    std::vector<int> hs((size_t) R0), hp((size_t) R0), ht((size_t) R0), hb((size_t) R0), hl((size_t) R0), hnp((size_t) B), hlast((size_t) B);
    {
        int r = 0;
        for (int b = 0; b < B; b++) {
            hnp[(size_t) b] = n_prompt[b];
            for (int i = 0; i < n_prompt[b]; i++, r++) { hs[(size_t) r] = b; hp[(size_t) r] = i; ht[(size_t) r] = (int) prompts[b][i]; hb[(size_t) r] = b * Tst; hl[(size_t) r] = i + 1; }
            hlast[(size_t) b] = r - 1;
        }
    }
    B2_CUDA(cudaEventRecord(ev[0], st));
    B2_CUDA(cudaMemcpyAsync(row_seq, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(row_pos, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(row_tok, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(row_base, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(row_len, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(d_np, hnp.data(), hnp.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(d_last, hlast.data(), hlast.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemsetAsync(d_step, 0, 4, st));
    B2_CUDA(cudaStreamSynchronize(st));   // the host vectors above are stack-owned

    const float theta_scale = powf(500000.0f, -2.0f / (float) head_dim);
    const float scale = 1.0f / sqrtf((float) head_dim);

    // one pass: R rows already described by row_* -> logits of B rows -> argmax into d_out[.][*d_step] and cur_tok; advances d_step
    const bool fuse = ar_fuse_enabled();
    auto run_pass = [&](int R, bool prefill) -> int {
        embed_kernel<<<R, 256, 0, st>>>(row_tok, embed, H, x);
        B2_LAUNCH_CHECK(ctx);
        for (int l = 0; l < n_layers; l++) {
            const OrpheusLayer & L = layers[(size_t) l];
            float * Kl = Kc + (size_t) l * B * Tst * KV, * Vl = Vc + (size_t) l * B * Tst * KV;
            rmsnorm_kernel<<<cdiv(R, 8), 256, 0, st>>>(x, L.in_norm, H, R, xn); B2_LAUNCH_CHECK(ctx);
            if (fuse) {                                                                        // q, k, v in one launch
                const ArW * W3[3] = {&L.wq, &L.wk, &L.wv}; const int N3[3] = {H, KV, KV}; float * Y3[3] = {q, kbuf, vbuf};
                if (Fw.gemv_n(xn, H, H, R, W3, N3, Y3, 3)) return 1;
            } else {
                if (Fw.gemv_w(xn, H, L.wq, H, H, R, nullptr, q, H)) return 1;
                if (Fw.gemv_w(xn, H, L.wk, H, KV, R, nullptr, kbuf, KV)) return 1;
                if (Fw.gemv_w(xn, H, L.wv, H, KV, R, nullptr, vbuf, KV)) return 1;
            }
            { dim3 grid(R, heads + kv_heads); rope_append_kernel<<<grid, 64, 0, st>>>(q, kbuf, vbuf, rope_ff, row_seq, row_pos, heads, kv_heads, head_dim, theta_scale, Kl, Vl, Tst, nullptr); B2_LAUNCH_CHECK(ctx); }
            if (Fw.attend(q, Kl, Vl, row_base, row_len, R, heads, kv_heads, head_dim, Tst, scale, att)) return 1;
            if (Fw.gemv_w(att, H, L.wo, H, H, R, x, xn, H)) return 1;                      // xn = attn_out + residual(x)
            rmsnorm_kernel<<<cdiv(R, 8), 256, 0, st>>>(xn, L.post_norm, H, R, q); B2_LAUNCH_CHECK(ctx);   // q reused as the normalised MLP input
            if (fuse) {                                                                        // gate and up in one launch
                const ArW * W2[2] = {&L.wgate, &L.wup}; const int N2[2] = {F, F}; float * Y2[2] = {g, u};
                if (Fw.gemv_n(q, H, H, R, W2, N2, Y2, 2)) return 1;
            } else {
                if (Fw.gemv_w(q, H, L.wgate, H, F, R, nullptr, g, F)) return 1;
                if (Fw.gemv_w(q, H, L.wup, H, F, R, nullptr, u, F)) return 1;
            }
            { const size_t n = (size_t) R * F; silu_mul_kernel<<<cdiv((int64_t) n, 256), 256, 0, st>>>(g, u, n); B2_LAUNCH_CHECK(ctx); }
            if (Fw.gemv_w(g, F, L.wdown, F, H, R, xn, x, H)) return 1;                       // x = mlp + residual(xn)
        }
        rmsnorm_kernel<<<cdiv(R, 8), 256, 0, st>>>(x, out_norm, H, R, xn); B2_LAUNCH_CHECK(ctx);
        const float * lastp = xn;
        if (prefill) { gather_rows_f32_kernel<<<B, 256, 0, st>>>(xn, d_last, H, last); B2_LAUNCH_CHECK(ctx); lastp = last; }   // logits of the last position only
        if (Fw.gemv_w(lastp, H, head, H, vocab, B, nullptr, logits, vocab)) return 1;
        if (samp.do_sample) {
            SampleParams sp = make_sample_params(samp, logits, B, vocab, s_last, s_cnt, s_scratch, d_step, d_out);
            sp.cur_tok = cur_tok; sp.out_stride_steps = n_steps;
            if (sample_rows(ctx, sp)) return 1;
        } else { argmax_kernel<<<B, 256, 0, st>>>(logits, vocab, cur_tok, d_out, n_steps, d_step); B2_LAUNCH_CHECK(ctx); }
        orpheus_stop_advance_kernel<<<1, 128, 0, st>>>(d_step, cur_tok, stopped, B, stopping_token); B2_LAUNCH_CHECK(ctx);
        return 0;
    };
    // every `exit_every` steps the stop flags are read back (one small sync): when EVERY sequence has produced its stopping token the remaining steps are skipped
    const int exit_every = [] { const char * e = getenv("B2TTS_AR_EXIT_EVERY"); const int v = e ? atoi(e) : 32; return v > 0 ? v : 32; }();
    std::vector<int32_t> hflags((size_t) B, -1);
    auto all_stopped = [&]() -> int {          // 1 all stopped, 0 not yet, -1 error
        if (!stopped) return 0;
        if (cudaMemcpyAsync(hflags.data(), stopped, (size_t) B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { set_error("orpheus: reading the stop flags failed"); return -1; }
        for (int b = 0; b < B; b++) if (hflags[(size_t) b] < 0) return 0;
        return 1;
    };
    auto copy_logits = [&](int s) -> int {
        if (out_logits)
            for (int b = 0; b < B; b++)
                B2_CUDA(cudaMemcpyAsync(out_logits + ((size_t) b * n_steps + s) * vocab, logits + (size_t) b * vocab, (size_t) vocab * 4, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    auto run_decode = [&]() -> int {
        decode_rows_kernel<<<cdiv(B, 128), 128, 0, st>>>(d_np, cur_tok, B, d_step, Tst, row_seq, row_pos, row_tok, row_base, row_len); B2_LAUNCH_CHECK(ctx);
        return run_pass(B, false);
    };
    if (run_pass(R0, true) || copy_logits(0)) return 1;                                   // step 0: the whole ragged batch of prompts
    if (use_pdk) {
        // ---- steps 1 .. n_steps - 1 inside the persistent kernel.  Program of a step: rows (token of the previous step -> embedding row, RoPE table) | per layer
        // { [RMSNorm] q|k|v GEMV with the NeoX rotation and the cache append in its epilogue -> GQA attention over the pages -> o GEMV + residual -> [RMSNorm] gate|up
        // GEMV with SwiGLU in its epilogue -> down GEMV + residual } | [RMSNorm] lm_head GEMV | argmax partials (combined by the next step's rows phase)
        unsigned char * pool = (unsigned char *) arena.alloc((size_t) n_layers * pk_layer_bytes);
        int * page_table = Fw.al<int>((size_t) B * pk_max_pages), * row_src = Fw.al<int>((size_t) R0), * pk_pos = Fw.al<int>(16);
        PkOp * d_ops = (PkOp *) arena.alloc((size_t) (10 * n_layers + 8) * sizeof(PkOp));
        // Q8_0: the phase inputs are quantised ONCE per phase by a PK_QUANT op into these (replicated) buffers; the GEMV phases copy them into shared memory
        const int Kq = std::max(H, F);
        const size_t qrep = (size_t) 16 * Kq, qdrep = (size_t) 16 * (Kq / 32);
        unsigned char * xq = pk_q8 ? (unsigned char *) arena.alloc(PK_REP * qrep) : nullptr;
        float * xqd = pk_q8 ? Fw.al<float>(PK_REP * qdrep) : nullptr;
        if (pk_q8 && (!xq || !xqd)) return 1;
        unsigned * d_bar = (unsigned *) arena.alloc(256);
        float * logits_all = out_logits ? Fw.al<float>((size_t) n_steps * B * vocab) : nullptr;
        const size_t xrep = (size_t) 16 * H, grep = (size_t) 16 * F;
        float * px = Fw.al<float>(PK_REP * xrep), * pxn = Fw.al<float>(PK_REP * xrep);
        __half * att16 = Fw.al<__half>(PK_REP * xrep), * g16 = Fw.al<__half>(PK_REP * grep);
        float2 * rope_cs = (float2 *) arena.alloc((size_t) 16 * (head_dim / 2) * sizeof(float2));
        float * amax_v = Fw.al<float>((size_t) B * pk_amax); int * amax_i = Fw.al<int>((size_t) B * pk_amax);
        if (!pool || !d_ops || !d_bar || !rope_cs || Fw.fail) return 1;
        std::vector<int> hpt((size_t) B * pk_max_pages, 0), hsrc((size_t) R0);
        { int next = 0, r = 0; for (int b = 0; b < B; b++) { const int np = cdiv(n_prompt[b] + n_steps, PK_PAGE); for (int i = 0; i < np; i++) hpt[(size_t) b * pk_max_pages + i] = next++; for (int i = 0; i < n_prompt[b]; i++) hsrc[(size_t) r++] = b * Tst + i; } }
        std::vector<PkOp> ops;
        auto seg = [&](const ArW & W, int N, int epi, int pair, int n_units) { PkSeg sg; memset(&sg, 0, sizeof sg); sg.W = (const __half *) W.p; sg.Wp = sg.W; sg.Ws = (const __half *) W.scales; sg.Wps = sg.Ws; sg.N = N; sg.epi = epi; sg.ldy = N; sg.pair = pair; sg.n_units = n_units; return sg; };
        auto gemv_op = [&](int layer, const float * X, const __half * X16, size_t xr, int K, const float * nw, std::initializer_list<PkSeg> segs) {
            PkOp op; memset(&op, 0, sizeof op);
            if (pk_q8) {                                        // first the quantisation of this phase's input rows (with the RMSNorm, if any) ...
                op.kind = PK_QUANT; op.layer = layer; op.X = X; op.X16 = X16; op.xrep = 0; op.ldx = K; op.K = K; op.norm = nw ? PKN_RMS : PKN_NONE; op.nw = nw; op.eps = 1e-5f;
                op.QY = xq; op.QD = xqd; op.qrep = qrep; op.qdrep = qdrep;
                ops.push_back(op);
                memset(&op, 0, sizeof op);
                op.XQ = xq; op.XD = xqd; op.qrep = qrep; op.qdrep = qdrep; nw = nullptr;      // ... then the GEMV over the quantised rows
            }
            op.kind = PK_GEMV; op.layer = layer; op.X = X; op.X16 = X16; op.xrep = xr; op.ldx = K; op.K = K; op.norm = nw ? PKN_RMS : PKN_NONE; op.nw = nw; op.eps = 1e-5f; op.q8 = pk_q8 ? 1 : 0;
            int u = 0;
            for (const PkSeg & sg : segs) { op.seg[op.nseg] = sg; op.seg[op.nseg].unit0 = u; u += sg.n_units; op.nseg++; }
            op.n_units = u;
            ops.push_back(op);
        };
        { PkOp op; memset(&op, 0, sizeof op); op.kind = PK_ROWS; ops.push_back(op); }
        const int rope_units_q = heads * (head_dim / 16), rope_units_k = kv_heads * (head_dim / 16);      // a unit = 8 rows of a head's first half + their partners in the second half
        for (int l = 0; l < n_layers; l++) {
            const OrpheusLayer & L = layers[(size_t) l];
            PkSeg sq = seg(L.wq, H, PKE_ROPE_Q, PKP_ROPE, rope_units_q); sq.Y = q;
            PkSeg sk = seg(L.wk, KV, PKE_ROPE_K, PKP_ROPE, rope_units_k);
            PkSeg sv = seg(L.wv, KV, PKE_KV, PKP_NONE, KV / 8); sv.kv = 1;
            gemv_op(l, px, nullptr, xrep, H, L.in_norm, {sq, sk, sv}); ops.back().kv_prefetch = 1;
            { PkOp op; memset(&op, 0, sizeof op); op.kind = PK_ATTN; op.layer = l; op.q = q; op.out16 = att16; op.orep = xrep; op.scale = scale; ops.push_back(op); }
            PkSeg so = seg(L.wo, H, PKE_RES, PKP_NONE, H / 8); so.Y = pxn; so.res = px; so.yrep = xrep;                 // xn = attention + residual(x)
            gemv_op(l, nullptr, att16, xrep, H, nullptr, {so});
            PkSeg sg2 = seg(L.wgate, F, PKE_SWIGLU, PKP_SWIGLU, F / 8); sg2.Wp = (const __half *) L.wup.p; sg2.Wps = (const __half *) L.wup.scales; sg2.Y16 = g16; sg2.yrep = grep;
            gemv_op(l, pxn, nullptr, xrep, H, L.post_norm, {sg2});
            PkSeg sd = seg(L.wdown, H, PKE_RES, PKP_NONE, H / 8); sd.Y = px; sd.res = pxn; sd.yrep = xrep;               // x = mlp + residual(xn)
            gemv_op(l, nullptr, g16, grep, F, nullptr, {sd});
        }
        { PkSeg sh = seg(head, vocab, PKE_LOGITS, PKP_NONE, cdiv(vocab, 8)); sh.Y = logits; gemv_op(0, px, nullptr, xrep, H, out_norm, {sh}); }
        { PkOp op; memset(&op, 0, sizeof op); op.kind = PK_ARGMAX; ops.push_back(op); }
        B2_CUDA(cudaMemcpyAsync(page_table, hpt.data(), hpt.size() * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(row_src, hsrc.data(), hsrc.size() * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_ops, ops.data(), ops.size() * sizeof(PkOp), cudaMemcpyHostToDevice, st));
        PkParams Pk; memset(&Pk, 0, sizeof Pk);
        Pk.ops = d_ops; Pk.n_ops = (int) ops.size(); Pk.R = B; Pk.H = H; Pk.heads = heads; Pk.kv_heads = kv_heads; Pk.hd = head_dim; Pk.n_out = 1; Pk.vocab = vocab;
        Pk.model = PKM_ORPHEUS; Pk.ak = pk_ak; Pk.pos_off = 1; Pk.n_steps_total = n_steps; Pk.stop_token = stopping_token;
        Pk.embed = embed; Pk.rope_ff = rope_ff; Pk.rope_cs = rope_cs; Pk.theta_scale = theta_scale; Pk.amax_v = amax_v; Pk.amax_i = amax_i; Pk.amax_ch = pk_amax;
        Pk.bar = d_bar; Pk.d_step = d_step; Pk.first_pos = d_np; Pk.d_out = d_out; Pk.stopped = stopped; Pk.row_pos = pk_pos; Pk.x0 = px; Pk.x0rep = xrep;
        Pk.kv_pool = pool; Pk.kv_layer_bytes = pk_layer_bytes; Pk.page_table = page_table; Pk.max_pages = pk_max_pages;
        Pk.logits = logits; Pk.logits_all = logits_all;
        PkLaunch pkl;
        if (pk_configure(Pk, kv_f32, std::min(std::max(H, F), pk_ak), 0, Tmax, pkl)) { set_error("orpheus: the persistent decode kernel does not fit this shape (%d positions) in shared memory", Tmax); return 1; }
        pk_prof_begin(Pk, ops.size(), pk_grid, st);
        pk_kv_import(kv_f32, Kc, Vc, (size_t) B * Tst * KV, row_src, row_seq, row_pos, Pk, R0, n_layers, st);      // the prompt pass' K / V rows (fp32, compact GQA rows) into the pages
        B2_LAUNCH_CHECK(ctx);
        for (int s0 = 1; s0 < n_steps; s0 += exit_every) {
            if (stopped && s0 > 1) { const int a = all_stopped(); if (a < 0) return 1; if (a) break; }
            Pk.step_begin = s0; Pk.n_steps = std::min(exit_every, n_steps - s0);
            B2_CUDA(pk_launch(pkl, Pk, pk_grid, st));
            ctx->launches++; pdk_launches++; pdk_steps += (uint64_t) Pk.n_steps;
        }
        pk_prof_end(Pk, ops, pk_grid, st);
        if (out_logits)
            for (int s = 1; s < n_steps; s++)
                for (int b = 0; b < B; b++)
                    B2_CUDA(cudaMemcpyAsync(out_logits + ((size_t) b * n_steps + s) * vocab, logits_all + ((size_t) s * B + b) * vocab, (size_t) vocab * 4, cudaMemcpyDeviceToHost, st));
    }
    // B2TTS_AR_GRAPH=1: capture one decode step into a CUDA graph and replay it (see parler.cu); not used when every step's logits go to the host
    const char * ge = getenv("B2TTS_AR_GRAPH");
    if (use_pdk) {
    } else if (!(ge && ge[0] == '0') && !out_logits && n_steps > 2) {      // on by default since it reproduced the reference's tokens on a B200 (round 2); B2TTS_AR_GRAPH=0 for A/B runs
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        if (run_decode()) return 1;                             // step 1 runs directly: every kernel instantiation has its attributes set before the capture
        const uint64_t l0 = ctx->launches;
        B2_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = run_decode();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (!rc) set_error("orpheus: stream capture failed: %s", cudaGetErrorString(ce)); return 1; }
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); set_error("orpheus: cudaGraphInstantiate failed"); return 1; }
        cudaError_t le = cudaSuccess;
        for (int s = 2; s < n_steps && le == cudaSuccess; s++) {
            le = cudaGraphLaunch(exec, st);
            if (stopped && le == cudaSuccess && (s + 1) % exit_every == 0 && s + 1 < n_steps) { const int a = all_stopped(); if (a < 0) le = cudaErrorUnknown; else if (a) break; }
        }
        ctx->launches += (uint64_t) (n_steps - 3) * (ctx->launches - l0);
        cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
        if (le != cudaSuccess) { set_error("orpheus: cudaGraphLaunch failed: %s", cudaGetErrorString(le)); return 1; }
    } else {
        for (int s = 1; s < n_steps; s++) {
            if (stopped && s % exit_every == 0) { const int a = all_stopped(); if (a < 0) return 1; if (a) break; }
            if (run_decode() || copy_logits(s)) return 1;
        }
    }
    B2_CUDA(cudaEventRecord(ev[1], st));
    B2_CUDA(cudaMemcpyAsync(out_tokens, d_out, (size_t) B * n_steps * 4, cudaMemcpyDeviceToHost, st));
    if (stopped) B2_CUDA(cudaMemcpyAsync(hflags.data(), stopped, (size_t) B * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (n_generated)
        for (int b = 0; b < B; b++) {                           // tokens past a sequence's stopping token are what the reference never computes: zeroed
            const int ng = hflags[(size_t) b] >= 0 ? hflags[(size_t) b] : n_steps;
            n_generated[b] = ng;
            for (int s2 = ng; s2 < n_steps; s2++) out_tokens[(size_t) b * n_steps + s2] = 0;
        }
    cudaEventElapsedTime(&timing_ms, ev[0], ev[1]);
    return 0;
}

}  // namespace b2
