// t5.cu -- T5 conditional-prompt encoder, first correct CUDA path.  See t5.h for what it replaces.
#include "t5.h"
#include "ar_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace b2 {

namespace {

// build_t5_norm (model.cpp:183-189): ggml_rms_norm with eps 1e-6 (float squares accumulated in a double, scale = 1/sqrtf(mean + eps)) x weight; a warp per row
__global__ void t5_rmsnorm_kernel(const float * __restrict__ x, const float * __restrict__ w, int H, int R, float * __restrict__ y) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float * row = x + (size_t) r * H;
    double s = 0.0;
    for (int c = lane; c < H; c += 32) s += (double) (row[c] * row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = (float) (s / (double) H);
    const float scale = 1.0f / sqrtf(mean + 1e-6f);
    for (int c = lane; c < H; c += 32) y[(size_t) r * H + c] = (row[c] * scale) * w[c];
}

// One query row x one head of the bidirectional self-attention (model.cpp:246-270): scores q.k (NO 1/sqrt(d): soft_max_ext scale 1.0) + the relative-position
// bias of (key - query), softmax over the row's whole prompt with ggml_soft_max's double-accumulated sum, then P.V.  q / k / v: [R][heads * hd] rows; a row
// attends to rows [row_base[r], row_base[r] + row_len[r]) (its own prompt) and sits at position row_pos[r] in it.  lut: [heads][2 * max_ctx - 1].
__global__ void __launch_bounds__(128) t5_attention_kernel(const float * __restrict__ q, const float * __restrict__ k, const float * __restrict__ v,
                                                           const int * __restrict__ row_base, const int * __restrict__ row_len, const int * __restrict__ row_pos,
                                                           const float * __restrict__ lut, int max_ctx, int heads, int hd, int Tcap, float * __restrict__ out) {
    extern __shared__ float t5_sc[];       // [Tcap] scores, then 128 floats + 128 doubles of reduction scratch behind them
    const int r = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    const size_t base = (size_t) row_base[r];
    const int T = row_len[r], pos = row_pos[r], H = heads * hd;
    float * redf = t5_sc + ((Tcap + 1) & ~1);
    double * redd = reinterpret_cast<double *>(redf + 128);
    const float * qv = q + (size_t) r * H + (size_t) h * hd;
    const float * bl = lut + (size_t) h * (2 * max_ctx - 1) + (max_ctx - 1) - pos;      // bl[t] = bias of key t for this query
    float mx = -INFINITY;
    for (int t = tid; t < T; t += 128) {
        const float * kr = k + (base + t) * H + (size_t) h * hd;
        float a = 0.f;
        for (int d = 0; d < hd; d += 4) {
            const float4 k4 = *reinterpret_cast<const float4 *>(kr + d);
            a = fmaf(qv[d + 3], k4.w, fmaf(qv[d + 2], k4.z, fmaf(qv[d + 1], k4.y, fmaf(qv[d], k4.x, a))));
        }
        a += bl[t];                                                                    // ggml_add(kq, pos_bias), model.cpp:260
        t5_sc[t] = a;
        mx = fmaxf(mx, a);
    }
    redf[tid] = mx;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (tid < o) redf[tid] = fmaxf(redf[tid], redf[tid + o]); __syncthreads(); }
    mx = redf[0];
    double sum = 0.0;
    for (int t = tid; t < T; t += 128) { const float e = expf(t5_sc[t] - mx); t5_sc[t] = e; sum += (double) e; }
    redd[tid] = sum;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (tid < o) redd[tid] += redd[tid + o]; __syncthreads(); }
    const float inv = (float) (1.0 / redd[0]);
    for (int d = tid; d < hd; d += 128) {
        float a = 0.f;
        for (int t = 0; t < T; t++) a = fmaf(t5_sc[t] * inv, v[(base + t) * H + (size_t) h * hd + d], a);
        out[(size_t) r * H + (size_t) h * hd + d] = a;
    }
}

// gelu(wi_0 x) * (wi_1 x) (model.cpp:278-279; ggml_gelu = the fp16 table), in place on the wi_0 branch
__global__ void t5_gated_gelu_kernel(float * __restrict__ up, const float * __restrict__ gate, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) up[i] = gelu_f16lut(up[i]) * gate[i];
}

// fp32 rows -> the fp16 operand of the tensor-core GEMM: the rounding ggml_mul_mat applies to the activations of an F16 matrix (vec_dot_type F16)
__global__ void t5_cast_h_kernel(const float * __restrict__ x, size_t n4, __half2 * __restrict__ y) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) { const float4 v = reinterpret_cast<const float4 *>(x)[i]; y[2 * i] = __floats2half2_rn(v.x, v.y); y[2 * i + 1] = __floats2half2_rn(v.z, v.w); }
}

__global__ void t5_add_bias_kernel(float * __restrict__ y, const float * __restrict__ b, int N, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += b[i % (size_t) N];
}

// t5_runner::set_inputs (model.cpp:318-332), one entry, with the reference's own arithmetic: ints divided inside the logarithm, a float denominator, double log
int t5_bucket(int key_pos, int query_pos, int relative_attn_buckets) {
    const int n_buckets = relative_attn_buckets / 2, max_exact = n_buckets / 2;
    const float logarithmic_denominator = (float) log(128.0 / max_exact);
    const int rpos = key_pos - query_pos, ab_rpos = abs(rpos);
    return (rpos > 0 ? n_buckets : 0) + (ab_rpos < max_exact ? ab_rpos : std::min(n_buckets - 1, max_exact + (int) ((log((double) (ab_rpos / max_exact)) / logarithmic_denominator) * max_exact)));
}

// From this many rows on an F16 matrix goes through the tensor-core GEMM (conv_gemm: the tcgen05 + TMA kernel of gemm_umma.cu as a K = 1 "convolution", the
// mma.sync kernel for shapes it does not take): ONE pass over the matrix for all rows, where the GEMV family -- built for decode batches -- streams it once per 64 rows.
constexpr int T5_GEMM_MIN_ROWS = 33;      // measured (profiles/r2x_t5.txt): 32 rows 2.66 ms GEMV vs 3.65 ms GEMM; 64 rows 4.49 vs 3.84 (the GEMV kernels run 33-64 rows as one 64-row tile)

struct TFwd : ArLaunch {
    T5 * m; bool fail = false, use_gemm = false;
    __half * xh = nullptr;                                     // [R][max(hidden, ffn)] fp16 operand scratch
    TFwd(T5 * m_, Ctx * c, cudaStream_t s) : m(m_) { ctx = c; st = s; }
    template <class T> T * al(size_t n) { T * p = (T *) m->arena.alloc(n * sizeof(T)); if (!p) fail = true; return p; }
    int rms(const float * x, const float * w, int H, int R, float * y) { t5_rmsnorm_kernel<<<cdiv(R, 8), 256, 0, st>>>(x, w, H, R, y); B2_LAUNCH_CHECK(ctx); return 0; }
    static bool gemm_shape(const ArW & W, int K, int N) { return W.f16 && !W.qtype && K % 64 == 0 && N % 128 == 0; }
    int cast(const float * x, int K, int R) {                  // xh = fp16(x), once per group of matrices that share the rows
        const size_t n4 = (size_t) R * K / 4;
        t5_cast_h_kernel<<<cdiv((int64_t) n4, 256), 256, 0, st>>>(x, n4, reinterpret_cast<__half2 *>(xh)); B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    // Y = X . W^T (+ res); `casted`: xh already holds fp16(X)
    int linear(const float * X, const ArW & W, int K, int N, int R, const float * res, float * Y, bool casted = false) {
#ifndef B2EMU
        if (use_gemm && gemm_shape(W, K, N)) {
            if (!casted && cast(X, K, R)) return 1;
            ConvGemmParams p;
            p.A = xh; p.lda = K; p.W = (const __half *) W.p; p.N = N; p.Npad = N; p.KW = 1; p.CinPad = K; p.CinTrue = K;
            p.B = 1; p.LmaxIn = R; p.LmaxOut = R; p.outF = Y; p.ldo = N; p.add1 = res; p.ldadd1 = N;
            return conv_gemm(ctx, p);
        }
#endif
        (void) casted;
        return gemv(X, K, W, K, N, R, res, Y, N);
    }
};

}  // namespace

int T5::assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {
    if (prepared) { set_error("t5: assign_weight after prepare"); return 1; }
    std::string nm(name);
    if (nm.rfind("t5encoder.", 0) == 0) nm = nm.substr(10);
    HostTensor t;
    if (host_tensor_from_gguf(t, name, type, n_dims, ne, data, nbytes, true)) return 1;
    host[nm] = std::move(t);
    return 0;
}

int T5::prepare() {
    if (prepared) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    // hyper-parameters: optional keys over the header's defaults, vocab_size required (t5_encoder::prep_constants, model.cpp:118-158)
    auto kvopt = [&](const char * k, int & out) { auto it = kv.find(k); if (it != kv.end()) out = (int) it->second; };
    kvopt("t5encoder.block_count", n_layers); kvopt("t5encoder.embedding_length", hidden); kvopt("t5encoder.attention.head_count", heads);
    kvopt("t5encoder.context_length", max_ctx); kvopt("tokenizer.ggml.eos_token_id", eos); kvopt("t5encoder.output_size", out_size);
    if (kv.find("t5encoder.vocab_size") == kv.end()) { set_error("key 't5encoder.vocab_size' must be specified in gguf file."); return 1; }
    vocab = (int) kv["t5encoder.vocab_size"];
    if (n_layers <= 0 || heads <= 0 || max_ctx <= 0 || heads * head_dim != hidden) { set_error("t5: hidden size %d is not %d heads of 64 (the reference fixes the head size, t5/model.h:46)", hidden, heads); return 1; }
    bool ok = true;
    auto find = [&](const std::string & n, int64_t expect, bool required = true) -> const HostTensor * {
        auto it = host.find(n);
        if (it == host.end()) { if (required) { set_error("missing tensor t5encoder.%s", n.c_str()); ok = false; } return nullptr; }
        if (expect && (int64_t) it->second.v.size() != expect) { set_error("tensor t5encoder.%s has %zu elements, expected %lld", n.c_str(), it->second.v.size(), (long long) expect); ok = false; return nullptr; }
        return &it->second;
    };
    auto dev = [&](const float * src, size_t n) -> float * {
        void * d = nullptr;
        if (cudaMalloc(&d, n * 4) != cudaSuccess) { cudaGetLastError(); set_error("t5: cudaMalloc of %zu bytes failed", n * 4); ok = false; return nullptr; }
        if (src) cudaMemcpy(d, src, n * 4, cudaMemcpyHostToDevice);
        dev_allocs.push_back(d); weight_bytes += n * 4;
        return (float *) d;
    };
    auto up = [&](const std::string & n, int64_t expect) -> float * { const HostTensor * t = find(n, expect); return t ? dev(t->v.data(), t->v.size()) : nullptr; };
    auto upw = [&](const std::string & n, int64_t expect) -> ArW {
        ArW w;
        const HostTensor * t = find(n, expect);
        if (!t) return w;
        if (t->qtype) { if (!upload_quant_planes(*t, w, dev_allocs, weight_bytes)) ok = false; return w; }
        w.f16 = t->f16;
        if (!t->f16) { w.p = dev(t->v.data(), t->v.size()); return w; }
        std::vector<__half> h(t->v.size());                                   // F16 tensors go to HBM as fp16 (their fp32 host copies are exact widenings)
        for (size_t i = 0; i < h.size(); i++) h[i] = __float2half_rn(t->v[i]);
        void * d = nullptr;
        if (cudaMalloc(&d, h.size() * 2) != cudaSuccess) { cudaGetLastError(); set_error("t5: cudaMalloc of %zu bytes failed", h.size() * 2); ok = false; return w; }
        cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
        dev_allocs.push_back(d); weight_bytes += h.size() * 2;
        w.p = d;
        return w;
    };

    embd = up("token_embd", (int64_t) vocab * hidden);
    out_norm = up("enc.final_layer_norm", hidden);
    { const HostTensor * t = find("enc.blk.0.ffn_up", 0); if (t && !t->shape.empty()) ffn = (int) t->shape[0]; }
    if (ok && (ffn <= 0 || ffn % 4 || hidden % 4)) { set_error("t5: feed-forward width %d / hidden %d must be multiples of 4", ffn, hidden); return 1; }
    layers.resize((size_t) n_layers);
    for (int l = 0; l < n_layers && ok; l++) {
        const std::string b = "enc.blk." + std::to_string(l);
        T5Layer & L = layers[(size_t) l];
        L.attn_norm = up(b + ".attn_norm", hidden); L.ffn_norm = up(b + ".ffn_norm", hidden);
        L.q = upw(b + ".attn_q", (int64_t) hidden * hidden); L.k = upw(b + ".attn_k", (int64_t) hidden * hidden);
        L.v = upw(b + ".attn_v", (int64_t) hidden * hidden); L.o = upw(b + ".attn_o", (int64_t) hidden * hidden);
        L.wi0 = upw(b + ".ffn_up", (int64_t) ffn * hidden); L.wi1 = upw(b + ".ffn_gate", (int64_t) ffn * hidden); L.wo = upw(b + ".ffn_down", (int64_t) hidden * ffn);
    }
    if (!ok) return 1;
    {   // the relative-position bias as a table over key - query: every layer adds layer 0's table (assign_to_t5_layer keeps ONE relative_attn_bias, model.cpp:57-60)
        const HostTensor * rb = nullptr;
        for (int l = 0; l < n_layers && !rb; l++) rb = find("enc.blk." + std::to_string(l) + ".attn_rel_b", 0, false);
        if (!rb || rb->shape.size() != 2 || rb->shape[1] != heads) { set_error("t5: missing or mis-shaped relative attention bias (expected [buckets][%d heads])", heads); return 1; }
        buckets = (int) rb->shape[0];
        if (buckets < 4 || buckets % 4) { set_error("t5: %d relative-attention buckets (a multiple of 4 expected)", buckets); return 1; }
        const int W = 2 * max_ctx - 1;
        std::vector<float> lut((size_t) heads * W);
        for (int d = -(max_ctx - 1); d <= max_ctx - 1; d++) {
            const int bk = t5_bucket(d, 0, buckets);
            for (int h = 0; h < heads; h++) lut[(size_t) h * W + (size_t) (d + max_ctx - 1)] = rb->v[(size_t) bk * heads + h];
        }
        bias_lut = dev(lut.data(), lut.size());
    }
    {
        const HostTensor * dp = find("down_proj", 0, false);
        if (dp) {
            if (dp->shape.size() != 2 || dp->shape[1] != hidden) { set_error("t5: down_proj must be [output_size][%d]", hidden); return 1; }
            out_size = (int) dp->shape[0];
            down = upw("down_proj", (int64_t) out_size * hidden); has_down = true;
            if (find("down_proj_bias", 0, false)) down_bias = up("down_proj_bias", out_size);
        }
    }
    if (!ok) return 1;
    B2_CUDA(cudaDeviceSynchronize());                         // the uploads above went through the legacy stream; the kernels run on ctx->stream (non-blocking)
    for (int i = 0; i < 2; i++) B2_CUDA(cudaEventCreate(&ev[i]));
    host.clear();
    prepared = true;
    return 0;
}

void T5::free_all() {
    for (void * p : dev_allocs) cudaFree(p);
    dev_allocs.clear();
    arena.release();
    for (int i = 0; i < 2; i++) if (ev[i]) cudaEventDestroy(ev[i]);
}

int T5::encode(int B, const uint32_t * const * tokens, const int32_t * n_tokens, float * out) {
    if (!prepared) { set_error("t5: model not prepared"); return 1; }
    if (B <= 0) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int R = 0, Tmax = 0;
    for (int b = 0; b < B; b++) {
        if (n_tokens[b] <= 0 || n_tokens[b] > max_ctx) { set_error("t5: prompt %d has %d tokens (1 .. %d, the model's context_length)", b, n_tokens[b], max_ctx); return 1; }
        for (int i = 0; i < n_tokens[b]; i++) if (tokens[b][i] >= (uint32_t) vocab) { set_error("t5: prompt %d token %u >= vocabulary %d", b, tokens[b][i], vocab); return 1; }
        R += n_tokens[b]; Tmax = std::max(Tmax, n_tokens[b]);
    }
    const int H = hidden, F = ffn, O = output_size();
    const size_t need = (size_t) R * ((size_t) 6 * H + 2 * (size_t) F + (size_t) O) * 4 + (size_t) R * std::max(H, F) * 2 + (size_t) R * 16 + (1 << 20);
    if (arena.reserve(need)) return 1;
    TFwd Fw(this, ctx, st);
    float * x = Fw.al<float>((size_t) R * H), * xn = Fw.al<float>((size_t) R * H), * q = Fw.al<float>((size_t) R * H), * k = Fw.al<float>((size_t) R * H),
          * v = Fw.al<float>((size_t) R * H), * att = Fw.al<float>((size_t) R * H), * g = Fw.al<float>((size_t) R * F), * u = Fw.al<float>((size_t) R * F),
          * y = Fw.al<float>((size_t) R * O);
    int * row_tok = Fw.al<int>((size_t) R), * row_base = Fw.al<int>((size_t) R), * row_len = Fw.al<int>((size_t) R), * row_pos = Fw.al<int>((size_t) R);
    Fw.xh = Fw.al<__half>((size_t) R * std::max(H, F));
    if (Fw.fail) return 1;
    { const char * e = getenv("B2TTS_T5_GEMM"); Fw.use_gemm = R >= T5_GEMM_MIN_ROWS && !(e && e[0] == '0'); }      // B2TTS_T5_GEMM=0: the GEMV family at every row count (A/B runs)
    last_used_gemm = false;
    {
        std::vector<int> ht((size_t) R), hb((size_t) R), hl((size_t) R), hp((size_t) R);
        int at = 0;
        for (int b = 0; b < B; b++) for (int i = 0; i < n_tokens[b]; i++) { ht[(size_t) at + i] = (int) tokens[b][i]; hb[(size_t) at + i] = at; hl[(size_t) at + i] = n_tokens[b]; hp[(size_t) at + i] = i; if (i == n_tokens[b] - 1) at += n_tokens[b]; }
        B2_CUDA(cudaMemcpyAsync(row_tok, ht.data(), (size_t) R * 4, cudaMemcpyHostToDevice, st)); B2_CUDA(cudaMemcpyAsync(row_base, hb.data(), (size_t) R * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(row_len, hl.data(), (size_t) R * 4, cudaMemcpyHostToDevice, st)); B2_CUDA(cudaMemcpyAsync(row_pos, hp.data(), (size_t) R * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaStreamSynchronize(st));                   // the host vectors are stack-owned
    }
    const size_t att_smem = (size_t) ((Tmax + 1) & ~1) * 4 + 128 * 4 + 128 * 8;
    if (att_smem > 200 * 1024) { set_error("t5: a prompt of %d tokens exceeds the attention kernel's shared memory", Tmax); return 1; }
    if (att_smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(t5_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) att_smem));
    B2_CUDA(cudaEventRecord(ev[0], st));
    embed_kernel<<<R, 256, 0, st>>>(row_tok, embd, H, x); B2_LAUNCH_CHECK(ctx);
    for (int l = 0; l < n_layers; l++) {
        const T5Layer & L = layers[(size_t) l];
        if (Fw.rms(x, L.attn_norm, H, R, xn)) return 1;
        const bool g3 = Fw.use_gemm && TFwd::gemm_shape(L.q, H, H) && TFwd::gemm_shape(L.k, H, H) && TFwd::gemm_shape(L.v, H, H);      // one cast for q, k and v
        if (g3) { if (Fw.cast(xn, H, R)) return 1; last_used_gemm = true; }
        if (Fw.linear(xn, L.q, H, H, R, nullptr, q, g3) || Fw.linear(xn, L.k, H, H, R, nullptr, k, g3) || Fw.linear(xn, L.v, H, H, R, nullptr, v, g3)) return 1;
        t5_attention_kernel<<<dim3((unsigned) R, (unsigned) heads), 128, att_smem, st>>>(q, k, v, row_base, row_len, row_pos, bias_lut, max_ctx, heads, head_dim, Tmax, att);
        B2_LAUNCH_CHECK(ctx);
        if (Fw.linear(att, L.o, H, H, R, x, xn)) return 1;                          // xn = attention output + residual(x)
        if (Fw.rms(xn, L.ffn_norm, H, R, x)) return 1;
        const bool g2 = Fw.use_gemm && TFwd::gemm_shape(L.wi0, H, F) && TFwd::gemm_shape(L.wi1, H, F);
        if (g2 && Fw.cast(x, H, R)) return 1;
        if (Fw.linear(x, L.wi0, H, F, R, nullptr, u, g2) || Fw.linear(x, L.wi1, H, F, R, nullptr, g, g2)) return 1;
        t5_gated_gelu_kernel<<<cdiv((int64_t) R * F, 256), 256, 0, st>>>(u, g, (size_t) R * F); B2_LAUNCH_CHECK(ctx);
        if (Fw.linear(u, L.wo, F, H, R, xn, x)) return 1;                            // x = mlp + residual(xn)
    }
    if (Fw.rms(x, out_norm, H, R, xn)) return 1;
    const float * res = xn;
    if (has_down) {
        if (Fw.linear(xn, down, H, O, R, nullptr, y)) return 1;
        if (down_bias) { t5_add_bias_kernel<<<cdiv((int64_t) R * O, 256), 256, 0, st>>>(y, down_bias, O, (size_t) R * O); B2_LAUNCH_CHECK(ctx); }
        res = y;
    }
    B2_CUDA(cudaEventRecord(ev[1], st));
    B2_CUDA(cudaMemcpyAsync(out, res, (size_t) R * O * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&timing_ms, ev[0], ev[1]);
    return 0;
}

}  // namespace b2
