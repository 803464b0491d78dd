// dac.cu -- DAC codec decoder (codebook indices -> PCM) on the B200.  See dac.h for what it replaces.
//
// Layout and kernels are the Kokoro path's: activations channels-last, batch-major, padded to the longest utterance with
// per-utterance length arrays; every convolution is the tcgen05 implicit GEMM (gemm_umma.cu); the ConvTranspose1d of each
// decoder block (K = 2 * stride) is one GEMM over all `stride` output phases (the polyphase form of kokoro.cu).  The reference keeps
// the codec's weights in F32 unless asked otherwise (examples/quantize/quantize_impl.cpp:44,265) and then computes every convolution
// in fp32, so here both GEMM operands are split into fp16 hi + lo parts and three products are accumulated in fp32 (error ~2^-22).
#include "dac.h"
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace b2 {

static inline float h2f_(uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); }

static int assign_host(std::map<std::string, HostTensor> & host, const char * prefix, const char * name, int type, int n_dims, const int64_t * ne,
                       const void * data, size_t nbytes) {
    std::string nm(name);
    if (nm.rfind(prefix, 0) == 0) nm = nm.substr(strlen(prefix));
    HostTensor t;
    int64_t n = 1;
    for (int i = n_dims - 1; i >= 0; i--) { t.shape.push_back(ne[i]); n *= ne[i]; }
    t.v.resize((size_t) n);
    if (type == 0) {
        if (nbytes < (size_t) n * 4) { set_error("tensor %s: short data", name); return 1; }
        memcpy(t.v.data(), data, (size_t) n * 4);
    } else if (type == 1) {
        if (nbytes < (size_t) n * 2) { set_error("tensor %s: short data", name); return 1; }
        const uint16_t * s = (const uint16_t *) data;
        for (int64_t i = 0; i < n; i++) t.v[(size_t) i] = h2f_(s[i]);
        t.f16 = true;
    } else {
        set_error("tensor %s: ggml type %d not supported (F32/F16 only)", name, type);
        return 1;
    }
    host[nm] = std::move(t);
    return 0;
}

int Dac::assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {
    if (prepared) { set_error("dac: assign_weight after prepare"); return 1; }
    return assign_host(host, "audio_encoder.", name, type, n_dims, ne, data, nbytes);
}
int Snac::assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {
    if (prepared) { set_error("snac: assign_weight after prepare"); return 1; }
    return assign_host(host, "snac.", name, type, n_dims, ne, data, nbytes);
}

namespace {

template <class M>
struct PrepT {
    M * m;
    bool ok = true;
    const HostTensor * get(const std::string & n) {
        auto it = m->host.find(n);
        if (it == m->host.end()) { set_error("missing tensor audio_encoder.%s", n.c_str()); ok = false; return nullptr; }
        return &it->second;
    }
    void * dev(const void * src, size_t bytes) {
        void * d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) failed for weights", bytes); ok = false; return nullptr; }
        cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
        m->dev_allocs.push_back(d);
        m->weight_bytes += bytes;
        return d;
    }
    float * f32(const std::string & n) { auto t = get(n); return t ? (float *) dev(t->v.data(), t->v.size() * 4) : nullptr; }
    float * f32v(const std::vector<float> & v) { return (float *) dev(v.data(), v.size() * 4); }
    // [N][Cin][K] (numpy order) -> fp16 [Npad][K][CinPad]
    W16 w16_from(const std::vector<float> & src, int N, int Cin, int K) {
        W16 w; w.N = N; w.Cin = Cin; w.KW = K; w.CinPad = round_up(Cin, 64); w.Npad = N > 64 ? round_up(N, 128) : 64;
        std::vector<__half> h((size_t) w.Npad * K * w.CinPad, __float2half(0.f));
        for (int n = 0; n < N; n++)
            for (int ci = 0; ci < Cin; ci++)
                for (int k = 0; k < K; k++) h[((size_t) n * K + k) * w.CinPad + ci] = __float2half(src[((size_t) n * Cin + ci) * K + k]);
        w.w = (__half *) dev(h.data(), h.size() * 2);
        return w;
    }
    // Conv1d kernel [Cout][Cin][K] (+ bias).  F32 storage -> split form over 3*Cin operand channels (x hi | x lo | x hi) x (W hi | W hi | W lo)
    DacConv conv(const std::string & base, int dil, int pad, int k_hint = 0) {
        DacConv c;
        auto t = get(base + ".weight");
        if (!t) return c;
        // a caller that hands over ggml tensors reports ggml_n_dims(), which drops outermost dimensions of size 1: the final [1][96][7] kernel arrives as [96][7].
        // Conv1d kernels are rank 3 in the file unless they are 1x1 projections stored as [Cout][Cin]: a rank-2 shape whose second extent is the kernel width of a
        // unit-output conv is told apart by the caller's `k_hint`
        std::vector<int64_t> sh = t->shape;
        if (k_hint > 1 && sh.size() == 2 && sh[1] == k_hint) sh.insert(sh.begin(), 1);
        c.Cout = (int) sh[0]; c.Cin = (int) sh[1]; c.K = sh.size() > 2 ? (int) sh[2] : 1; c.dil = dil; c.pad = pad;
        c.split = !t->f16;
        if (c.split) {
            const int C3 = 3 * c.Cin;
            std::vector<float> src((size_t) c.Cout * C3 * c.K);
            for (int n = 0; n < c.Cout; n++)
                for (int ci = 0; ci < c.Cin; ci++)
                    for (int k = 0; k < c.K; k++) {
                        const float wv = t->v[((size_t) n * c.Cin + ci) * c.K + k];
                        const float whi = __half2float(__float2half(wv)), wlo = wv - whi;
                        src[((size_t) n * C3 + ci) * c.K + k] = whi;
                        src[((size_t) n * C3 + c.Cin + ci) * c.K + k] = whi;
                        src[((size_t) n * C3 + 2 * c.Cin + ci) * c.K + k] = wlo;
                    }
            c.w = w16_from(src, c.Cout, C3, c.K);
            c.w.Cin = c.Cin;   // roofline accounting counts the conv's own products
        } else {
            c.w = w16_from(t->v, c.Cout, c.Cin, c.K);
        }
        c.b = f32(base + ".bias");
        return c;
    }
};

// ---------------------------------------------------------------- kernels
// quantizer: x[b][t][c] = sum over heads (in order) of table_h[code[b][t][h]][c]   (dac_build_audio_inputs, dac_model.cpp:100-123)
__global__ void dac_embed_kernel(const uint32_t * __restrict__ codes, const float * __restrict__ tables, int H, int n_codes, int latent, int Fmax,
                                 const int * __restrict__ len, float * __restrict__ out) {
    const int b = blockIdx.y, t = blockIdx.x;
    if (t >= len[b]) return;
    const uint32_t * cd = codes + ((size_t) b * Fmax + t) * H;
    float * o = out + ((size_t) b * Fmax + t) * latent;
    for (int c = threadIdx.x; c < latent; c += blockDim.x) {
        float acc = tables[((size_t) 0 * n_codes + cd[0]) * latent + c];
        for (int h = 1; h < H; h++) acc = acc + tables[((size_t) h * n_codes + cd[h]) * latent + c];
        o[c] = acc;
    }
}

// fp32 activations -> GEMM operand rows: optional snake (x + sin^2(alpha x) / alpha, src/util.cpp:86-101), then either fp16 (plain) or the
// split form [hi | lo | hi]; pad channels zero; the TAIL rows past an utterance's end zero (conv padding); rows = Lq per utterance
constexpr int DAC_TAIL = 32;
__global__ void dac_operand_kernel(const float * __restrict__ x, int ldx, int C, int Lmax, const int * __restrict__ len, const float * __restrict__ alpha,
                                   int split, __half * outH, int ldo, int Lq) {
    const int b = blockIdx.y;
    const int L = len[b];
    const int q = blockIdx.x * blockDim.y + threadIdx.y;
    if (q >= Lq || q >= L + DAC_TAIL) return;
    __half * orow = outH + ((size_t) b * Lq + q) * ldo;
    const int CW = split ? 3 * C : C;
    if (q >= L) { for (int c = threadIdx.x; c < ldo; c += blockDim.x) orow[c] = __float2half_rn(0.f); return; }
    const float * row = x + ((size_t) b * Lmax + q) * ldx;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v = row[c];
        if (alpha) { const float a = alpha[c]; const float s = sinf(v * a); v = v + (s * s) * (1.0f / a); }
        const __half hi = __float2half_rn(v);
        if (split) { const __half lo = __float2half_rn(v - __half2float(hi)); orow[c] = hi; orow[C + c] = lo; orow[2 * C + c] = hi; }
        else orow[c] = hi;
    }
    for (int c = CW + threadIdx.x; c < ldo; c += blockDim.x) orow[c] = __float2half_rn(0.f);
}

// 16-byte flavour (C % 4 == 0, 16-byte aligned rows): a thread converts 4 channels of a row (one float4 load, three 8-byte stores in
// split mode); blockDim = (64 channel groups, 4 rows), each thread walks the row's groups with stride 64 and 4 rows per block pass
__global__ void __launch_bounds__(256) dac_operand4_kernel(const float * __restrict__ x, int ldx, int C, int Lmax, const int * __restrict__ len,
                                                           const float * __restrict__ alpha, int split, __half * outH, int ldo, int Lq) {
    const int b = blockIdx.y;
    const int L = len[b];
    const int CW = split ? 3 * C : C;
    for (int u = 0; u < 4; u++) {
        const int q = (blockIdx.x * 4 + u) * 4 + threadIdx.y;
        if (q >= Lq || q >= L + DAC_TAIL) continue;
        __half * orow = outH + ((size_t) b * Lq + q) * ldo;
        if (q >= L) { for (int c = threadIdx.x * 4; c < ldo; c += 256) *reinterpret_cast<uint2 *>(orow + c) = make_uint2(0u, 0u); continue; }
        const float * row = x + ((size_t) b * Lmax + q) * ldx;
        for (int c = threadIdx.x * 4; c < C; c += 256) {
            const float4 v4 = *reinterpret_cast<const float4 *>(row + c);
            float v[4] = {v4.x, v4.y, v4.z, v4.w};
            if (alpha) {
                const float4 a4 = *reinterpret_cast<const float4 *>(alpha + c);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int k = 0; k < 4; k++) { const float s = sinf(v[k] * a[k]); v[k] = v[k] + (s * s) * (1.0f / a[k]); }
            }
            __half hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { hi[k] = __float2half_rn(v[k]); lo[k] = __float2half_rn(v[k] - __half2float(hi[k])); }
            uint2 ph, pl;
            ph.x = (uint32_t) __half_as_ushort(hi[0]) | ((uint32_t) __half_as_ushort(hi[1]) << 16);
            ph.y = (uint32_t) __half_as_ushort(hi[2]) | ((uint32_t) __half_as_ushort(hi[3]) << 16);
            *reinterpret_cast<uint2 *>(orow + c) = ph;
            if (split) {
                pl.x = (uint32_t) __half_as_ushort(lo[0]) | ((uint32_t) __half_as_ushort(lo[1]) << 16);
                pl.y = (uint32_t) __half_as_ushort(lo[2]) | ((uint32_t) __half_as_ushort(lo[3]) << 16);
                *reinterpret_cast<uint2 *>(orow + C + c) = pl;
                *reinterpret_cast<uint2 *>(orow + 2 * C + c) = ph;
            }
        }
        for (int c = CW + threadIdx.x * 4; c < ldo; c += 256) *reinterpret_cast<uint2 *>(orow + c) = make_uint2(0u, 0u);
    }
}

template <class M>
struct FwdT {
    M * m; Ctx * ctx; int B; bool fail = false;
    template <class T> T * al(size_t n) { T * p = (T *) m->arena.alloc(n * sizeof(T)); if (!p) fail = true; return p; }

    int operand(const float * x, int ldx, int C, int Lmax, const int * len, const float * alpha, bool split, __half * out, int ldo, int Lq) {
        if (C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ((((uintptr_t) x) | ((uintptr_t) alpha)) & 15) == 0 && (((uintptr_t) out) & 7) == 0 &&
            !getenv("B2TTS_DAC_SCALAR_OPERAND")) {
            dim3 blk4(64, 4), grid4(cdiv(Lq, 16), B);
            dac_operand4_kernel<<<grid4, blk4, 0, ctx->stream>>>(x, ldx, C, Lmax, len, alpha, split ? 1 : 0, out, ldo, Lq);
            B2_LAUNCH_CHECK(ctx);
            return 0;
        }
        dim3 blk(128, 2), grid(cdiv(Lq, 2), B);
        dac_operand_kernel<<<grid, blk, 0, ctx->stream>>>(x, ldx, C, Lmax, len, alpha, split ? 1 : 0, out, ldo, Lq);
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    int conv(const DacConv & c, const __half * a16, int L, const int * len, float * out, const float * add1, int act = ACT_NONE) {
        ConvGemmParams p;
        p.A = a16; p.lda = c.w.CinPad; p.W = c.w.w; p.bias = c.b; p.outF = out; p.ldo = c.Cout; p.add1 = add1; p.ldadd1 = c.Cout; p.act = act;
        p.B = B; p.LmaxIn = L; p.LmaxOut = L; p.lenIn = len; p.lenOut = len;
        p.N = c.w.N; p.Npad = c.w.Npad; p.KW = c.w.KW; p.CinPad = c.w.CinPad; p.CinTrue = c.w.Cin; p.stride = 1; p.dil = c.dil; p.pad = c.pad;
        p.tailClean = true;
        return conv_gemm(ctx, p);
    }
    // x [B][L][C] fp32 -> (snake) -> operand -> conv
    int snake_conv(const DacConv & c, const float * alpha, const float * x, int L, const int * len, float * out, const float * add1, int act = ACT_NONE) {
        const size_t mark = m->arena.off;            // the operand is a temporary: later launches are stream-ordered after its consumer
        __half * a16 = al<__half>((size_t) B * L * c.w.CinPad);
        if (fail) return 1;
        if (operand(x, c.Cin, c.Cin, L, len, alpha, c.split, a16, c.w.CinPad, L)) return 1;
        const int rc = conv(c, a16, L, len, out, add1, act);
        m->arena.off = mark;
        return rc;
    }
};

}  // namespace

int Dac::prepare() {
    if (prepared) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    PrepT<Dac> P{this};
    auto kvget = [&](std::initializer_list<const char *> keys, uint32_t dflt) { for (auto k : keys) { auto it = kv.find(k); if (it != kv.end()) return it->second; } return dflt; };
    n_heads = (int) kvget({"parler-tts.decoder.output_heads", "output_heads", "dia.decoder.output_heads"}, 9);     // dac_model.cpp:15-18
    up_factor = (int) kvget({"dac.up_sampling_factor", "up_sampling_factor"}, 512);                               // dac_model.cpp:20-23

    // quantizer tables: out_proj (1x1 conv over the 8-dim codebook row) + bias, one table per head.  The F32 mul_mat accumulates its
    // (fewer than one SIMD step of) products in a double (ggml_vec_dot_f32 leftovers), reproduced here.
    {
        auto cb0 = P.get("quantizers.0.codebook.weight");
        if (!cb0) return 1;
        n_codes = (int) cb0->shape[0]; const int cbd = (int) cb0->shape[1];
        auto w0 = P.get("quantizers.0.out_proj.weight");
        if (!w0) return 1;
        latent = (int) w0->shape[0];
        std::vector<float> tab((size_t) n_heads * n_codes * latent);
        for (int h = 0; h < n_heads; h++) {
            auto cb = P.get("quantizers." + std::to_string(h) + ".codebook.weight"), w = P.get("quantizers." + std::to_string(h) + ".out_proj.weight"),
                 bs = P.get("quantizers." + std::to_string(h) + ".out_proj.bias");
            if (!cb || !w || !bs) return 1;
            const bool r16 = w->f16;   // an F16 projection kernel re-rounds the codebook rows to fp16 (ggml-cpu.c:262-267)
            for (int code = 0; code < n_codes; code++)
                for (int c = 0; c < latent; c++) {
                    double s = 0.0;
                    for (int d = 0; d < cbd; d++) {
                        float xv = cb->v[(size_t) code * cbd + d];
                        if (r16) xv = __half2float(__float2half(xv));
                        s += (double) (xv * w->v[(size_t) c * cbd + d]);
                    }
                    tab[((size_t) h * n_codes + code) * latent + c] = (float) s + bs->v[c];
                }
        }
        tables = P.f32v(tab);
    }
    initial = P.conv("initial", 1, 3);
    for (int l = 0; l < 4; l++) {
        DacLayer & L = layers[l];
        const std::string b = "decoder_block." + std::to_string(l + 1);
        auto sk = kvget({("dac.dac_layer_stride_" + std::to_string(l)).c_str(), ("dac_layer_stride_" + std::to_string(l)).c_str()}, 0);
        auto pk = kvget({("dac.dac_layer_padding_" + std::to_string(l)).c_str(), ("dac_layer_padding_" + std::to_string(l)).c_str()}, 0xffffffffu);
        if (sk == 0 || pk == 0xffffffffu) { set_error("key dac_layer_stride_%d / dac_layer_padding_%d must be specified in gguf file inorder to initialize the DAC audio decoder.", l, l); return 1; }
        L.stride = (int) sk; L.pad = (int) pk;
        L.alpha = P.f32(b + ".final.alpha");
        auto t = P.get(b + ".final.weight"), bt = P.get(b + ".final.bias");
        if (!t || !bt) return 1;
        if (t->f16) { set_error("%s.final.weight: F16 ConvTranspose1d kernels are not supported (the reference's F16 ConvTranspose path is mis-indexed, ggml-cpu.c:10091); keep them F32", b.c_str()); return 1; }
        L.Cin = (int) t->shape[0]; L.Cout = (int) t->shape[1];
        const int K = (int) t->shape[2], s = L.stride;
        if (K != 2 * s || L.pad >= s) { set_error("%s: ConvTranspose1d with K=%d stride=%d pad=%d is outside the polyphase form (K == 2*stride, pad < stride)", b.c_str(), K, s, L.pad); return 1; }
        // out[q*s + r - p] = x[q] . W[:, :, r] + x[q-1] . W[:, :, r+s]  (see kokoro.cu): N = s*Cout, 2 taps, 3*Cin split channels
        const int C3 = 3 * L.Cin, N = s * L.Cout;
        std::vector<float> src((size_t) N * C3 * 2), brep((size_t) N);
        for (int r = 0; r < s; r++)
            for (int co = 0; co < L.Cout; co++) {
                const size_t n = (size_t) r * L.Cout + co;
                brep[n] = bt->v[co];
                for (int ci = 0; ci < L.Cin; ci++)
                    for (int k = 0; k < 2; k++) {
                        const float wv = t->v[((size_t) ci * L.Cout + co) * K + (k == 0 ? r + s : r)];   // tap 0 reads x[q-1]
                        const float whi = __half2float(__float2half(wv)), wlo = wv - whi;
                        src[(n * C3 + ci) * 2 + k] = whi;
                        src[(n * C3 + L.Cin + ci) * 2 + k] = whi;
                        src[(n * C3 + 2 * L.Cin + ci) * 2 + k] = wlo;
                    }
            }
        L.w3 = P.w16_from(src, N, C3, 2);
        L.w3.Cin = L.Cin;
        L.b_rep = P.f32v(brep);
        for (int i = 0; i < 3; i++) {
            const std::string r = b + ".residual_unit." + std::to_string(i) + ".res";
            const int d = (int) std::lround(std::pow(3.0, i));          // general_neural_audio_codec.h:44-48: dilation 3^i, padding 3^(i+1)
            L.res[i].a1 = P.f32(r + ".initial.alpha"); L.res[i].c1 = P.conv(r + ".initial", d, 3 * d);
            L.res[i].a2 = P.f32(r + ".final.alpha");   L.res[i].c2 = P.conv(r + ".final", 1, 0);
        }
    }
    final_alpha = P.f32("final.alpha");
    final_conv = P.conv("final", 1, 3, 7);
    if (!P.ok) return 1;
    for (int i = 0; i < 2; i++) B2_CUDA(cudaEventCreate(&ev[i]));
    host.clear();
    cudaDeviceSynchronize();               // legacy-stream uploads above vs kernels on the non-blocking ctx->stream
    prepared = true;
    return 0;
}

void Dac::free_all() {
    for (void * p : dev_allocs) cudaFree(p);
    dev_allocs.clear();
    arena.release();
    if (pcm_pinned) cudaFreeHost(pcm_pinned);
    for (int i = 0; i < 2; i++) if (ev[i]) cudaEventDestroy(ev[i]);
}

int Dac::decode_batch(int B, const uint32_t * const * codes, const int32_t * frames, const float ** pcm, int64_t * n_samples) {
    if (!prepared) { set_error("dac: model not prepared"); return 1; }
    if (B <= 0) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int Fmax = 0;
    for (int b = 0; b < B; b++) {
        if (frames[b] <= 0) { set_error("dac: utterance %d has %d frames", b, frames[b]); return 1; }
        Fmax = std::max(Fmax, (int) frames[b]);
    }
    // row pitch per level: the polyphase ConvTranspose GEMM writes (L + 1) * stride rows per utterance, the level's tensors share that pitch
    int P[5]; P[0] = Fmax;
    for (int l = 0; l < 4; l++) P[l + 1] = (P[l] + 1) * layers[l].stride;
    size_t need = (size_t) B * Fmax * (n_heads * 4 + latent * 4 + initial.w.CinPad * 2 + initial.Cout * 4) + (64 << 20);
    for (int l = 0; l < 4; l++) {
        const DacLayer & L = layers[l];
        need += (size_t) B * (P[l] + 1) * L.w3.CinPad * 2;                                   // split operand of the ConvTranspose
        need += ((size_t) B * P[l + 1] + 64) * L.Cout * 4 * 3;                               // stream, branch, ping-pong
        need += (size_t) B * P[l + 1] * (size_t) std::max(L.res[0].c1.w.CinPad, L.res[0].c2.w.CinPad) * 2 * 2 + (4 << 20);
    }
    need += (size_t) B * P[4] * (final_conv.w.CinPad * 2 + 4) + (16 << 20);
    if (arena.reserve(need)) return 1;
    FwdT<Dac> F{this, ctx, B};

    // ---- inputs
    std::vector<uint32_t> hc((size_t) B * Fmax * n_heads, 0u);
    std::vector<int> hl((size_t) 9 * B);
    for (int b = 0; b < B; b++) {
        for (int64_t i = 0; i < (int64_t) frames[b] * n_heads; i++) {
            if (codes[b][i] >= (uint32_t) n_codes) { set_error("dac: utterance %d code %u >= codebook size %d", b, codes[b][i], n_codes); return 1; }
            hc[(size_t) b * Fmax * n_heads + i] = codes[b][i];
        }
        int L = frames[b];
        for (int l = 0; l <= 4; l++) { hl[(size_t) l * B + b] = L; if (l < 4) { hl[(size_t) (5 + l) * B + b] = L + 1; L *= layers[l].stride; } }
    }
    uint32_t * d_codes = F.al<uint32_t>(hc.size());
    int * d_len = F.al<int>(hl.size());
    if (F.fail) return 1;
    B2_CUDA(cudaEventRecord(ev[0], st));
    B2_CUDA(cudaMemcpyAsync(d_codes, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(d_len, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaStreamSynchronize(st));   // hc / hl are stack-owned

    // ---- quantizer + initial conv (dac_model.cpp:156-158)
    float * emb = F.al<float>((size_t) B * Fmax * latent);
    float * h = F.al<float>((size_t) B * Fmax * initial.Cout);
    if (F.fail) return 1;
    {
        dim3 grid(Fmax, B);
        dac_embed_kernel<<<grid, 256, 0, st>>>(d_codes, tables, n_heads, n_codes, latent, Fmax, d_len, emb);
        B2_LAUNCH_CHECK(ctx);
    }
    if (F.snake_conv(initial, nullptr, emb, Fmax, d_len, h, nullptr)) return 1;

    // ---- decoder blocks (general_neural_audio_codec::build_layer / build_residual_unit)
    const float * x = h;
    for (int l = 0; l < 4; l++) {
        const DacLayer & L = layers[l];
        const int Lq = P[l] + 1, Pn = P[l + 1], C = L.Cout;
        const int * len_in = d_len + (size_t) l * B, * len_q = d_len + (size_t) (5 + l) * B, * len_out = d_len + (size_t) (l + 1) * B;
        __half * a3 = F.al<__half>((size_t) B * Lq * L.w3.CinPad);
        float * ubuf = F.al<float>(((size_t) B * Pn + 64) * C);
        float * y = F.al<float>((size_t) B * Pn * C);
        float * alt = F.al<float>((size_t) B * Pn * C);
        if (F.fail) return 1;
        // snake -> ConvTranspose1d as ONE GEMM over the `stride` output phases: row q of the result holds out[q*s - pad .. q*s - pad + s)
        if (F.operand(x, L.Cin, L.Cin, P[l], len_in, L.alpha, true, a3, L.w3.CinPad, Lq)) return 1;
        {
            ConvGemmParams p;
            p.A = a3; p.lda = L.w3.CinPad; p.W = L.w3.w; p.bias = L.b_rep; p.outF = ubuf; p.ldo = L.stride * C;
            p.B = B; p.LmaxIn = Lq; p.LmaxOut = Lq; p.lenIn = len_in; p.lenOut = len_q;
            p.N = L.w3.N; p.Npad = L.w3.Npad; p.KW = 2; p.CinPad = L.w3.CinPad; p.CinTrue = L.w3.Cin; p.stride = 1; p.dil = 1; p.pad = 1;
            p.tailClean = true;
            if (conv_gemm(ctx, p)) return 1;
        }
        float * cur = ubuf + (size_t) L.pad * C;       // out[o] = Y[o + pad], viewed with a pitch of Pn rows of C channels per utterance
        for (int i = 0; i < 3; i++) {
            const DacUnit & U = L.res[i];
            float * nxt = (cur == alt) ? ubuf : alt;
            if (F.snake_conv(U.c1, U.a1, cur, Pn, len_out, y, nullptr)) return 1;
            if (F.snake_conv(U.c2, U.a2, y, Pn, len_out, nxt, cur)) return 1;
            cur = nxt;
        }
        x = cur;
    }
    // ---- snake -> conv k7 -> tanh (dac_model.cpp:162-165)
    float * pcm_d = F.al<float>((size_t) B * P[4]);
    if (F.fail) return 1;
    if (F.snake_conv(final_conv, final_alpha, x, P[4], d_len + (size_t) 4 * B, pcm_d, nullptr, ACT_TANH)) return 1;
    B2_CUDA(cudaEventRecord(ev[1], st));

    // ---- D2H into the runner-owned pinned buffer
    size_t total = 0;
    for (int b = 0; b < B; b++) total += (size_t) frames[b] * up_factor;
    if (pcm_pinned_cap < total) {
        if (pcm_pinned) cudaFreeHost(pcm_pinned);
        pcm_pinned = nullptr; pcm_pinned_cap = 0;
        B2_CUDA(cudaMallocHost(&pcm_pinned, total * 4));
        pcm_pinned_cap = total;
    }
    size_t off = 0;
    for (int b = 0; b < B; b++) {
        const size_t n = (size_t) frames[b] * up_factor;
        if ((int64_t) n > (int64_t) P[4]) { set_error("dac: up_sampling_factor %d does not match the layer strides", up_factor); return 1; }
        B2_CUDA(cudaMemcpyAsync(pcm_pinned + off, pcm_d + (size_t) b * P[4], n * 4, cudaMemcpyDeviceToHost, st));
        if (pcm) pcm[b] = pcm_pinned + off;
        if (n_samples) n_samples[b] = (int64_t) n;
        off += n;
    }
    B2_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&timing_ms, ev[0], ev[1]);
    return 0;
}

}  // namespace b2

// ============================================================================================ SNAC
#include <random>

namespace b2 {
namespace {

// the reference's static generator (src/util.cpp:74-80): libstdc++'s own engine and distribution, so the stream is identical by construction
struct NormalGen {
    std::default_random_engine e;
    std::normal_distribution<float> dis{0.0f, 1.0f};
};

// quantizer: x[b][t][c] = T0[c0[t/4]][c] + T1[c1[t/2]][c] + T2[c2[t]][c]   (snac_build_audio_inputs, snac_model.cpp:86-109)
__global__ void snac_embed_kernel(const uint32_t * __restrict__ codes, const int * __restrict__ code_off, const float * __restrict__ tables, int n_codes,
                                  int latent, int Lmax, const int * __restrict__ len, float * __restrict__ out) {
    const int b = blockIdx.y, t = blockIdx.x;
    const int L = len[b];
    if (t >= L) return;
    const uint32_t * cd = codes + code_off[b];
    const uint32_t c0 = cd[t >> 2], c1 = cd[(L >> 2) + (t >> 1)], c2 = cd[(L >> 2) + (L >> 1) + t];
    float * o = out + ((size_t) b * Lmax + t) * latent;
    for (int c = threadIdx.x; c < latent; c += blockDim.x) {
        float acc = tables[((size_t) 0 * n_codes + c0) * latent + c];
        acc = acc + tables[((size_t) 1 * n_codes + c1) * latent + c];
        acc = acc + tables[((size_t) 2 * n_codes + c2) * latent + c];
        o[c] = acc;
    }
}

// depthwise Conv1d k7 (ggml_conv_1d_dw: F32 im2col + mul_mat, the 7-tap dot accumulated in a double like ggml_vec_dot_f32's leftovers,
// ggml.c:3847-3868), optional snake on the input, + bias.  x, y: [b][t][C]
__global__ void snac_dwconv7_kernel(const float * __restrict__ x, int C, int Lmax, const int * __restrict__ len, const float * __restrict__ alpha,
                                    const float * __restrict__ w, const float * __restrict__ bias, int dil, float * __restrict__ y) {
    const int b = blockIdx.y;
    const int L = len[b];
    const int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= L) return;
    const float * xb = x + (size_t) b * Lmax * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float a = alpha ? alpha[c] : 0.f;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int tt = t + (k - 3) * dil;
            float v = 0.f;
            if (tt >= 0 && tt < L) {
                v = xb[(size_t) tt * C + c];
                if (alpha) { const float sn = sinf(v * a); v = v + (sn * sn) * (1.0f / a); }
            }
            s += (double) (v * w[c * 7 + k]);
        }
        y[((size_t) b * Lmax + t) * C + c] = (float) s + bias[c];
    }
}

// noise block: x[b][t][c] += nx[b][t][c] * noise[b][t]   (general_neural_audio_codec.cpp:153-157)
__global__ void snac_noise_add_kernel(float * x, const float * __restrict__ nx, const float * __restrict__ noise, int C, int Lmax, const int * __restrict__ len) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= len[b]) return;
    const float n = noise[(size_t) b * Lmax + t];
    const size_t row = ((size_t) b * Lmax + t) * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) x[row + c] = x[row + c] + nx[row + c] * n;
}

}  // namespace

int Snac::prepare() {
    if (prepared) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    PrepT<Snac> P{this};
    auto kvget = [&](const std::string & k, uint32_t dflt) { auto it = kv.find(k); return it != kv.end() ? it->second : dflt; };
    up_factor = (int) kvget("snac.up_sampling_factor", 512);
    if (kvget("snac.audio_token_channels", 3) != 3) { set_error("snac: audio_token_channels must be 3"); return 1; }
    {
        auto cb0 = P.get("quantizers.0.codebook.weight"), w0 = P.get("quantizers.0.out_proj.weight");
        if (!cb0 || !w0) return 1;
        n_codes = (int) cb0->shape[0]; const int cbd = (int) cb0->shape[1]; latent = (int) w0->shape[0];
        std::vector<float> tab((size_t) 3 * n_codes * latent);
        for (int h = 0; h < 3; h++) {
            auto cb = P.get("quantizers." + std::to_string(h) + ".codebook.weight"), w = P.get("quantizers." + std::to_string(h) + ".out_proj.weight"),
                 bs = P.get("quantizers." + std::to_string(h) + ".out_proj.bias");
            if (!cb || !w || !bs) return 1;
            for (int code = 0; code < n_codes; code++)
                for (int c = 0; c < latent; c++) {
                    double s = 0.0;
                    for (int d = 0; d < cbd; d++) s += (double) (cb->v[(size_t) code * cbd + d] * w->v[(size_t) c * cbd + d]);
                    tab[((size_t) h * n_codes + code) * latent + c] = (float) s + bs->v[c];
                }
        }
        tables = P.f32v(tab);
    }
    { auto t = P.get("in.weight"); if (!t) return 1; if (t->shape.size() != 3 || t->shape[1] != 1 || t->shape[2] != 7) { set_error("snac.in.weight must be a depthwise k7 kernel"); return 1; } }
    in_w = P.f32("in.weight"); in_b = P.f32("in.bias");
    up = P.conv("up", 1, 0);
    for (int l = 0; l < 4; l++) {
        SnacLayer & L = layers[l];
        const std::string b = "layers." + std::to_string(l);
        const uint32_t sk = kvget("snac.snac_layer_stride_" + std::to_string(l), 0), pk = kvget("snac.snac_layer_padding_" + std::to_string(l), 0xffffffffu);
        if (sk == 0 || pk == 0xffffffffu) { set_error("key snac.snac_layer_stride_%d must be specified in gguf file inorder to initialize the SNAC audio decoder.", l); return 1; }
        L.stride = (int) sk; L.pad = (int) pk;
        L.alpha = P.f32(b + ".alpha");
        auto t = P.get(b + ".weight"), bt = P.get(b + ".bias");
        if (!t || !bt) return 1;
        if (t->f16) { set_error("snac %s.weight: F16 ConvTranspose1d kernels are not supported; keep them F32", b.c_str()); return 1; }
        L.Cin = (int) t->shape[0]; L.Cout = (int) t->shape[1];
        const int K = (int) t->shape[2], s = L.stride;
        if (K != 2 * s || L.pad >= s) { set_error("snac %s: ConvTranspose1d K=%d stride=%d pad=%d is outside the polyphase form", b.c_str(), K, s, L.pad); return 1; }
        const int C3 = 3 * L.Cin, N = s * L.Cout;
        std::vector<float> src((size_t) N * C3 * 2), brep((size_t) N);
        for (int r = 0; r < s; r++)
            for (int co = 0; co < L.Cout; co++) {
                const size_t n = (size_t) r * L.Cout + co;
                brep[n] = bt->v[co];
                for (int ci = 0; ci < L.Cin; ci++)
                    for (int k = 0; k < 2; k++) {
                        const float wv = t->v[((size_t) ci * L.Cout + co) * K + (k == 0 ? r + s : r)];
                        const float whi = __half2float(__float2half(wv)), wlo = wv - whi;
                        src[(n * C3 + ci) * 2 + k] = whi;
                        src[(n * C3 + L.Cin + ci) * 2 + k] = whi;
                        src[(n * C3 + 2 * L.Cin + ci) * 2 + k] = wlo;
                    }
            }
        L.w3 = P.w16_from(src, N, C3, 2);
        L.w3.Cin = L.Cin;
        L.b_rep = P.f32v(brep);
        {   // noise block kernel: 1x1, no bias
            auto nt = P.get(b + ".noise_weight");
            if (!nt) return 1;
            host[b + ".noise.weight"] = *nt;
            HostTensor zb; zb.shape = {(int64_t) L.Cout}; zb.v.assign((size_t) L.Cout, 0.f);
            host[b + ".noise.bias"] = zb;
            L.noise = P.conv(b + ".noise", 1, 0);
            L.noise.b = nullptr;
        }
        for (int i = 0; i < 3; i++) {
            const std::string r = b + ".residual_unit." + std::to_string(i) + ".res";
            auto dw = P.get(r + ".initial.weight");
            if (!dw) return 1;
            if (dw->shape.size() != 3 || dw->shape[1] != 1 || dw->shape[2] != 7) { set_error("snac %s.initial.weight must be a depthwise k7 kernel (grouping == channels)", r.c_str()); return 1; }
            L.res[i].dil = (int) std::lround(std::pow(3.0, i));
            L.res[i].a1 = P.f32(r + ".initial.alpha"); L.res[i].dw_w = P.f32(r + ".initial.weight"); L.res[i].dw_b = P.f32(r + ".initial.bias");
            L.res[i].a2 = P.f32(r + ".final.alpha");   L.res[i].c2 = P.conv(r + ".final", 1, 0);
        }
    }
    final_alpha = P.f32("alpha_out");
    final_conv = P.conv("final", 1, 3, 7);
    if (!P.ok) return 1;
    for (int i = 0; i < 2; i++) B2_CUDA(cudaEventCreate(&ev[i]));
    noise_engine = new NormalGen();
    host.clear();
    cudaDeviceSynchronize();               // legacy-stream uploads above vs kernels on the non-blocking ctx->stream
    prepared = true;
    return 0;
}

void Snac::reset_noise() { if (noise_engine) { delete (NormalGen *) noise_engine; noise_engine = new NormalGen(); } }

void Snac::free_all() {
    for (void * p : dev_allocs) cudaFree(p);
    dev_allocs.clear();
    arena.release();
    if (pcm_pinned) cudaFreeHost(pcm_pinned);
    if (noise_engine) { delete (NormalGen *) noise_engine; noise_engine = nullptr; }
    for (int i = 0; i < 2; i++) if (ev[i]) cudaEventDestroy(ev[i]);
}

int Snac::decode_batch(int B, const uint32_t * const * codes, const int32_t * fine_frames, const float ** pcm, int64_t * n_samples) {
    if (!prepared) { set_error("snac: model not prepared"); return 1; }
    if (B <= 0) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int Fmax = 0; size_t n_codes_total = 0;
    for (int b = 0; b < B; b++) {
        if (fine_frames[b] <= 0 || fine_frames[b] % 4) { set_error("snac: utterance %d has %d fine frames (must be a positive multiple of 4)", b, fine_frames[b]); return 1; }
        Fmax = std::max(Fmax, (int) fine_frames[b]);
        n_codes_total += (size_t) fine_frames[b] / 4 * 7;
    }
    int P[5]; P[0] = Fmax;
    for (int l = 0; l < 4; l++) P[l + 1] = (P[l] + 1) * layers[l].stride;
    size_t need = n_codes_total * 4 + (size_t) B * Fmax * ((size_t) latent * 8 + up.w.CinPad * 2 + up.Cout * 4) + (64 << 20);
    for (int l = 0; l < 4; l++) {
        const SnacLayer & L = layers[l];
        need += (size_t) B * (P[l] + 1) * L.w3.CinPad * 2;
        need += ((size_t) B * P[l + 1] + 64) * L.Cout * 4 * 4 + (size_t) B * P[l + 1] * 4;
        need += (size_t) B * P[l + 1] * (size_t) std::max(L.noise.w.CinPad, L.res[0].c2.w.CinPad) * 2 + (4 << 20);
    }
    need += (size_t) B * P[4] * (final_conv.w.CinPad * 2 + 4) + (16 << 20);
    if (arena.reserve(need)) return 1;
    FwdT<Snac> F{this, ctx, B};

    // ---- inputs: codes, lengths per level, and the noise every utterance would have seen had they been decoded one after another
    std::vector<uint32_t> hc(n_codes_total);
    std::vector<int> hoff((size_t) B), hl((size_t) 9 * B);
    size_t o = 0;
    for (int b = 0; b < B; b++) {
        const size_t n = (size_t) fine_frames[b] / 4 * 7;
        for (size_t i = 0; i < n; i++) {
            if (codes[b][i] >= (uint32_t) n_codes) { set_error("snac: utterance %d code %u >= codebook size %d", b, codes[b][i], n_codes); return 1; }
            hc[o + i] = codes[b][i];
        }
        hoff[(size_t) b] = (int) o; o += n;
        int L = fine_frames[b];
        for (int l = 0; l <= 4; l++) { hl[(size_t) l * B + b] = L; if (l < 4) { hl[(size_t) (5 + l) * B + b] = L + 1; L *= layers[l].stride; } }
    }
    std::vector<std::vector<float>> hnoise(4);
    for (int l = 0; l < 4; l++) hnoise[(size_t) l].assign((size_t) B * P[l + 1], 0.f);
    {
        NormalGen * g = (NormalGen *) noise_engine;
        for (int b = 0; b < B; b++)            // snac_runner::set_inputs: one random_normal_gen(840 * L) per run, consumed layer by layer
            for (int l = 0; l < 4; l++) {
                float * dst = hnoise[(size_t) l].data() + (size_t) b * P[l + 1];
                const int n = noise_steps[l] * fine_frames[b];
                if (n > P[l + 1]) { set_error("snac: noise_steps do not match the layer strides"); return 1; }
                for (int i = 0; i < n; i++) dst[i] = g->dis(g->e);
            }
    }
    uint32_t * d_codes = F.al<uint32_t>(hc.size());
    int * d_off = F.al<int>(hoff.size());
    int * d_len = F.al<int>(hl.size());
    float * d_noise[4];
    for (int l = 0; l < 4; l++) d_noise[l] = F.al<float>(hnoise[(size_t) l].size());
    if (F.fail) return 1;
    B2_CUDA(cudaEventRecord(ev[0], st));
    B2_CUDA(cudaMemcpyAsync(d_codes, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(d_off, hoff.data(), hoff.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(d_len, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice, st));
    for (int l = 0; l < 4; l++) B2_CUDA(cudaMemcpyAsync(d_noise[l], hnoise[(size_t) l].data(), hnoise[(size_t) l].size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaStreamSynchronize(st));

    // ---- quantizer -> depthwise k7 -> 1x1 up (snac_model.cpp:136-140)
    float * emb = F.al<float>((size_t) B * Fmax * latent);
    float * e2 = F.al<float>((size_t) B * Fmax * latent);
    float * h = F.al<float>((size_t) B * Fmax * up.Cout);
    if (F.fail) return 1;
    {
        dim3 grid(Fmax, B);
        snac_embed_kernel<<<grid, 256, 0, st>>>(d_codes, d_off, tables, n_codes, latent, Fmax, d_len, emb);
        B2_LAUNCH_CHECK(ctx);
        dim3 blk(128, 2), g2(cdiv(Fmax, 2), B);
        snac_dwconv7_kernel<<<g2, blk, 0, st>>>(emb, latent, Fmax, d_len, nullptr, in_w, in_b, 1, e2);
        B2_LAUNCH_CHECK(ctx);
    }
    if (F.snake_conv(up, nullptr, e2, Fmax, d_len, h, nullptr)) return 1;

    const float * x = h;
    for (int l = 0; l < 4; l++) {
        const SnacLayer & L = layers[l];
        const int Lq = P[l] + 1, Pn = P[l + 1], C = L.Cout;
        const int * len_in = d_len + (size_t) l * B, * len_q = d_len + (size_t) (5 + l) * B, * len_out = d_len + (size_t) (l + 1) * B;
        __half * a3 = F.al<__half>((size_t) B * Lq * L.w3.CinPad);
        float * ubuf = F.al<float>(((size_t) B * Pn + 64) * C);
        float * y = F.al<float>((size_t) B * Pn * C);
        float * alt = F.al<float>((size_t) B * Pn * C);
        if (F.fail) return 1;
        if (F.operand(x, L.Cin, L.Cin, P[l], len_in, L.alpha, true, a3, L.w3.CinPad, Lq)) return 1;
        {
            ConvGemmParams p;
            p.A = a3; p.lda = L.w3.CinPad; p.W = L.w3.w; p.bias = L.b_rep; p.outF = ubuf; p.ldo = L.stride * C;
            p.B = B; p.LmaxIn = Lq; p.LmaxOut = Lq; p.lenIn = len_in; p.lenOut = len_q;
            p.N = L.w3.N; p.Npad = L.w3.Npad; p.KW = 2; p.CinPad = L.w3.CinPad; p.CinTrue = L.w3.Cin; p.stride = 1; p.dil = 1; p.pad = 1;
            p.tailClean = true;
            if (conv_gemm(ctx, p)) return 1;
        }
        float * cur = ubuf + (size_t) L.pad * C;
        // noise block: cur += conv1x1(cur) * noise[t]
        if (F.snake_conv(L.noise, nullptr, cur, Pn, len_out, y, nullptr)) return 1;
        {
            dim3 blk(128, 2), grid(cdiv(Pn, 2), B);
            snac_noise_add_kernel<<<grid, blk, 0, st>>>(cur, y, d_noise[l], C, Pn, len_out);
            B2_LAUNCH_CHECK(ctx);
        }
        for (int i = 0; i < 3; i++) {
            const SnacLayer::Unit & U = L.res[i];
            float * nxt = (cur == alt) ? ubuf : alt;
            dim3 blk(128, 2), grid(cdiv(Pn, 2), B);
            snac_dwconv7_kernel<<<grid, blk, 0, st>>>(cur, C, Pn, len_out, U.a1, U.dw_w, U.dw_b, U.dil, y);
            B2_LAUNCH_CHECK(ctx);
            if (F.snake_conv(U.c2, U.a2, y, Pn, len_out, nxt, cur)) return 1;
            cur = nxt;
        }
        x = cur;
    }
    float * pcm_d = F.al<float>((size_t) B * P[4]);
    if (F.fail) return 1;
    if (F.snake_conv(final_conv, final_alpha, x, P[4], d_len + (size_t) 4 * B, pcm_d, nullptr, ACT_TANH)) return 1;
    B2_CUDA(cudaEventRecord(ev[1], st));

    size_t total = 0;
    for (int b = 0; b < B; b++) total += (size_t) fine_frames[b] * up_factor;
    if (pcm_pinned_cap < total) {
        if (pcm_pinned) cudaFreeHost(pcm_pinned);
        pcm_pinned = nullptr; pcm_pinned_cap = 0;
        B2_CUDA(cudaMallocHost(&pcm_pinned, total * 4));
        pcm_pinned_cap = total;
    }
    size_t off = 0;
    for (int b = 0; b < B; b++) {
        const size_t n = (size_t) fine_frames[b] * up_factor;
        if ((int64_t) n > (int64_t) P[4]) { set_error("snac: up_sampling_factor %d does not match the layer strides", up_factor); return 1; }
        B2_CUDA(cudaMemcpyAsync(pcm_pinned + off, pcm_d + (size_t) b * P[4], n * 4, cudaMemcpyDeviceToHost, st));
        if (pcm) pcm[b] = pcm_pinned + off;
        if (n_samples) n_samples[b] = (int64_t) n;
        off += n;
    }
    B2_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&timing_ms, ev[0], ev[1]);
    return 0;
}

}  // namespace b2
