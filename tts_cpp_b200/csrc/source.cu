// source.cu -- harmonic source, STFT / iSTFT, ConvTranspose1d and the small patched ops, for sm_100a.
//
// Replaces (reference file:line):
//   build_sin_gen + uv_noise_compute        src/models/kokoro/model.cpp:173-193, src/util.cpp:140-172
//   random_uniform_gen (host, serial)       src/util.cpp:66-72      -> jump-ahead minstd_rand0 on the device, bit-identical stream
//   ggml_mod / ggml_cumsum / ggml_round / ggml_reciprocal / ggml_upscale_linear   ggml-cpu.c:1797-1798,5526-5720,10912-10964
//   ggml_stft / ggml_istft (+ util wrappers, window-square-sum)                   ggml-cpu.c:8476-8760, src/util.cpp:111-137,203-217
//   ggml_conv_transpose_1d (F32 kernel)                                           ggml-cpu.c:10104-10200
// All of it is HBM-bound byte/float shuffling: one pass over the data, coalesced, no tensor cores.
#include "kernels.cuh"
#include <math.h>
#include <vector>

namespace b2 {
namespace {

// ---------------------------------------------------------------- minstd_rand0 with jump-ahead
constexpr unsigned long long LCG_M = 2147483647ull, LCG_A = 16807ull;
// (a * b) mod (2^31 - 1) for a, b < 2^31: Mersenne folding (2^31 == 1 mod M) instead of a 64-bit remainder
__device__ __forceinline__ unsigned long long mulmod(unsigned long long a, unsigned long long b) {
    const unsigned long long p = a * b;                                   // < 2^62
    unsigned long long r = (p & LCG_M) + (p >> 31);                       // < 2^32
    r = (r & LCG_M) + (r >> 31);                                          // <= M + 1
    return r >= LCG_M ? r - LCG_M : r;
}
__device__ unsigned long long lcg_state(unsigned long long k) {   // A^k mod M == state after k draws from seed 1
    unsigned long long r = 1, a = LCG_A;
    k %= (LCG_M - 1);
    while (k) { if (k & 1) r = mulmod(r, a); a = mulmod(a, a); k >>= 1; }
    return r;
}
__device__ __forceinline__ float lcg_to_uniform(unsigned long long x) {
    // libstdc++ generate_canonical<float,24> over minstd_rand0: (float)(x - 1) / (float)2147483646  [== 2^31 in fp32], clamp < 1
    float u = (float) (x - 1ull) * 4.656612873077392578125e-10f;   // / 2^31: a power of two, so the product is the exactly rounded quotient
    return u >= 1.0f ? 0.99999994f : u;
}

__constant__ float c_cos20[20];
__constant__ float c_sin20[20];
__constant__ float c_hann20[20];
__constant__ float c_hann20sq[20];
bool g_tables_ready = false;

int ensure_tables() {
    if (g_tables_ready) return 0;
    float cs[20], sn[20], hw[20], hw2[20];
    for (int i = 0; i < 20; i++) {
        cs[i] = (float) cos(2.0 * M_PI * i / 20.0);
        sn[i] = (float) sin(2.0 * M_PI * i / 20.0);
        hw[i] = (float) pow(sin(M_PI * (double) i / 20.0), 2.0);   // hann_window (src/util.cpp:132-137)
        hw2[i] = powf(hw[i], 2);                                   // compute_window_squared_sum (src/util.cpp:214)
    }
    B2_CUDA(cudaMemcpyToSymbol(c_cos20, cs, sizeof(cs)));
    B2_CUDA(cudaMemcpyToSymbol(c_sin20, sn, sizeof(sn)));
    B2_CUDA(cudaMemcpyToSymbol(c_hann20, hw, sizeof(hw)));
    B2_CUDA(cudaMemcpyToSymbol(c_hann20sq, hw2, sizeof(hw2)));
    g_tables_ready = true;
    return 0;
}

// ggml_upscale_linear along time for one output index (ggml-cpu.c:10934-10960)
__device__ __forceinline__ float upscale_linear_at(const float * __restrict__ row, int n, int factor, int i0) {
    const int   ne0  = n * factor;
    const float sf0  = (float) ne0 / (float) n;
    const float hsf0 = sf0 / 2.0f;
    const int   sf = (int) sf0, hsf = (int) hsf0;
    if (i0 < hsf) return row[0];
    if (i0 >= ne0 - hsf) return row[n - 1];
    const int   i00 = (int) (((float) i0 - hsf0) / sf0);
    const float base = row[i00], top = row[i00 + 1];
    const float diff_adj = (top - base) / sf0;
    const float adj = fmaf((float) ((i0 - hsf) % sf), diff_adj, diff_adj / 2.0f);
    return base + adj;
}

// ---------------------------------------------------------------- phase accumulator: mod -> serial cumsum -> scale
__global__ void f0_phase_kernel(const float * __restrict__ f0, int L2max, const int * __restrict__ len2, int B, float * phase) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = i / 9, h = i - b * 9;
    if (b >= B) return;
    const float hn = ((float) h + 1.0f) / 24000.0f;                  // harmonic_sampling_norm (model.cpp:375-377)
    const float scal = (float) (600.0 * M_PI);                       // upsample_scale*2.0f*M_PI (model.cpp:388)
    const float * f = f0 + (size_t) b * L2max;
    float * out = phase + ((size_t) b * 9 + h) * L2max;
    float run = 0.f;
    for (int t = 0; t < len2[b]; t++) {
        run = run + fmodf(f[t] * hn, 1.0f);                          // ggml_mod then ggml_cumsum (serial fp32 prefix sum)
        out[t] = run * scal;
    }
}

// ---------------------------------------------------------------- harmonic source: upsample, sin, uv/noise, m_source linear, tanh
// sine of the (large: up to ~4e5 rad) accumulated phase.  sinf() takes its slow Payne-Hanek path above 1e5; reducing in double
// (x - k*2pi with 2pi split hi/lo, error < 1e-10) and calling sinf on the small remainder is as accurate and several times cheaper.
__device__ __forceinline__ float sin_phase(float x) {
    const double xd = (double) x;
    const double k = rint(xd * 0.15915494309189535);
    double r = fma(-k, 6.283185307179586, xd);
    r = fma(-k, 2.4492935982947064e-16, r);
    return sinf((float) r);
}
constexpr int SRC_RUN = 16;  // consecutive samples per thread (amortises the LCG jump-ahead)
__global__ void source_har_kernel(const SourceParams p) {
    const int b = blockIdx.y;
    const int L2 = p.len2[b];
    const int S = L2 * 300;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * SRC_RUN;
    if (j0 >= S) return;
    const int j1 = min(S, j0 + SRC_RUN);
    const unsigned long long skip = p.noise_skip ? p.noise_skip[b] : 0ull;
    double acc[SRC_RUN];
#pragma unroll
    for (int r = 0; r < SRC_RUN; r++) acc[r] = 0.0;
    const float * f0b = p.f0 + (size_t) b * p.L2max;
    for (int h = 0; h < 9; h++) {
        const float * ph = p.phase + ((size_t) b * 9 + h) * p.L2max;
        // noise index of (h, j) inside this utterance's block of 9*S draws is h*S + j  (util.cpp:157-158)
        unsigned long long st = lcg_state(skip + (unsigned long long) h * S + j0 + 1ull);
        const float wh = p.w_src[h];
#pragma unroll
        for (int r = 0; r < SRC_RUN; r++) {
            const int j = j0 + r;
            if (j < j1) {
                const float u = lcg_to_uniform(st);
                st = mulmod(st, LCG_A);
                const float fv = f0b[j / 300];                                     // ggml_upscale_ext nearest (ggml-cpu.c:10870)
                const bool voiced = fv > 10.0f;
                const float uv = voiced ? 0.1f : 0.0f;
                const float nz = voiced ? 0.003f * u : (0.1f / 3.0f) * u;
                const float sv = sin_phase(upscale_linear_at(ph, L2, 300, j)) * uv + nz;
                if (p.sing) p.sing[((size_t) b * p.Smax + j) * 9 + h] = sv;
                // m_source_weight is stored F16 -> activation re-rounded to fp16; ggml_vec_dot_f16 tail accumulates in double
                acc[r] += (double) (__half2float(__float2half_rn(sv)) * wh);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < SRC_RUN; r++) {
        const int j = j0 + r;
        if (j < j1) p.har[(size_t) b * p.Smax + j] = tanhf((float) acc[r] + p.b_src);
    }
}

// ---------------------------------------------------------------- STFT n_fft=20 hop=5, centre reflect, |X| and angle
__global__ void stft20_kernel(const float * __restrict__ har, int Smax, const int * __restrict__ lenS, int Fmax, __half * outH, int ldoh, int Cpad,
                              float * outF, int ldof, int FpitchH) {
    const int b = blockIdx.y;
    const int S = lenS[b];
    const int nfr = S / 5 + 1;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfr) return;
    const float * x = har + (size_t) b * Smax;
    float fr[20];
#pragma unroll
    for (int i = 0; i < 20; i++) {
        int ai = f * 5 - 10 + i;
        if (ai < 0) ai = -ai; else if (ai >= S) ai = S - (ai - S + 1);          // ggml-cpu.c:8601-8611
        fr[i] = x[ai] * c_hann20[i];
    }
    const size_t row = (size_t) b * Fmax + f;
    float mg[11], an[11];
#pragma unroll
    for (int k = 0; k <= 10; k++) {
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int i = 0; i < 20; i++) {
            const int idx = (k * i) % 20;
            re = fmaf(fr[i], c_cos20[idx], re);
            im = fmaf(fr[i], -c_sin20[idx], im);
        }
        if (k == 0 || k == 10) im = 0.0f;   // the reference's radix-2/DFT yields exactly +0.0 here for a real frame
        const float mag = sqrtf(re * re + im * im);
        const float ang = atan2f(im, re);
        mg[k] = mag; an[k] = ang;
        if (outF) { outF[row * ldof + k] = mag; outF[row * ldof + 11 + k] = ang; }
    }
    if (outH) {
        __half * o = outH + ((size_t) b * FpitchH + f) * ldoh;
        if (Cpad % 8 == 0 && Cpad >= 24 && ldoh % 8 == 0 && ((((uintptr_t) outH) & 15) == 0)) {
            // the operand row (22 values + zero pad channels) leaves as 16-byte stores: a row is one or two full 64-byte sectors
            __half hv[24];
#pragma unroll
            for (int k = 0; k < 11; k++) { hv[k] = __float2half_rn(mg[k]); hv[11 + k] = __float2half_rn(an[k]); }
            hv[22] = hv[23] = __float2half_rn(0.f);
#pragma unroll
            for (int q = 0; q < 3; q++) {
                uint4 u;
                u.x = (uint32_t) __half_as_ushort(hv[8 * q]) | ((uint32_t) __half_as_ushort(hv[8 * q + 1]) << 16);
                u.y = (uint32_t) __half_as_ushort(hv[8 * q + 2]) | ((uint32_t) __half_as_ushort(hv[8 * q + 3]) << 16);
                u.z = (uint32_t) __half_as_ushort(hv[8 * q + 4]) | ((uint32_t) __half_as_ushort(hv[8 * q + 5]) << 16);
                u.w = (uint32_t) __half_as_ushort(hv[8 * q + 6]) | ((uint32_t) __half_as_ushort(hv[8 * q + 7]) << 16);
                *reinterpret_cast<uint4 *>(o + 8 * q) = u;
            }
            for (int c = 24; c < Cpad; c += 8) *reinterpret_cast<uint4 *>(o + c) = make_uint4(0u, 0u, 0u, 0u);
        } else {
#pragma unroll
            for (int k = 0; k < 11; k++) { o[k] = __float2half_rn(mg[k]); o[11 + k] = __float2half_rn(an[k]); }
            for (int c = 22; c < Cpad; c++) o[c] = __float2half_rn(0.f);
        }
    }
}

// ---------------------------------------------------------------- iSTFT: (mag, phase) -> (re, im) in place, then overlap-add / window^2 sum
__global__ void istft_polar_kernel(float * specph, int ld, const int * __restrict__ lenF, int Fmax) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = i / 11, k = i - f * 11;
    if (f >= lenF[b]) return;
    float * row = specph + ((size_t) b * Fmax + f) * ld;
    const float m = row[k], ph = row[11 + k];
    row[k] = m * cosf(ph);              // ggml-cpu.c:8745-8751
    row[11 + k] = m * sinf(ph);
}

__global__ void istft_ola_kernel(const float * __restrict__ reim, int ld, const int * __restrict__ lenF, int Fmax, float * pcm, int Smax) {
    const int b = blockIdx.y;
    const int nfr = lenF[b];
    const int S = (nfr - 1) * 5;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= S) return;
    // frames f with 0 <= n + 10 - 5f < 20, accumulated in increasing f (ggml-cpu.c:8665-8688)
    int f_lo = (n - 9 + 4) / 5; if (n - 9 < 0) f_lo = 0;
    int f_hi = (n + 10) / 5;
    float acc = 0.f;
    for (int f = f_lo; f <= f_hi && f < nfr; f++) {
        const int i = n + 10 - 5 * f;
        if (i < 0 || i >= 20) continue;
        const float * row = reim + ((size_t) b * Fmax + f) * ld;
        float v = row[0] + ((i & 1) ? -row[10] : row[10]);
#pragma unroll
        for (int k = 1; k < 10; k++) {
            const int idx = (k * i) % 20;
            v = fmaf(2.0f * row[k], c_cos20[idx], v);
            v = fmaf(-2.0f * row[11 + k], c_sin20[idx], v);
        }
        acc = acc + (v / 20.0f) * c_hann20[i];
    }
    // window-square-sum with the reference's frame range (n_frames + half/hop frames, util.cpp:203-217)
    float wss = 0.f;
    const int n_frames = S / 5;
    int g_lo = (n - 9 + 4) / 5; if (n - 9 < 0) g_lo = 0;
    for (int g = g_lo; g < n_frames + 2; g++) {
        const int ii = n + 10 - 5 * g;
        if (ii < 0) break;
        if (ii < 20) wss = fmaf(c_hann20[ii], c_hann20[ii], wss);   // `tgt += powf(w, 2)` is one FMA in the reference build
    }
    pcm[(size_t) b * Smax + n] = acc / wss;
}

// ---------------------------------------------------------------- ConvTranspose1d, channels-last, fp32 (generator up-convs)
// y[b][o][co] = bias[co] + sum over (t,k) with t*s + k - p == o of sum_ci lrelu(x[b][t][ci]) * w[k][ci][co]
// Block: 64 output positions x Cout; the (<= K/s + 1) contributing input rows are staged in shared memory.
template <int TO>
__global__ void convt_cl_kernel(const float * __restrict__ x, int ldx, int Cin, int LmaxIn, const int * __restrict__ lenIn, const float * __restrict__ w,
                                const float * __restrict__ bias, int K, int Cout, int stride, int pad, float ns, int reflect1, float * y, int ldy,
                                int LmaxOut, const int * __restrict__ lenOut) {
    extern __shared__ float sx[];   // [nrows][Cin]
    const int b = blockIdx.y;
    const int Lo = lenOut[b], Li = lenIn[b];
    const int o0 = blockIdx.x * TO;
    if (o0 >= Lo) return;
    // output index in un-padded coordinates: oo = o - reflect1  (o == 0 with reflect1 mirrors oo = 1)
    const int oo_lo = max(0, o0 - reflect1), oo_hi = min(Lo - 1 - reflect1, o0 + TO - 1 - reflect1);
    const int oo_min = (reflect1 && o0 == 0) ? 0 : oo_lo;
    const int oo_max = (reflect1 && o0 == 0) ? max(oo_hi, 1) : oo_hi;
    int t_lo = (oo_min + pad - (K - 1) + stride - 1) / stride; if (oo_min + pad - (K - 1) < 0) t_lo = 0;
    int t_hi = min(Li - 1, (oo_max + pad) / stride);
    const int nrows = t_hi - t_lo + 1;
    const float * xb = x + (size_t) b * LmaxIn * ldx;
    for (int i = threadIdx.x; i < nrows * Cin; i += blockDim.x) {
        const int r = i / Cin, c = i - r * Cin;
        const float v = xb[(size_t) (t_lo + r) * ldx + c];
        sx[i] = (v > 0.f ? v : 0.f) + ns * (v < 0.f ? v : 0.f);
    }
    __syncthreads();
    // thread -> (co, group of output positions)
    const int nco = Cout;
    for (int item = threadIdx.x; item < nco * TO; item += blockDim.x) {
        const int co = item % nco, oi = item / nco;
        const int o = o0 + oi;
        if (o >= Lo) continue;
        int oo = o - reflect1; if (oo < 0) oo = 1;
        float acc = 0.f;
        // contributions in increasing t (the reference accumulates dst += v in t order, one dot product over Cin each)
        int ta = (oo + pad - (K - 1) + stride - 1) / stride; if (oo + pad - (K - 1) < 0) ta = 0;
        const int tb = min(Li - 1, (oo + pad) / stride);
        for (int t = ta; t <= tb; t++) {
            const int k = oo + pad - t * stride;
            const float * xr = sx + (size_t) (t - t_lo) * Cin;
            const float * wr = w + ((size_t) k * Cin) * Cout + co;
            float d = 0.f;
            for (int ci = 0; ci < Cin; ci++) d = fmaf(xr[ci], wr[(size_t) ci * Cout], d);
            acc = acc + d;
        }
        y[((size_t) b * LmaxOut + o) * ldy + co] = acc + bias[co];
    }
}

// ---------------------------------------------------------------- op-level kernels (ggml layout: x[c][l])
__global__ void op_cumsum_kernel(const float * x, int L, int rows, float * y) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float run = 0.f;
    for (int i = 0; i < L; i++) { run += x[(size_t) r * L + i]; y[(size_t) r * L + i] = run; }
}
__global__ void op_unary_kernel(int which, const float * x, int64_t n, float arg, float * y) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    y[i] = which == 0 ? fmodf(v, arg) : which == 1 ? (float) ((int) (v + 0.5f)) : 1.0f / v;
}
__global__ void op_upscale_linear_kernel(const float * x, int L, int rows, int factor, float * y) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ne0 = (int64_t) L * factor;
    if (i >= ne0 * rows) return;
    const int r = (int) (i / ne0), i0 = (int) (i - (int64_t) r * ne0);
    y[i] = upscale_linear_at(x + (size_t) r * L, L, factor, i0);
}
__global__ void op_snake_kernel(const float * alpha, int C, const float * x, int L, float * y) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t) C * L) return;
    const float a = alpha[i / L], v = x[i];
    const float s = sinf(v * a);
    y[i] = v + (s * s) * (1.0f / a);
}
__global__ void op_uniform_kernel(unsigned long long skip, int64_t count, float * y) {
    const int64_t j0 = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (j0 >= count) return;
    unsigned long long st = lcg_state(skip + (unsigned long long) j0 + 1ull);
    for (int r = 0; r < 16 && j0 + r < count; r++) { y[j0 + r] = lcg_to_uniform(st); st = mulmod(st, LCG_A); }
}
// generic ggml-layout ConvTranspose1d (groups, padding, output padding), one thread per output element
__global__ void op_convt_kernel(const float * w, int K, int coutg, int cin, const float * x, int L, int s, int p, int g, float * y, int Lout) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int cout = coutg * g;
    if (i >= (int64_t) cout * Lout) return;
    const int co = (int) (i / Lout), o = (int) (i - (int64_t) co * Lout);
    const int cing = cin / g, grp = co / coutg, col = co - grp * coutg;
    float acc = 0.f;
    for (int t = 0; t < L; t++) {
        const int k = o + p - t * s;
        if (k < 0 || k >= K) continue;
        float d = 0.f;
        for (int ci = grp * cing; ci < (grp + 1) * cing; ci++) d = fmaf(x[(size_t) ci * L + t], w[((size_t) ci * coutg + col) * K + k], d);
        acc += d;
    }
    y[i] = acc;
}

}  // namespace

int source_har(Ctx * ctx, const SourceParams & p) {
    if (ensure_tables()) return 1;
    f0_phase_kernel<<<cdiv(p.B * 9, 64), 64, 0, ctx->stream>>>(p.f0, p.L2max, p.len2, p.B, p.phase);
    B2_LAUNCH_CHECK(ctx);
    dim3 grid(cdiv(cdiv(p.Smax, SRC_RUN), 128), p.B);
    source_har_kernel<<<grid, 128, 0, ctx->stream>>>(p);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int stft20(Ctx * ctx, const float * har, int Smax, int B, const int * lenS, int Fmax, __half * outH, int ldoh, int Cpad, float * outF, int ldof, int FpitchH) {
    if (ensure_tables()) return 1;
    dim3 grid(cdiv(Fmax, 128), B);
    stft20_kernel<<<grid, 128, 0, ctx->stream>>>(har, Smax, lenS, Fmax, outH, ldoh, Cpad, outF, ldof, FpitchH > 0 ? FpitchH : Fmax);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int istft20(Ctx * ctx, float * specph, int ld, int B, const int * lenF, int Fmax, float * pcm, int Smax) {
    if (ensure_tables()) return 1;
    dim3 g1(cdiv((int64_t) Fmax * 11, 256), B);
    istft_polar_kernel<<<g1, 256, 0, ctx->stream>>>(specph, ld, lenF, Fmax);
    B2_LAUNCH_CHECK(ctx);
    dim3 g2(cdiv(Smax, 256), B);
    istft_ola_kernel<<<g2, 256, 0, ctx->stream>>>(specph, ld, lenF, Fmax, pcm, Smax);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int convt_cl(Ctx * ctx, const float * x, int ldx, int Cin, int B, int LmaxIn, const int * lenIn, const float * w, const float * bias, int K, int Cout,
             int stride, int pad, float ns, int reflect1, float * y, int ldy, int LmaxOut, const int * lenOut) {
    constexpr int TO = 32;
    const int nrows_max = (TO + 1 + K) / stride + 3;
    const size_t smem = (size_t) nrows_max * Cin * sizeof(float);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        B2_CUDA(cudaFuncSetAttribute(convt_cl_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        smem_set = smem;
    }
    dim3 grid(cdiv(LmaxOut, TO), B);
    ctx->prof_begin(PROF_CONVT, 2.0 * B * LmaxIn * (double) K * Cin * Cout, 0.0);
    convt_cl_kernel<TO><<<grid, 256, smem, ctx->stream>>>(x, ldx, Cin, LmaxIn, lenIn, w, bias, K, Cout, stride, pad, ns, reflect1, y, ldy, LmaxOut, lenOut);
    ctx->prof_end();
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int op_cumsum(Ctx * ctx, const float * x, int L, int rows, float * y) {
    op_cumsum_kernel<<<cdiv(rows, 64), 64, 0, ctx->stream>>>(x, L, rows, y);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}
int op_unary(Ctx * ctx, int which, const float * x, int64_t n, float arg, float * y) {
    op_unary_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(which, x, n, arg, y);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}
int op_upscale_linear(Ctx * ctx, const float * x, int L, int rows, int factor, float * y) {
    op_upscale_linear_kernel<<<cdiv((int64_t) L * factor * rows, 256), 256, 0, ctx->stream>>>(x, L, rows, factor, y);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}
int op_snake(Ctx * ctx, const float * alpha, int C, const float * x, int L, float * y) {
    op_snake_kernel<<<cdiv((int64_t) C * L, 256), 256, 0, ctx->stream>>>(alpha, C, x, L, y);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}
int op_uniform(Ctx * ctx, unsigned long long skip, int64_t count, float * y) {
    op_uniform_kernel<<<cdiv(cdiv(count, 16), 128), 128, 0, ctx->stream>>>(skip, count, y);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}
int op_conv_transpose_1d(Ctx * ctx, const float * w, int K, int coutg, int cin, const float * x, int L, int s, int p, int op, int g, float * y,
                         int Lout) {
    (void) op;
    op_convt_kernel<<<cdiv((int64_t) coutg * g * Lout, 128), 128, 0, ctx->stream>>>(w, K, coutg, cin, x, L, s, p, g, y, Lout);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b2
