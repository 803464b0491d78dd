// elementwise.cu -- normalisation, activation, gather and small-matrix kernels of the Kokoro forward (sm_100a).
// Everything here is HBM/L2-bound: channels-last rows are read/written with consecutive threads on consecutive
// channels (coalesced), per-channel parameters live in registers, and each kernel fuses the chain of GGML nodes the
// reference spends separate passes on (norm -> mul -> add -> add -> leaky_relu/snake -> cont(transpose) -> fp16 im2col).
#include "kernels.cuh"
#include <type_traits>
#include <math.h>

namespace b2 {
namespace {

constexpr int KMAX = 5;  // channels per thread (C <= 1280)

__device__ __forceinline__ float lrelu(float v, float ns) { return (v > 0.f ? v : 0.f) + ns * (v < 0.f ? v : 0.f); }  // ggml_vec_leaky_relu_f32

// ---------------------------------------------------------------- InstanceNorm statistics
__global__ void inorm_stats_kernel(const float * __restrict__ x, int ldx, int C, int Lmax, const int * __restrict__ len, double * sums,
                                   int rows_per_block) {
    const int b = blockIdx.y;
    const int L = len[b];
    const int t0 = blockIdx.x * rows_per_block;
    if (t0 >= L) return;
    const int t1 = min(L, t0 + rows_per_block);
    const int tx = threadIdx.x, ty = threadIdx.y, nx = blockDim.x, ny = blockDim.y;
    double s[KMAX], q[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) s[k] = q[k] = 0.0;
    const float * xb = x + (size_t) b * Lmax * ldx;
    for (int t = t0 + ty; t < t1; t += ny) {
        const float * row = xb + (size_t) t * ldx;
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            const int c = tx + k * nx;
            if (c < C) { const double v = (double) row[c]; s[k] += v; q[k] += v * v; }
        }
    }
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
        const int c = tx + k * nx;
        if (c < C) {
            atomicAdd(&sums[((size_t) b * C + c) * 2 + 0], s[k]);
            atomicAdd(&sums[((size_t) b * C + c) * 2 + 1], q[k]);
        }
    }
}

// ---------------------------------------------------------------- AdaIN apply + activation
constexpr int TAIL_ROWS = 32;   // rows past an utterance's end that a conv may read as padding: (K-1)*dil - pad <= 25 for every Kokoro layer

__global__ void adain_apply_kernel(const AdainParams p, int rows_per_block) {
    const int b = blockIdx.y;
    const int L = p.len[b];
    const int t0 = blockIdx.x * rows_per_block;
    if (t0 >= L) return;
    const int t1 = min(L, t0 + rows_per_block);
    const int tx = threadIdx.x, ty = threadIdx.y, nx = blockDim.x, ny = blockDim.y;
    float mean[KMAX], rstd[KMAX], gam[KMAX], bet[KMAX], al[KMAX], ial[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
        const int c = tx + k * nx;
        mean[k] = 0.f; rstd[k] = 0.f; gam[k] = 0.f; bet[k] = 0.f; al[k] = 1.f; ial[k] = 1.f;
        if (c < p.C) {
            const double s = p.sums[((size_t) b * p.C + c) * 2], q = p.sums[((size_t) b * p.C + c) * 2 + 1];
            const double m = s / (double) L;
            double var = q / (double) L - m * m;
            if (var < 0.0) var = 0.0;
            mean[k] = (float) m;
            rstd[k] = 1.0f / sqrtf((float) var + 1e-5f);
            gam[k]  = p.gb[(size_t) b * p.ldgb + p.goff + c];
            bet[k]  = p.gb[(size_t) b * p.ldgb + p.boff + c];
            if (p.act == NACT_SNAKE) { al[k] = p.alpha[c]; ial[k] = 1.0f / al[k]; }
        }
    }
    const float * xb = p.x + (size_t) b * p.Lmax * p.ldx;
    for (int t = t0 + ty; t < t1; t += ny) {
        const float * row = xb + (size_t) t * p.ldx;
        const size_t orow = (size_t) b * p.Lmax + t;
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            const int c = tx + k * nx;
            if (c < p.C) {
                const float n = (row[c] - mean[k]) * rstd[k];
                float v = (n + n * gam[k]) + bet[k];                       // model.cpp:99-100 / :148
                if (p.act == NACT_LRELU02) v = lrelu(v, 0.2f);
                else if (p.act == NACT_SNAKE) { const float sn = sinf(v * al[k]); v = v + (sn * sn) * ial[k]; }   // util.cpp:99-101
                if (p.outH) p.outH[orow * p.ldoh + c] = __float2half_rn(v);
                if (p.outF) p.outF[orow * p.ldof + c] = v;
            } else if (c < p.Cpad && p.outH) {
                p.outH[orow * p.ldoh + c] = __float2half_rn(0.f);
            }
        }
    }
    if (p.outH && t1 == L)   // clear the rows past the utterance's end (see adain_apply4_kernel)
        for (int t = L + ty; t < min(L + TAIL_ROWS, p.Lmax); t += ny)
            for (int c = tx; c < p.Cpad; c += nx) p.outH[((size_t) b * p.Lmax + t) * p.ldoh + c] = __float2half_rn(0.f);
}

// ---------------------------------------------------------------- vectorised variants (C % 4 == 0): one thread = 4 channels,
// 16-byte loads, 8-byte fp16 stores, 4 rows in flight per thread -> these passes run at HBM speed instead of latency-bound.
constexpr int V4_ROWS = 64;   // rows per block

// sin for the snake activation: 2-term Cody-Waite reduction to [-pi, pi] + the SFU sine (abs error ~5e-7 here, i.e. ~1e-3 of
// an fp16 ulp of the value this feeds -- the result is re-rounded to fp16 for the next conv).  libm-grade sinf made this
// HBM-bound pass issue-bound.
__device__ __forceinline__ float sin_snake(float x) {
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, x);
    r = fmaf(-k, -1.7484556000744883e-7f, r);
    return __sinf(r);
}

__global__ void __launch_bounds__(256) inorm_stats4_kernel(const float * __restrict__ x, int ldx, int C, int Lmax, const int * __restrict__ len, double * sums) {
    const int b = blockIdx.y;
    const int L = len[b];
    const int t0 = blockIdx.x * V4_ROWS;
    if (t0 >= L) return;
    const int t1 = min(L, t0 + V4_ROWS);
    const int c4 = threadIdx.x, ny = blockDim.y;
    const int cg = (blockIdx.z * blockDim.x + c4) * 4;    // first of this thread's 4 channels (channels beyond 1024 use a second z slice);
    if (cg >= C) return;                                  // C need not be a multiple of 4: rows are padded to ldx % 4 == 0
    const float * xb = x + (size_t) b * Lmax * ldx + cg;
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int t = t0 + threadIdx.y; t < t1; t += 4 * ny) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int tt = t + u * ny;
            v[u] = tt < t1 ? *reinterpret_cast<const float4 *>(xb + (size_t) tt * ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // fp32 partial over 4 rows, then into the double accumulators (ggml accumulates in double: ggml-cpu.c:7138-7152)
        float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            ps[0] += v[u].x; ps[1] += v[u].y; ps[2] += v[u].z; ps[3] += v[u].w;
            pq[0] = fmaf(v[u].x, v[u].x, pq[0]); pq[1] = fmaf(v[u].y, v[u].y, pq[1]); pq[2] = fmaf(v[u].z, v[u].z, pq[2]); pq[3] = fmaf(v[u].w, v[u].w, pq[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) { s[k] += (double) ps[k]; q[k] += (double) pq[k]; }
    }
    // combine the ny row-lanes of the block through shared memory, then one atomic per channel per block
    __shared__ double sh[256 * 8];
    double * mine = sh + (threadIdx.y * blockDim.x + c4) * 8;
#pragma unroll
    for (int k = 0; k < 4; k++) { mine[k] = s[k]; mine[4 + k] = q[k]; }
    __syncthreads();
    if (threadIdx.y == 0) {
        for (int y = 1; y < ny; y++) {
            const double * o = sh + (y * blockDim.x + c4) * 8;
#pragma unroll
            for (int k = 0; k < 4; k++) { s[k] += o[k]; q[k] += o[4 + k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (cg + k >= C) break;
            atomicAdd(&sums[((size_t) b * C + cg + k) * 2 + 0], s[k]);
            atomicAdd(&sums[((size_t) b * C + cg + k) * 2 + 1], q[k]);
        }
    }
}

// RAGGED: C is not a multiple of 4 (rows padded to ldx % 4 == 0): per-channel liveness checks, kept out of the common instantiation
// because a per-element test in the inner loop cost 18 % on the 585 MB generator tensors
// ACT (NACT_*) and the output flavour are template constants so that the 4 x 4 unrolled element chains of an iteration form one
// basic block and interleave (runtime-uniform `if`s inside the loop ended a block per element and serialised them)
template <bool RAGGED, int ACT, bool OUTH, bool OUTF>
__global__ void __launch_bounds__(256) adain_apply4_kernel(const AdainParams p, const int rows_per_block) {
    const int b = blockIdx.y;
    const int L = p.len[b];
    const int t0 = blockIdx.x * rows_per_block;
    if (t0 >= L) return;
    const int t1 = min(L, t0 + rows_per_block);
    const int c0 = (blockIdx.z * blockDim.x + threadIdx.x) * 4, ny = blockDim.y;
    if (c0 >= (p.outH ? p.Cpad : p.C)) return;
    const bool live = c0 < p.C;                           // some of the 4 channels are real (C need not be a multiple of 4)
    float mean[4], rstd[4], gam[4], bet[4], al[4], ial[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        mean[k] = 0.f; rstd[k] = 0.f; gam[k] = 0.f; bet[k] = 0.f; al[k] = 1.f; ial[k] = 1.f;
        if (RAGGED ? (c0 + k < p.C) : live) {
            const int c = c0 + k;
            const double s = p.sums[((size_t) b * p.C + c) * 2], q = p.sums[((size_t) b * p.C + c) * 2 + 1];
            const double m = s / (double) L;
            double var = q / (double) L - m * m;
            if (var < 0.0) var = 0.0;
            mean[k] = (float) m;
            rstd[k] = 1.0f / sqrtf((float) var + 1e-5f);
            gam[k]  = p.gb[(size_t) b * p.ldgb + p.goff + c];
            bet[k]  = p.gb[(size_t) b * p.ldgb + p.boff + c];
            if (p.act == NACT_SNAKE) { al[k] = p.alpha[c]; ial[k] = 1.0f / al[k]; }
        }
    }
    const float * xb = p.x + (size_t) b * p.Lmax * p.ldx + c0;
    for (int t = t0 + threadIdx.y; t < t1; t += 4 * ny) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int tt = t + u * ny;
            v[u] = (live && tt < t1) ? *reinterpret_cast<const float4 *>(xb + (size_t) tt * p.ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int tt = t + u * ny;
            const bool ok = tt < t1;
            float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float n = (f[k] - mean[k]) * rstd[k];
                float w = (n + n * gam[k]) + bet[k];
                if constexpr (ACT == NACT_LRELU02) w = lrelu(w, 0.2f);
                else if constexpr (ACT == NACT_SNAKE) { const float sn = sin_snake(w * al[k]); w = w + (sn * sn) * ial[k]; }
                f[k] = (RAGGED ? (c0 + k < p.C) : live) ? w : 0.f;     // pad channels of the fp16 operand are zeros
            }
            const size_t orow = (size_t) b * p.Lmax + tt;
            if constexpr (OUTH) {
                __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
                uint2 pk; pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
                if (ok) *reinterpret_cast<uint2 *>(p.outH + orow * p.ldoh + c0) = pk;
            }
            if constexpr (OUTF) { if (ok && live) *reinterpret_cast<float4 *>(p.outF + orow * p.ldof + c0) = make_float4(f[0], f[1], f[2], f[3]); }
        }
    }
    if constexpr (OUTH) {
        // the block that holds the utterance's last row also clears the TAIL_ROWS rows past it: the conv that consumes this operand
        // must read zeros there (ragged batches), which saves it a separate clearing launch (ConvGemmParams::tailClean)
        if (t1 == L)
            for (int tt = L + threadIdx.y; tt < min(L + TAIL_ROWS, p.Lmax); tt += ny)
                *reinterpret_cast<uint2 *>(p.outH + ((size_t) b * p.Lmax + tt) * p.ldoh + c0) = make_uint2(0u, 0u);
    }
}

// ---------------------------------------------------------------- depthwise ConvTranspose1d k3 s2 p1 op1 ("pool")
__global__ void pool_convt_kernel(const float * __restrict__ x, int ldx, int C, int Lmax, const int * __restrict__ len,
                                  const float * __restrict__ w3, const float * __restrict__ bias, __half * outH, int ldoh, int Cpad) {
    const int b = blockIdx.y;
    const int L = len[b];
    const int o = blockIdx.x * blockDim.y + threadIdx.y;   // output position in [0, 2L)
    if (o >= 2 * L) return;
    const float * xb = x + (size_t) b * Lmax * ldx;
    const size_t orow = (size_t) b * (2 * Lmax) + o;
    const int q = o >> 1;
    for (int c = threadIdx.x; c < Cpad; c += blockDim.x) {
        float v = 0.f;
        if (c < C) {
            // dst is zero-initialised and accumulated in (t, k) order (ggml-cpu.c:10180-10195); kernel layout w3[c*3 + k]
            if ((o & 1) == 0) {
                v = 0.f + xb[(size_t) q * ldx + c] * w3[c * 3 + 1];
            } else {
                v = 0.f + xb[(size_t) q * ldx + c] * w3[c * 3 + 2];
                if (q + 1 < L) v = v + xb[(size_t) (q + 1) * ldx + c] * w3[c * 3 + 0];
            }
            v = v + bias[c];
        }
        outH[orow * ldoh + c] = __float2half_rn(v);
    }
}

// ---------------------------------------------------------------- row LayerNorm (one warp per row)
__global__ void row_norm_kernel(const RowNormParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int b = warp / p.Lmax, t = warp - b * p.Lmax;
    if (b >= p.B || t >= p.len[b]) return;
    const float * grow = p.x + ((size_t) b * p.Lmax + t) * p.ldx;
    float row_r[32];                                       // the row, read once (C <= 1024: 32 values per lane)
#pragma unroll
    for (int k = 0; k < 32; k++) { const int c = lane + 32 * k; row_r[k] = c < p.C ? grow[c] : 0.f; }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 32; k++) { if (lane + 32 * k < p.C) s += (double) row_r[k]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = (float) (s / (double) p.C);
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < 32; k++) { if (lane + 32 * k < p.C) { const float v = row_r[k] - mean; q += (double) (v * v); } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float var = (float) (q / (double) p.C);
    const float scale = 1.0f / sqrtf(var + p.eps);
    const size_t orow = (size_t) b * p.Lmax + t;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const int c = lane + 32 * k;
        if (c >= p.C) break;
        const float n = (row_r[k] - mean) * scale;
        float v;
        if (p.mode == LN_AFFINE) v = n * p.w[c] + p.bias[c];
        else v = (n + n * p.gb[(size_t) b * p.ldgb + p.goff + c]) + p.gb[(size_t) b * p.ldgb + p.boff + c];
        if (p.lrelu02) v = lrelu(v, 0.2f);
        if (p.outF) p.outF[orow * p.ldof + p.coff + c] = v;
        if (p.outH) p.outH[orow * p.ldoh + p.coffh + c] = __float2half_rn(v);
    }
}

// ---------------------------------------------------------------- helpers
// 16-byte flavour of cast_rows (C % 4 == 0, aligned rows): thread = 4 channels x 4 rows in flight
__global__ void __launch_bounds__(256) cast_rows4_kernel(const float * __restrict__ x, int ldx, int C, int LmaxIn, const int * __restrict__ lenOut, int LmaxOut,
                                                         int up2, float ns, __half * outH, int ldoh, int Cpad, int rows_per_block) {
    const int b = blockIdx.y;
    const int L = lenOut[b];
    const int t0 = blockIdx.x * rows_per_block;
    if (t0 >= L) return;
    const int t1 = min(L, t0 + rows_per_block);
    const int c0 = threadIdx.x * 4, ny = blockDim.y;
    if (c0 >= Cpad) return;
    const bool live = c0 < C;
    const float * xb = x + (size_t) b * LmaxIn * ldx + c0;
    for (int t = t0 + threadIdx.y; t < t1; t += 4 * ny) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int tt = t + u * ny;
            v[u] = (live && tt < t1) ? *reinterpret_cast<const float4 *>(xb + (size_t) (up2 ? (tt >> 1) : tt) * ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int tt = t + u * ny;
            const __half2 h0 = __floats2half2_rn(lrelu(v[u].x, ns), lrelu(v[u].y, ns)), h1 = __floats2half2_rn(lrelu(v[u].z, ns), lrelu(v[u].w, ns));
            uint2 pk; pk.x = *reinterpret_cast<const uint32_t *>(&h0); pk.y = *reinterpret_cast<const uint32_t *>(&h1);
            if (tt < t1) *reinterpret_cast<uint2 *>(outH + ((size_t) b * LmaxOut + tt) * ldoh + c0) = pk;
        }
    }
}

__global__ void cast_rows_kernel(const float * __restrict__ x, int ldx, int C, int LmaxIn, const int * __restrict__ lenOut, int LmaxOut, int up2,
                                 float ns, __half * outH, int ldoh, int Cpad) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= lenOut[b]) return;
    const int ts = up2 ? (t >> 1) : t;
    const float * row = x + ((size_t) b * LmaxIn + ts) * ldx;
    __half * orow = outH + ((size_t) b * LmaxOut + t) * ldoh;
    for (int c = threadIdx.x; c < Cpad; c += blockDim.x) orow[c] = __float2half_rn(c < C ? lrelu(row[c], ns) : 0.f);
}

__global__ void split3_rows_kernel(const float * __restrict__ x, int ldx, int C, int LmaxIn, const int * __restrict__ len, float ns, __half * outH, int Lq) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.y + threadIdx.y;
    if (q >= Lq) return;
    __half * orow = outH + ((size_t) b * Lq + q) * (3 * C);
    const bool live = q < len[b];
    const float * row = x + ((size_t) b * LmaxIn + (live ? q : 0)) * ldx;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float v = live ? lrelu(row[c], ns) : 0.f;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        orow[c] = hi; orow[C + c] = lo; orow[2 * C + c] = hi;
    }
}

__global__ void copy_cols_kernel(const float * __restrict__ src, int lds, int scoff, float * dst, int ldd, int dcoff, int C, int Lmax,
                                 const int * __restrict__ len) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= len[b]) return;
    const size_t r = (size_t) b * Lmax + t;
    for (int c = threadIdx.x; c < C; c += blockDim.x) dst[r * ldd + dcoff + c] = src[r * lds + scoff + c];
}

__global__ void bcast_cols_kernel(const float * __restrict__ v, int ldv, int C, int Lmax, const int * __restrict__ len, float * dstF, int ldf,
                                  int cofff, __half * dstH, int ldh, int coffh) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= len[b]) return;
    const size_t r = (size_t) b * Lmax + t;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float val = v[(size_t) b * ldv + c];
        if (dstF) dstF[r * ldf + cofff + c] = val;
        if (dstH) dstH[r * ldh + coffh + c] = __float2half_rn(val);
    }
}

__global__ void gather_rows_kernel(const float * __restrict__ src, int lds, int LmaxSrc, const int * __restrict__ idx, int C, int Lmax,
                                   const int * __restrict__ len, float * dstF, int ldf, __half * dstH, int ldh, int Cpad) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= len[b]) return;
    const int j = idx[(size_t) b * Lmax + t];
    const float * row = src + ((size_t) b * LmaxSrc + j) * lds;
    const size_t r = (size_t) b * Lmax + t;
    for (int c = threadIdx.x; c < Cpad; c += blockDim.x) {
        const float val = c < C ? row[c] : 0.f;
        if (dstF && c < C) dstF[r * ldf + c] = val;
        if (dstH) dstH[r * ldh + c] = __float2half_rn(val);
    }
}

__global__ void linear_f32_kernel(const float * __restrict__ x, int ldx, const float * __restrict__ W, const float * __restrict__ bias, int rows,
                                  int K, int N, float * y, int ldy) {
    // one warp per (row, 32 outputs): lanes stride over K for each output -> coalesced W reads, shuffle reduce
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int ngroups = (N + 31) / 32;
    const int r = warp / ngroups, g = warp - r * ngroups;
    if (r >= rows) return;
    const float * xr = x + (size_t) r * ldx;
    for (int j = 0; j < 32; j++) {
        const int n = g * 32 + j;
        if (n >= N) break;
        const float * wr = W + (size_t) n * K;
        float acc = 0.f;
        for (int k = lane; k < K; k += 32) acc = fmaf(xr[k], wr[k], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[(size_t) r * ldy + n] = acc + (bias ? bias[n] : 0.f);
    }
}

// K == 128 flavour (the AdaIN gamma/beta projections: 32 style rows x ~30 K outputs, and ALBERT's embedding projection): a block
// stages 32 rows of x in shared memory, a warp owns one output n at a time: its lanes hold W[n][4l..4l+3] (one coalesced 512 B read,
// W is read ONCE, not once per row), form the 32 row partials and a shuffle reduce-scatter leaves lane r with y[r][n].
__global__ void __launch_bounds__(256) linear_f32_k128_kernel(const float * __restrict__ x, int ldx, const float * __restrict__ W,
                                                              const float * __restrict__ bias, int rows, int N, float * y, int ldy, int n_per_block) {
    __shared__ float4 xs[32][32];
    const int r0 = blockIdx.y * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
        const int r = i >> 5, k4 = i & 31;
        xs[r][k4] = (r0 + r < rows) ? *reinterpret_cast<const float4 *>(x + (size_t) (r0 + r) * ldx + 4 * k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int n_lo = blockIdx.x * n_per_block, n_hi = min(N, n_lo + n_per_block);
    for (int n = n_lo + warp; n < n_hi; n += 8) {
        const float4 w = *reinterpret_cast<const float4 *>(W + (size_t) n * 128 + 4 * lane);
        float p[32];
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const float4 xv = xs[r][lane];
            p[r] = fmaf(xv.w, w.w, fmaf(xv.z, w.z, fmaf(xv.y, w.y, xv.x * w.x)));
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; i++) {
                const float send = up ? p[i] : p[i + off], keep = up ? p[i + off] : p[i];
                p[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        if (r0 + lane < rows) y[(size_t) (r0 + lane) * ldy + n] = p[0] + (bias ? bias[n] : 0.f);
    }
}

__global__ void albert_embed_kernel(const int * __restrict__ tokens, const int * __restrict__ tok_off, const float * __restrict__ tok_embd,
                                    const float * __restrict__ pos_embd, const float * __restrict__ type_embd, const float * __restrict__ nw,
                                    const float * __restrict__ nb, int B, int Lmax, const int * __restrict__ len, float * out, int ldo) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int b = warp / Lmax, t = warp - b * Lmax;
    if (b >= B || t >= len[b]) return;
    const int tok = tokens[tok_off[b] + t];
    float v[4];
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = lane + k * 32;
        v[k] = (tok_embd[(size_t) tok * 128 + c] + pos_embd[(size_t) t * 128 + c]) + type_embd[c];
        s += (double) v[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = (float) (s / 128.0);
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = v[k] - mean; q += (double) (v[k] * v[k]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float scale = 1.0f / sqrtf((float) (q / 128.0) + 1e-12f);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = lane + k * 32;
        out[((size_t) b * Lmax + t) * ldo + c] = (v[k] * scale) * nw[c] + nb[c];
    }
}

__global__ void embed_rows_h_kernel(const int * __restrict__ tokens, const int * __restrict__ tok_off, const __half * __restrict__ table, int C,
                                    int Lmax, const int * __restrict__ len, __half * outH, int ldoh) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.y + threadIdx.y;
    if (t >= len[b]) return;
    const int tok = tokens[tok_off[b] + t];
    for (int c = threadIdx.x; c < C; c += blockDim.x) outH[((size_t) b * Lmax + t) * ldoh + c] = table[(size_t) tok * C + c];
}

// ---------------------------------------------------------------- ALBERT attention (fp32)
// one block per (query tile of 8, head, utterance); scores kept in shared memory
__global__ void albert_attention_kernel(const float * __restrict__ qkv, int Lmax, const int * __restrict__ len, int heads, int hd, float scale,
                                        __half * outH, int ldoh) {
    extern __shared__ float sm[];
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 8;
    const int n = len[b];
    if (q0 >= n) return;
    const int D = heads * hd;
    float * sq = sm;                 // [8][hd]
    float * sc = sm + 8 * hd;        // [8][n]
    const int nq = min(8, n - q0);
    const float * base = qkv + (size_t) b * Lmax * 3 * D;
    for (int i = threadIdx.x; i < nq * hd; i += blockDim.x) sq[i] = base[(size_t) (q0 + i / hd) * 3 * D + h * hd + (i % hd)];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    // scores: warp per key
    for (int j = warp; j < n; j += nw) {
        const float * kr = base + (size_t) j * 3 * D + D + h * hd;
        float kv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) kv[u] = (lane + u * 32 < hd) ? kr[lane + u * 32] : 0.f;
        for (int qi = 0; qi < nq; qi++) {
            float a = 0.f;
#pragma unroll
            for (int u = 0; u < 4; u++) if (lane + u * 32 < hd) a = fmaf(sq[qi * hd + lane + u * 32], kv[u], a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) sc[qi * n + j] = a * scale;
        }
    }
    __syncthreads();
    // softmax: warp per query (ggml_compute_forward_soft_max_f32: max, exp, double sum, scale)
    for (int qi = warp; qi < nq; qi += nw) {
        float mx = -INFINITY;
        for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sc[qi * n + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double sum = 0.0;
        for (int j = lane; j < n; j += 32) { const float e = expf(sc[qi * n + j] - mx); sc[qi * n + j] = e; sum += (double) e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = (float) (1.0 / sum);
        for (int j = lane; j < n; j += 32) sc[qi * n + j] *= inv;
    }
    __syncthreads();
    // PV: thread per (query, dim)
    for (int i = threadIdx.x; i < nq * hd; i += blockDim.x) {
        const int qi = i / hd, d = i % hd;
        const float * vr = base + 2 * D + h * hd + d;
        float a = 0.f;
        for (int j = 0; j < n; j++) a = fmaf(sc[qi * n + j], vr[(size_t) j * 3 * D], a);
        outH[((size_t) b * Lmax + q0 + qi) * ldoh + h * hd + d] = __float2half_rn(a);
    }
}

// head dim 64 (Kokoro's ALBERT): one block per (tile of 16 queries, head, utterance), 128 threads.  Keys / values stream through
// shared memory in tiles of 64; a thread owns one key (scores) or one output dim (PV) for 8 queries, so every shared-memory read is
// either a broadcast or conflict-free and there are no shuffle reductions in the inner loops.
constexpr int ATT_Q = 16, ATT_K = 128, ATT_HD = 64;   // key / value tile of 128 rows: a 66-token prompt is one tile (two with 64 wasted a pass)
__global__ void __launch_bounds__(128) albert_attention64_kernel(const float * __restrict__ qkv, int Lmax, const int * __restrict__ len, int heads,
                                                                 float scale, __half * outH, int ldoh, int n_pad) {
    extern __shared__ float sm[];
    float * sQ = sm;                               // [16][64]
    float * sK = sQ + ATT_Q * ATT_HD;              // [64][65]   (keys, later values [64][64])
    float * sS = sK + ATT_K * (ATT_HD + 1);        // [16][n_pad]
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_Q;
    const int n = len[b];
    if (q0 >= n) return;
    const int D = heads * ATT_HD, tid = threadIdx.x;
    const int nq = min(ATT_Q, n - q0);
    const float * base = qkv + (size_t) b * Lmax * 3 * D + h * ATT_HD;
    for (int i = tid; i < ATT_Q * ATT_HD; i += 128) {
        const int qi = i >> 6, d = i & 63;
        sQ[i] = qi < nq ? base[(size_t) (q0 + qi) * 3 * D + d] : 0.f;
    }
    const int jj = tid & 63, qg = (tid >> 6) * 8;
    // ---- scores
    for (int k0 = 0; k0 < n; k0 += ATT_K) {
        __syncthreads();
        const int kn = min(ATT_K, n - k0);
        for (int i = tid; i < kn * ATT_HD; i += 128) {
            const int j = i >> 6, d = i & 63;
            sK[j * (ATT_HD + 1) + d] = base[(size_t) (k0 + j) * 3 * D + D + d];
        }
        __syncthreads();
        // a thread scores key jj and, where the tile has that many, key jj + 64 (the test is warp-uniform: 32 consecutive keys per warp)
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            const int kj = jj + 64 * half;
            if ((kj & ~31) >= kn) break;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kr = kj < kn ? kj : 0;
#pragma unroll 4
            for (int d = 0; d < ATT_HD; d += 4) {      // 4 key values + 8 broadcast 16-byte query loads per 32 FMAs
                const float k0v = sK[kr * (ATT_HD + 1) + d], k1v = sK[kr * (ATT_HD + 1) + d + 1], k2v = sK[kr * (ATT_HD + 1) + d + 2], k3v = sK[kr * (ATT_HD + 1) + d + 3];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float4 q4 = *reinterpret_cast<const float4 *>(sQ + (qg + u) * ATT_HD + d);
                    acc[u] = fmaf(q4.w, k3v, fmaf(q4.z, k2v, fmaf(q4.y, k1v, fmaf(q4.x, k0v, acc[u]))));
                }
            }
            if (kj < kn) {
#pragma unroll
                for (int u = 0; u < 8; u++) sS[(qg + u) * n_pad + k0 + kj] = acc[u] * scale;
            }
        }
    }
    __syncthreads();
    // ---- softmax: warp per query (ggml_compute_forward_soft_max_f32: max, exp, double sum, scale)
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int qi = warp; qi < nq; qi += 4) {
            float mx = -INFINITY;
            for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sS[qi * n_pad + j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            double sum = 0.0;
            for (int j = lane; j < n; j += 32) { const float e = expf(sS[qi * n_pad + j] - mx); sS[qi * n_pad + j] = e; sum += (double) e; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = (float) (1.0 / sum);
            for (int j = lane; j < n; j += 32) sS[qi * n_pad + j] *= inv;
        }
    }
    // ---- PV: thread = output dim jj for queries qg..qg+7
    float out[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < n; k0 += ATT_K) {
        __syncthreads();
        const int kn = min(ATT_K, n - k0);
        for (int i = tid; i < kn * ATT_HD; i += 128) {
            const int j = i >> 6, d = i & 63;
            sK[j * ATT_HD + d] = base[(size_t) (k0 + j) * 3 * D + 2 * D + d];
        }
        __syncthreads();
        int j = 0;
        for (; j + 4 <= kn; j += 4) {                  // n_pad % 4 == 0 and k0 % 4 == 0: the probability rows are 16-byte readable
            const float v0 = sK[j * ATT_HD + jj], v1 = sK[(j + 1) * ATT_HD + jj], v2 = sK[(j + 2) * ATT_HD + jj], v3 = sK[(j + 3) * ATT_HD + jj];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float4 p4 = *reinterpret_cast<const float4 *>(sS + (qg + u) * n_pad + k0 + j);
                out[u] = fmaf(p4.w, v3, fmaf(p4.z, v2, fmaf(p4.y, v1, fmaf(p4.x, v0, out[u]))));
            }
        }
        for (; j < kn; j++) {
            const float vv = sK[j * ATT_HD + jj];
#pragma unroll
            for (int u = 0; u < 8; u++) out[u] = fmaf(sS[(qg + u) * n_pad + k0 + j], vv, out[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
        if (qg + u < nq) outH[((size_t) b * Lmax + q0 + qg + u) * ldoh + h * ATT_HD + jj] = __float2half_rn(out[u]);
}

__global__ void duration_tail_kernel(const float * __restrict__ logits, int ldl, int n_bins, int B, int Lmax, const int * __restrict__ len,
                                     float * lens_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = i / Lmax, t = i - b * Lmax;
    if (b >= B || t >= len[b]) return;
    const float * row = logits + ((size_t) b * Lmax + t) * ldl;
    double s = 0.0;                                       // ggml_compute_forward_sum_rows -> ggml_vec_sum_f32 accumulates in ggml_float
    for (int j = 0; j < n_bins; j++) s += (double) (1.0f / (1.0f + expf(-row[j])));
    float v = (float) s;
    v = (float) ((int) (v + 0.5f));                       // ggml_round (ggml-cpu.c:1797)
    v = fminf(fmaxf(v, 1.0f), 50.0f);                     // ggml_clamp [1,50] (model.cpp:1040)
    lens_out[(size_t) b * Lmax + t] = v;
}

__global__ void build_alignment_kernel(const float * __restrict__ lens, int B, int LmaxTok, const int * __restrict__ ntok, int LmaxFrames,
                                       int * idx, int * T) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int pos = 0;
    for (int i = 0; i < ntok[b]; i++) {
        const int d = (int) (unsigned) lens[(size_t) b * LmaxTok + i];     // (uint32_t) cast, model.cpp:1286
        if (idx) for (int k = 0; k < d && pos + k < LmaxFrames; k++) idx[(size_t) b * LmaxFrames + pos + k] = i;
        pos += d;
    }
    T[b] = pos;
}

__global__ void curve_conv_s2_kernel(const float * __restrict__ x, int ldx, int B, const int * __restrict__ lenOut, int LmaxOut,
                                     const int * __restrict__ lenIn, float w0, float w1, float w2, float bias, float * dst, int ldd, int coff) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = i / LmaxOut, t = i - b * LmaxOut;
    if (b >= B || t >= lenOut[b]) return;
    const float * xb = x + (size_t) b * ldx;
    const int Li = lenIn[b];
    // F16 kernel -> im2col in fp16 (ggml.c:3878): activations re-rounded to fp16, fp32 accumulate over the 3 taps
    float a = 0.f;
    const int p0 = 2 * t - 1;
    const float ws[3] = { w0, w1, w2 };
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int ps = p0 + k;
        const float xv = (ps >= 0 && ps < Li) ? __half2float(__float2half_rn(xb[ps])) : 0.f;
        a = fmaf(xv, ws[k], a);
    }
    dst[((size_t) b * LmaxOut + t) * ldd + coff] = a + bias;
}

static dim3 row_block(int Cpad, int & rows_per_blk) {
    int nx = Cpad >= 256 ? 256 : (Cpad <= 32 ? 32 : round_up(Cpad, 32));
    if (nx > 256) nx = 256;
    int ny = 256 / nx;
    if (ny < 1) ny = 1;
    rows_per_blk = ny;
    return dim3(nx, ny);
}

}  // namespace

int inorm_stats(Ctx * ctx, const float * x, int ldx, int C, int B, int Lmax, const int * len, double * sums) {
    if (C > KMAX * 256) { set_error("inorm_stats: C=%d too large", C); return 1; }
    B2_CUDA(cudaMemsetAsync(sums, 0, (size_t) B * C * 2 * sizeof(double), ctx->stream));
    if (ldx % 4 == 0 && ldx >= round_up(C, 4) && (((uintptr_t) x) & 15) == 0) {
        const int groups = cdiv(C, 4), bx = groups < 256 ? groups : 256, by = 256 / bx > 0 ? 256 / bx : 1;
        dim3 grid(cdiv(Lmax, V4_ROWS), B, cdiv(groups, bx)), blk(bx, by);
        ctx->prof_begin(PROF_NORM, 0.0, (double) B * Lmax * C * 4.0);
        inorm_stats4_kernel<<<grid, blk, 0, ctx->stream>>>(x, ldx, C, Lmax, len, sums);
        ctx->prof_end();
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    int rpb; dim3 blk = row_block(C, rpb);
    const int rows_per_block = 32;
    dim3 grid(cdiv(Lmax, rows_per_block), B);
    ctx->prof_begin(PROF_NORM, 0.0, (double) B * Lmax * C * 4.0);
    inorm_stats_kernel<<<grid, blk, 0, ctx->stream>>>(x, ldx, C, Lmax, len, sums, rows_per_block);
    ctx->prof_end();
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int adain_apply(Ctx * ctx, const AdainParams & p) {
    if (p.C > KMAX * 256 || p.Cpad > KMAX * 256) { set_error("adain_apply: C=%d too large", p.C); return 1; }
    if (!p.outH && !p.outF) { set_error("adain_apply: no output"); return 1; }
    {
        const int cw = p.outH ? p.Cpad : p.C;
        const bool al = (((uintptr_t) p.x) & 15) == 0 && (p.outH == nullptr || ((((uintptr_t) p.outH) & 7) == 0 && p.ldoh % 4 == 0)) &&
                        (p.outF == nullptr || ((((uintptr_t) p.outF) & 15) == 0 && p.ldof % 4 == 0));
        if (cw % 4 == 0 && p.ldx % 4 == 0 && p.ldx >= round_up(p.C, 4) && al) {
            const int groups = cw / 4, bx = groups < 256 ? groups : 256, by = 256 / bx > 0 ? 256 / bx : 1;
            // long sequences: 256 rows per block so the per-block statistics preamble (double-precision mean / variance) is amortised
            const int rpb4 = p.Lmax >= 8192 ? 256 : V4_ROWS;
            dim3 grid(cdiv(p.Lmax, rpb4), p.B, cdiv(groups, bx)), blk(bx, by);
            ctx->prof_begin(PROF_NORM, 0.0, (double) p.B * p.Lmax * p.C * (4.0 + (p.outH ? 2.0 : 0.0) + (p.outF ? 4.0 : 0.0)));
            // the generator's snake applies (fp16 operand only) get the fully specialised instantiation; the decoder's leaky-relu ones too
            auto launch = [&](auto ragged, auto act, auto oh, auto of) {
                adain_apply4_kernel<decltype(ragged)::value, decltype(act)::value, decltype(oh)::value, decltype(of)::value><<<grid, blk, 0, ctx->stream>>>(p, rpb4);
            };
            using T = std::true_type; using F = std::false_type;
            const bool rg = p.C % 4 != 0, oh = p.outH != nullptr, of = p.outF != nullptr;
            auto by_out = [&](auto ragged, auto act) {
                if (oh && !of) launch(ragged, act, T{}, F{}); else if (!oh && of) launch(ragged, act, F{}, T{}); else launch(ragged, act, T{}, T{});
            };
            auto by_act = [&](auto ragged) {
                if (p.act == NACT_SNAKE) by_out(ragged, std::integral_constant<int, NACT_SNAKE>{});
                else if (p.act == NACT_LRELU02) by_out(ragged, std::integral_constant<int, NACT_LRELU02>{});
                else by_out(ragged, std::integral_constant<int, NACT_NONE>{});
            };
            if (rg) by_act(T{}); else by_act(F{});
            ctx->prof_end();
            B2_LAUNCH_CHECK(ctx);
            return 0;
        }
    }
    int rpb; dim3 blk = row_block(p.outH ? p.Cpad : p.C, rpb);
    const int rows_per_block = 16;
    dim3 grid(cdiv(p.Lmax, rows_per_block), p.B);
    ctx->prof_begin(PROF_NORM, 0.0, (double) p.B * p.Lmax * p.C * (4.0 + (p.outH ? 2.0 : 0.0) + (p.outF ? 4.0 : 0.0)));
    adain_apply_kernel<<<grid, blk, 0, ctx->stream>>>(p, rows_per_block);
    ctx->prof_end();
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int pool_convt(Ctx * ctx, const float * x, int ldx, int C, int B, int Lmax, const int * len, const float * w3, const float * bias, __half * outH,
               int ldoh, int Cpad) {
    int rpb; dim3 blk = row_block(Cpad, rpb);
    dim3 grid(cdiv(2 * Lmax, rpb), B);
    pool_convt_kernel<<<grid, blk, 0, ctx->stream>>>(x, ldx, C, Lmax, len, w3, bias, outH, ldoh, Cpad);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int row_norm(Ctx * ctx, const RowNormParams & p) {
    if (p.C > 1024) { set_error("row_norm: C=%d > 1024", p.C); return 1; }
    const int64_t warps = (int64_t) p.B * p.Lmax;
    row_norm_kernel<<<cdiv(warps * 32, 256), 256, 0, ctx->stream>>>(p);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int cast_rows(Ctx * ctx, const float * x, int ldx, int C, int B, int LmaxIn, const int * lenOut, int LmaxOut, int up2, float ns, __half * outH,
              int ldoh, int Cpad) {
    if (C % 4 == 0 && Cpad % 4 == 0 && ldx % 4 == 0 && ldoh % 4 == 0 && Cpad / 4 <= 256 && ((((uintptr_t) x) & 15) == 0) && ((((uintptr_t) outH) & 7) == 0)) {
        const int bx = Cpad / 4, by = 256 / bx > 0 ? 256 / bx : 1;
        const int rows = LmaxOut >= 8192 ? 256 : 64;
        dim3 grid4(cdiv(LmaxOut, rows), B), blk4(bx, by);
        cast_rows4_kernel<<<grid4, blk4, 0, ctx->stream>>>(x, ldx, C, LmaxIn, lenOut, LmaxOut, up2, ns, outH, ldoh, Cpad, rows);
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    int rpb; dim3 blk = row_block(Cpad, rpb);
    dim3 grid(cdiv(LmaxOut, rpb), B);
    cast_rows_kernel<<<grid, blk, 0, ctx->stream>>>(x, ldx, C, LmaxIn, lenOut, LmaxOut, up2, ns, outH, ldoh, Cpad);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

// block = 32 channels x 8 partial groups: group g sums tiles g, g+8, ...; the 8 group sums are combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256) stats_finalize_kernel(const float * __restrict__ part, int n_mt, int C, double * sums) {
    __shared__ double sh[8][32][2];
    const int b = blockIdx.y;
    const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    double s = 0.0, q = 0.0;
    if (c < C) {
        const float * p = part + ((size_t) b * n_mt * C + c) * 2;
        for (int m = g; m < n_mt; m += 8) { const float2 v = *reinterpret_cast<const float2 *>(p + (size_t) m * C * 2); s += (double) v.x; q += (double) v.y; }
    }
    sh[g][cx][0] = s; sh[g][cx][1] = q;
    __syncthreads();
    if (g == 0 && c < C) {
#pragma unroll
        for (int k = 1; k < 8; k++) { s += sh[k][cx][0]; q += sh[k][cx][1]; }
        sums[((size_t) b * C + c) * 2] = s;
        sums[((size_t) b * C + c) * 2 + 1] = q;
    }
}

__global__ void zero_rows_past_end_kernel(__half * A, int lda, int C, int Lmax, const int * __restrict__ len) {
    const int b = blockIdx.y;
    const int nrow = Lmax - len[b];
    for (int r = blockIdx.x; r < nrow; r += gridDim.x) {
        __half * row = A + ((size_t) b * Lmax + len[b] + r) * lda;
        for (int c = threadIdx.x; c < C; c += blockDim.x) row[c] = __float2half_rn(0.f);
    }
}
int zero_rows_past_end(Ctx * ctx, __half * A, int lda, int C, int B, int Lmax, const int * len) {
    dim3 grid(64, B);
    zero_rows_past_end_kernel<<<grid, 64, 0, ctx->stream>>>(A, lda, C, Lmax, len);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int stats_finalize(Ctx * ctx, const float * part, int B, int n_mt, int C, double * sums) {
    dim3 grid(cdiv(C, 32), B);
    stats_finalize_kernel<<<grid, 256, 0, ctx->stream>>>(part, n_mt, C, sums);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int split3_rows(Ctx * ctx, const float * x, int ldx, int C, int B, int LmaxIn, const int * len, float ns, __half * outH, int Lq) {
    int rpb; dim3 blk = row_block(C, rpb);
    dim3 grid(cdiv(Lq, rpb), B);
    split3_rows_kernel<<<grid, blk, 0, ctx->stream>>>(x, ldx, C, LmaxIn, len, ns, outH, Lq);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int copy_cols(Ctx * ctx, const float * src, int lds, int scoff, float * dst, int ldd, int dcoff, int C, int B, int Lmax, const int * len) {
    int rpb; dim3 blk = row_block(C, rpb);
    dim3 grid(cdiv(Lmax, rpb), B);
    copy_cols_kernel<<<grid, blk, 0, ctx->stream>>>(src, lds, scoff, dst, ldd, dcoff, C, Lmax, len);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int bcast_cols(Ctx * ctx, const float * v, int ldv, int C, int B, int Lmax, const int * len, float * dstF, int ldf, int cofff, __half * dstH,
               int ldh, int coffh) {
    int rpb; dim3 blk = row_block(C, rpb);
    dim3 grid(cdiv(Lmax, rpb), B);
    bcast_cols_kernel<<<grid, blk, 0, ctx->stream>>>(v, ldv, C, Lmax, len, dstF, ldf, cofff, dstH, ldh, coffh);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int gather_rows(Ctx * ctx, const float * src, int lds, int LmaxSrc, const int * idx, int C, int B, int Lmax, const int * len, float * dstF, int ldf,
                __half * dstH, int ldh, int Cpad) {
    int rpb; dim3 blk = row_block(Cpad, rpb);
    dim3 grid(cdiv(Lmax, rpb), B);
    gather_rows_kernel<<<grid, blk, 0, ctx->stream>>>(src, lds, LmaxSrc, idx, C, Lmax, len, dstF, ldf, dstH, ldh, Cpad);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int linear_f32(Ctx * ctx, const float * x, int ldx, const float * W, const float * bias, int rows, int K, int N, float * y, int ldy) {
    if (K == 128 && ldx % 4 == 0 && ((((uintptr_t) x) | ((uintptr_t) W)) & 15) == 0) {
        const int npb = 64;
        dim3 grid(cdiv(N, npb), cdiv(rows, 32));
        linear_f32_k128_kernel<<<grid, 256, 0, ctx->stream>>>(x, ldx, W, bias, rows, N, y, ldy, npb);
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    const int64_t warps = (int64_t) rows * ((N + 31) / 32);
    linear_f32_kernel<<<cdiv(warps * 32, 256), 256, 0, ctx->stream>>>(x, ldx, W, bias, rows, K, N, y, ldy);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int albert_embed(Ctx * ctx, const int * tokens, const int * tok_off, const float * tok_embd, const float * pos_embd, const float * type_embd,
                 const float * nw, const float * nb, int B, int Lmax, const int * len, float * out, int ldo) {
    const int64_t warps = (int64_t) B * Lmax;
    albert_embed_kernel<<<cdiv(warps * 32, 256), 256, 0, ctx->stream>>>(tokens, tok_off, tok_embd, pos_embd, type_embd, nw, nb, B, Lmax, len, out, ldo);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int embed_rows_h(Ctx * ctx, const int * tokens, const int * tok_off, const __half * table, int C, int B, int Lmax, const int * len, __half * outH,
                 int ldoh) {
    int rpb; dim3 blk = row_block(C, rpb);
    dim3 grid(cdiv(Lmax, rpb), B);
    embed_rows_h_kernel<<<grid, blk, 0, ctx->stream>>>(tokens, tok_off, table, C, Lmax, len, outH, ldoh);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int albert_attention(Ctx * ctx, const float * qkv, int B, int Lmax, const int * len, int heads, int hd, float scale, __half * outH, int ldoh) {
    if (hd > 128) { set_error("albert_attention: head dim %d > 128", hd); return 1; }
    if (hd == ATT_HD) {
        const int n_pad = round_up(Lmax, 4) + 4;   // multiple of 4 floats (16-byte row reads), not of 32 (bank spread)
        const size_t smem64 = (size_t) (ATT_Q * ATT_HD + ATT_K * (ATT_HD + 1) + ATT_Q * n_pad) * sizeof(float);
        if (smem64 <= 200 * 1024) {
            static bool attr_set = false;
            if (!attr_set) { cudaFuncSetAttribute(albert_attention64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
            dim3 grid64(cdiv(Lmax, ATT_Q), heads, B);
            albert_attention64_kernel<<<grid64, 128, smem64, ctx->stream>>>(qkv, Lmax, len, heads, scale, outH, ldoh, n_pad);
            B2_LAUNCH_CHECK(ctx);
            return 0;
        }
    }
    const size_t smem = (size_t) (8 * hd + 8 * Lmax) * sizeof(float);
    dim3 grid(cdiv(Lmax, 8), heads, B);
    albert_attention_kernel<<<grid, 256, smem, ctx->stream>>>(qkv, Lmax, len, heads, hd, scale, outH, ldoh);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int duration_tail(Ctx * ctx, const float * logits, int ldl, int n_bins, int B, int Lmax, const int * len, float * lens_out) {
    duration_tail_kernel<<<cdiv((int64_t) B * Lmax, 128), 128, 0, ctx->stream>>>(logits, ldl, n_bins, B, Lmax, len, lens_out);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int build_alignment(Ctx * ctx, const float * lens, int B, int LmaxTok, const int * ntok, int LmaxFrames, int * idx, int * T) {
    build_alignment_kernel<<<cdiv(B, 32), 32, 0, ctx->stream>>>(lens, B, LmaxTok, ntok, LmaxFrames, idx, T);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

int curve_conv_s2(Ctx * ctx, const float * x, int ldx, int B, const int * lenOut, int LmaxOut, const int * lenIn, const float * w3,
                  const float * bias, float * dst, int ldd, int coff) {
    curve_conv_s2_kernel<<<cdiv((int64_t) B * LmaxOut, 128), 128, 0, ctx->stream>>>(x, ldx, B, lenOut, LmaxOut, lenIn, w3[0], w3[1], w3[2], bias[0],
                                                                                    dst, ldd, coff);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b2
