// ar_kernels.cuh -- plain CUDA-core kernels shared by the autoregressive decode paths (orpheus.cu, parler.cu, dia.cu).
//
// First correct versions: fp32 weights and activations as the reference computes them (ggml_mul_mat on F32 operands), one launch per op.
// Each kernel states the ggml op it restates.  Their LOGIC is checked under the CPU emulation in tests/emu (tests/test_emu_cpu.py);
// the -m gpu tests are the parity gate.  Included into each model's translation unit (internal linkage).
#pragma once
#include "kokoro.h"   // HostTensor, ArW
#include <cstdlib>
#include <cstring>

namespace b2 {
namespace {

// rows of a step: (sequence, position, token).  Decode steps have one row per sequence, built on the device from the last argmax.
__global__ void decode_rows_kernel(const int * __restrict__ n_prompt, const int * __restrict__ cur_tok, int B, const int * __restrict__ d_step, int Tmax, int * row_seq, int * row_pos,
                                   int * row_tok, int * row_base, int * row_len) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int pos = n_prompt[b] + *d_step - 1;          // the step number is device-resident so that one captured graph of a step can be replayed
    row_seq[b] = b; row_pos[b] = pos; row_tok[b] = cur_tok[b];
    row_base[b] = b * Tmax; row_len[b] = pos + 1;      // causal: the new row sees its sequence's cache up to and including itself
}

__global__ void embed_kernel(const int * __restrict__ row_tok, const float * __restrict__ embed, int H, float * __restrict__ x) {   // ggml_get_rows
    const int r = blockIdx.x;
    const float * src = embed + (size_t) row_tok[r] * H;
    for (int c = threadIdx.x; c < H; c += blockDim.x) x[(size_t) r * H + c] = src[c];
}

// ggml_rms_norm (float squares accumulated in a double, scale = 1/sqrtf(mean + eps)) followed by the weight multiply (model.cpp:122-125)
__global__ void rmsnorm_kernel(const float * __restrict__ x, const float * __restrict__ w, int H, int R, float * __restrict__ y) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float * row = x + (size_t) r * H;
    double s = 0.0;
    if ((H & 3) == 0) {                                         // 16-byte loads, 8 in flight per lane (the scalar lane-strided loop was a chain of H / 32 dependent round trips)
        const float4 * row4 = reinterpret_cast<const float4 *>(row);
        const int n4 = H >> 2;
        for (int j0 = lane; j0 < n4; j0 += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = j0 + 32 * u < n4 ? row4[j0 + 32 * u] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; u++) s += (double) (v[u].x * v[u].x) + (double) (v[u].y * v[u].y) + (double) (v[u].z * v[u].z) + (double) (v[u].w * v[u].w);
        }
    } else {
        for (int c = lane; c < H; c += 32) s += (double) (row[c] * row[c]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = (float) (s / (double) H);
    const float scale = 1.0f / sqrtf(mean + 1e-5f);
    for (int c = lane; c < H; c += 32) y[(size_t) r * H + c] = (row[c] * scale) * w[c];
}

// ggml's GELU for F32 tensors: an fp16 lookup table of the tanh approximation (ggml-cpu.c:1816-1830)
__device__ __forceinline__ float gelu_f16lut(float x) {
    if (x <= -10.0f) return 0.0f;
    if (x >= 10.0f) return x;
    const float xh = __half2float(__float2half_rn(x));
    return __half2float(__float2half_rn(0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh)))));
}

// where a GEMV puts its results.  act = 1: ggml_gelu on the product (the op that follows fc1); row_dst: output row r goes to row row_dst[r] of Y (the K / V
// projections write straight into their cache slots: ggml_cpy into the cache views); res (indexed by r, may alias Y when there is no row_dst) is added last.
struct GemvOut { const float * res; float * Y; const int * row_dst; int ldy; int act; };
__device__ __forceinline__ void gemv_store(const GemvOut & o, int r, int n, float a) {
    if (o.act == 1) a = gelu_f16lut(a);
    const size_t row = o.row_dst ? (size_t) o.row_dst[r] : (size_t) r;
    o.Y[row * o.ldy + n] = o.res ? a + o.res[(size_t) r * o.ldy + n] : a;
}
// up to three matrices of one dtype against the same activation rows in ONE launch (q / k / v, gate / up): segment j owns the blocks [blk0_j, blk0_j+1)
struct GemvSeg { const void * W; const void * W2; const void * W3; GemvOut o; int N; int blk0; };      // W2 / W3: split low halves (tensor-core F32) or quantised scales / fifth bits
struct GemvGroup { GemvSeg s[3]; int n; };
__device__ __forceinline__ GemvSeg gemv_group_pick(const GemvGroup & g, int blk) {
    const int j = (g.n > 2 && blk >= g.s[2].blk0) ? 2 : ((g.n > 1 && blk >= g.s[1].blk0) ? 1 : 0);
    return j == 2 ? g.s[2] : (j == 1 ? g.s[1] : g.s[0]);
}

// Y[r][n] = sum_k X[r][k] * W[n][k] (+ res[r][n]; res may alias Y: each element is read and written by the same thread).  A warp owns GN (1, 2 or 4) output rows
// and walks K in lane-strided 16-byte steps; every activation load is reused for the 4 weight rows and every weight load for the 8 batch rows of a chunk
// (GN = 4: 2 activation loads per weight load instead of 8 -- the one-row-per-warp form is LSU-bound long before HBM; GN shrinks for small N so that the grid
// still covers the 148 SMs).  Per output the summation order is
// lane-strided k, then the xor-shuffle tree -- independent of GN.  ggml_mul_mat with F32 weights and activations; K % 4 == 0.
constexpr int GR = 8;
template <typename WT, bool ROUND_X, int GN>
__device__ __forceinline__ void gemv_rows_body(const float * __restrict__ X, int ldx, const WT * __restrict__ W, int K, int N, int R, const GemvOut out, int blk) {
    const int n0 = (blk * 8 + (threadIdx.x >> 5)) * GN, lane = threadIdx.x & 31;
    if (n0 >= N) return;
    for (int r0 = 0; r0 < R; r0 += GR) {
        float acc[GN][GR];
#pragma unroll
        for (int i = 0; i < GN; i++)
#pragma unroll
            for (int j = 0; j < GR; j++) acc[i][j] = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
            float w[GN][4];
#pragma unroll
            for (int i = 0; i < GN; i++) {
                const int n = n0 + i < N ? n0 + i : N - 1;                 // rows past N recompute the last row and are not stored
                if constexpr (sizeof(WT) == 4) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(W + (size_t) n * K + k);
                    w[i][0] = w4.x; w[i][1] = w4.y; w[i][2] = w4.z; w[i][3] = w4.w;
                } else {
                    const __half2 * wp = reinterpret_cast<const __half2 *>(W + (size_t) n * K + k);
                    const float2 w01 = __half22float2(wp[0]), w23 = __half22float2(wp[1]);
                    w[i][0] = w01.x; w[i][1] = w01.y; w[i][2] = w23.x; w[i][3] = w23.y;
                }
            }
#pragma unroll
            for (int j = 0; j < GR; j++) {
                if (r0 + j < R) {
                    float4 x4 = *reinterpret_cast<const float4 *>(X + (size_t) (r0 + j) * ldx + k);
                    if constexpr (ROUND_X) {                               // F16 matrix: ggml_mul_mat rounds the activations to fp16 first
                        x4.x = __half2float(__float2half_rn(x4.x)); x4.y = __half2float(__float2half_rn(x4.y));
                        x4.z = __half2float(__float2half_rn(x4.z)); x4.w = __half2float(__float2half_rn(x4.w));
                    }
#pragma unroll
                    for (int i = 0; i < GN; i++) acc[i][j] = fmaf(x4.w, w[i][3], fmaf(x4.z, w[i][2], fmaf(x4.y, w[i][1], fmaf(x4.x, w[i][0], acc[i][j]))));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < GN; i++)
#pragma unroll
            for (int j = 0; j < GR; j++) {
                float a = acc[i][j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0 && r0 + j < R && n0 + i < N) gemv_store(out, r0 + j, n0 + i, a);
            }
    }
}
template <int GN>
__global__ void __launch_bounds__(256) gemv_rows_kernel(const float * __restrict__ X, int ldx, const float * __restrict__ W, int K, int N, int R,
                                                        const float * res, float * Y, int ldy) {
    gemv_rows_body<float, false, GN>(X, ldx, W, K, N, R, GemvOut{res, Y, nullptr, ldy, 0}, (int) blockIdx.x);
}
// the same product for an F16 weight matrix: ggml_mul_mat converts the activation rows to fp16 first (the vec_dot_type of F16 is F16, ggml-cpu.c
// mul_mat from_float) and accumulates the exact fp16 x fp16 products in fp32 -- also what halves the bytes streamed per step
template <int GN>
__global__ void __launch_bounds__(256) gemv_rows_h_kernel(const float * __restrict__ X, int ldx, const __half * __restrict__ W, int K, int N, int R,
                                                          const float * res, float * Y, int ldy) {
    gemv_rows_body<__half, true, GN>(X, ldx, W, K, N, R, GemvOut{res, Y, nullptr, ldy, 0}, (int) blockIdx.x);
}
template <typename WT, bool ROUND_X, int GN>
__global__ void __launch_bounds__(256) gemv_rows_group_kernel(const float * __restrict__ X, int ldx, int K, int R, const GemvGroup g) {
    const GemvSeg s = gemv_group_pick(g, (int) blockIdx.x);
    gemv_rows_body<WT, ROUND_X, GN>(X, ldx, (const WT *) s.W, K, s.N, R, s.o, (int) blockIdx.x - s.blk0);
}
// rows per warp for N outputs: as many as still give every SM a block (148 SMs x 8 warps)
static inline int gemv_rows_gn(int N) {
    static const int forced = [] { const char * e = getenv("B2TTS_GEMV_GN"); return e ? atoi(e) : 0; }();      // 1 / 2 / 4: A/B runs and tests
    if (forced == 1 || forced == 2 || forced == 4) return forced;
    return N >= 148 * 8 * 4 ? 4 : (N >= 148 * 8 * 2 ? 2 : 1);
}
static inline void gemv_rows_launch(cudaStream_t st, const float * X, int ldx, const void * W, bool f16, int K, int N, int R, const float * res, float * Y, int ldy) {
    const int gn = gemv_rows_gn(N), grid = cdiv(N, 8 * gn);
    if (f16) {
        const __half * w = (const __half *) W;
        if (gn == 4) gemv_rows_h_kernel<4><<<grid, 256, 0, st>>>(X, ldx, w, K, N, R, res, Y, ldy);
        else if (gn == 2) gemv_rows_h_kernel<2><<<grid, 256, 0, st>>>(X, ldx, w, K, N, R, res, Y, ldy);
        else gemv_rows_h_kernel<1><<<grid, 256, 0, st>>>(X, ldx, w, K, N, R, res, Y, ldy);
    } else {
        const float * w = (const float *) W;
        if (gn == 4) gemv_rows_kernel<4><<<grid, 256, 0, st>>>(X, ldx, w, K, N, R, res, Y, ldy);
        else if (gn == 2) gemv_rows_kernel<2><<<grid, 256, 0, st>>>(X, ldx, w, K, N, R, res, Y, ldy);
        else gemv_rows_kernel<1><<<grid, 256, 0, st>>>(X, ldx, w, K, N, R, res, Y, ldy);
    }
}

// ---- block-quantised matrices (Q4_0 / Q5_0 / Q8_0; the reference's `quantize` tool, its perf battery runs Parler as Q5_0 / Q8_0).  ggml_mul_mat quantises the
// activation rows to Q8_0 first (vec_dot_type): per 32 columns d = amax / 127 (stored as fp16), q = round-to-nearest-even(x * 127 / amax) (the AVX2
// quantize_row_q8_0), then per block an exact integer dot product scaled by d_w * d_x (ggml_vec_dot_q{4,5,8}_0_q8_0).  Here: stage 1 quantises the 8 rows of a
// chunk into shared memory once per block; stage 2 is a warp per output row, a lane per weight block, dp4a over the eight 4-byte words of a block.
// The matrix lives in HBM as planes (values / fp16 scales / Q5_0 fifth bits, split at load time from ggml's unaligned 34 / 22 / 18-byte blocks): a lane fetches a
// block's values with aligned 16-byte loads; the weight stream is a quarter to an eighth of the fp32 one.
static inline size_t gemv_q_smem(int K) { return (size_t) GR * K + (size_t) GR * (K / 32) * 4; }
__device__ __forceinline__ void gemv_rows_q_body(const float * __restrict__ X, int ldx, const uint8_t * __restrict__ W, const __half * __restrict__ Ws,
                                                 const unsigned * __restrict__ Wh, int qtype, int K, int N, int R, const GemvOut out, int blk) {
    extern __shared__ __align__(16) float gq_smem[];
    const int nb = K >> 5;
    int * xq = reinterpret_cast<int *>(gq_smem);                    // [GR][2 planes][nb][4] packed int8 activations (plane p = words 4p .. 4p+3 of every block)
    float * xd = gq_smem + (size_t) GR * (K >> 2);                   // [GR][nb] block scales (fp16-rounded)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blk * 8 + warp;
    for (int r0 = 0; r0 < R; r0 += GR) {
        if (r0) __syncthreads();
        for (int i = tid; i < GR * nb; i += 256) {                   // stage 1: one thread quantises one (row, block)
            const int j = i / nb, b = i - j * nb;
            int * q = xq + (size_t) j * (K >> 2) + b * 4;             // words 0..3 of block b; words 4..7 sit nb * 4 further (second plane)
            if (r0 + j >= R) { for (int w = 0; w < 8; w++) q[(w >> 2) * nb * 4 + (w & 3)] = 0; xd[j * nb + b] = 0.f; continue; }
            const float4 * src = reinterpret_cast<const float4 *>(X + (size_t) (r0 + j) * ldx + b * 32);
            float4 v[8];
            float amax = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) { v[w] = src[w]; amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[w].x), fabsf(v[w].y)), fmaxf(fabsf(v[w].z), fabsf(v[w].w)))); }
            const float id = amax != 0.f ? 127.f / amax : 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const int a0 = __float2int_rn(v[w].x * id), a1 = __float2int_rn(v[w].y * id), a2 = __float2int_rn(v[w].z * id), a3 = __float2int_rn(v[w].w * id);
                q[(w >> 2) * nb * 4 + (w & 3)] = (a0 & 0xff) | ((a1 & 0xff) << 8) | ((a2 & 0xff) << 16) | ((a3 & 0xff) << 24);
            }
            xd[j * nb + b] = __half2float(__float2half_rn(amax / 127.f));
        }
        __syncthreads();
        if (n < N) {
            float acc[GR];
#pragma unroll
            for (int j = 0; j < GR; j++) acc[j] = 0.f;
            for (int b = lane; b < nb; b += 32) {
                const size_t bi = (size_t) n * nb + b;
                const float dw = __half2float(Ws[bi]);
                int wq[8];
                if (qtype == 8) {
                    const uint4 v0 = *reinterpret_cast<const uint4 *>(W + bi * 32), v1 = *reinterpret_cast<const uint4 *>(W + bi * 32 + 16);
                    wq[0] = (int) v0.x; wq[1] = (int) v0.y; wq[2] = (int) v0.z; wq[3] = (int) v0.w; wq[4] = (int) v1.x; wq[5] = (int) v1.y; wq[6] = (int) v1.z; wq[7] = (int) v1.w;
                } else {
                    const uint4 v = *reinterpret_cast<const uint4 *>(W + bi * 16);
                    const unsigned q4s[4] = {v.x, v.y, v.z, v.w};
                    const unsigned qh = qtype == 6 ? Wh[bi] : 0u;
#pragma unroll
                    for (int w = 0; w < 4; w++) {                    // bytes 4w .. 4w+3 of the block: low nibbles are elements 4w.., high nibbles elements 16 + 4w..
                        const unsigned q4 = q4s[w];
                        unsigned lo = q4 & 0x0F0F0F0Fu, hi = (q4 >> 4) & 0x0F0F0F0Fu;
                        if (qtype == 6) {
                            const unsigned bl = (qh >> (4 * w)) & 0xFu, bh = (qh >> (16 + 4 * w)) & 0xFu;
                            lo |= ((bl & 1u) << 4) | ((bl & 2u) << 11) | ((bl & 4u) << 18) | ((bl & 8u) << 25);
                            hi |= ((bh & 1u) << 4) | ((bh & 2u) << 11) | ((bh & 4u) << 18) | ((bh & 8u) << 25);
                            wq[w] = (int) __vsub4(lo, 0x10101010u); wq[w + 4] = (int) __vsub4(hi, 0x10101010u);
                        } else {
                            wq[w] = (int) __vsub4(lo, 0x08080808u); wq[w + 4] = (int) __vsub4(hi, 0x08080808u);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < GR; j++) {
                    // two 16-byte loads per (row, block); consecutive lanes read consecutive 16 B of a plane: no bank conflicts (block-major, 32 B per lane, was 8-way)
                    const int4 q0 = *reinterpret_cast<const int4 *>(xq + (size_t) j * (K >> 2) + b * 4), q1 = *reinterpret_cast<const int4 *>(xq + (size_t) j * (K >> 2) + nb * 4 + b * 4);
                    int sumi = __dp4a(wq[0], q0.x, 0);
                    sumi = __dp4a(wq[1], q0.y, sumi); sumi = __dp4a(wq[2], q0.z, sumi); sumi = __dp4a(wq[3], q0.w, sumi);
                    sumi = __dp4a(wq[4], q1.x, sumi); sumi = __dp4a(wq[5], q1.y, sumi); sumi = __dp4a(wq[6], q1.z, sumi); sumi = __dp4a(wq[7], q1.w, sumi);
                    acc[j] = fmaf((float) sumi, dw * xd[j * nb + b], acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < GR; j++) {
                float a = acc[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0 && r0 + j < R) gemv_store(out, r0 + j, n, a);
            }
        }
    }
}
__global__ void __launch_bounds__(256) gemv_rows_q_kernel(const float * __restrict__ X, int ldx, const uint8_t * __restrict__ W, const __half * __restrict__ Ws,
                                                          const unsigned * __restrict__ Wh, int qtype, int K, int N, int R, const float * res, float * Y, int ldy) {
    gemv_rows_q_body(X, ldx, W, Ws, Wh, qtype, K, N, R, GemvOut{res, Y, nullptr, ldy, 0}, (int) blockIdx.x);
}
__global__ void __launch_bounds__(256) gemv_rows_q_group_kernel(const float * __restrict__ X, int ldx, int qtype, int K, int R, const GemvGroup g) {
    const GemvSeg s = gemv_group_pick(g, (int) blockIdx.x);
    gemv_rows_q_body(X, ldx, (const uint8_t *) s.W, (const __half *) s.W2, (const unsigned *) s.W3, qtype, K, s.N, R, s.o, (int) blockIdx.x - s.blk0);
}

// ---- tensor-core batched GEMV for F16 matrices: the decode step of a batch of <= 16 sequences.
// At batch 16 an F16 weight byte carries 16 flops: 6.6 TB/s of weights would need ~105 TFLOP/s of fp32 FMA, above what the CUDA cores deliver, and the plain
// kernel above issues 8 activation loads per weight load.  Here the batch IS the M = 16 of mma.sync.m16n8k16 (exact fp16 products, fp32 accumulation -- the
// reference's numerics for an F16 matrix): a block owns 8 output rows and splits K over its 8 warps; every lane streams 16 contiguous bytes of its weight row per
// step (8 rows x 64 B per warp instruction, whole sectors), the activations sit in shared memory as fp16 (rounded once per block), and the warps' partial
// 16 x 8 tiles are summed in a fixed order.  The k index is permuted consistently on both operands (a dot product does not care): the 8 halves a lane loads
// are k-slots {2t, 2t+1, 2t+8, 2t+9} of two consecutive k16 steps, so no ldmatrix / transposition is needed.  tcgen05 would bring nothing here: its M is 64+.
#ifdef B2EMU
static inline void mma16816_f16f32(float * c, const unsigned * a, unsigned b0, unsigned b1) {      // functional model of the PTX fragment layout (tests/emu)
    unsigned mine[6] = {a[0], a[1], a[2], a[3], b0, b1}, all[32][6];
    b2emu::warp_exchange(mine, 6, &all[0][0]);
    const int lane = b2emu_lane(), g = lane >> 2, t = lane & 3;
    auto h = [](unsigned v, int i) { __half_raw r; r.x = (unsigned short) (i ? v >> 16 : v & 0xffff); return __half2float(__half(r)); };
    for (int ci = 0; ci < 4; ci++) {
        const int row = g + (ci >> 1) * 8, col = 2 * t + (ci & 1);
        float acc = c[ci];
        for (int tp = 0; tp < 4; tp++)                          // k = 2tp, 2tp+1 (a0/a1, b0) and 2tp+8, 2tp+9 (a2/a3, b1)
            for (int hi = 0; hi < 2; hi++)
                for (int i = 0; i < 2; i++) {
                    const unsigned av = all[(row & 7) * 4 + tp][(row >> 3) + 2 * hi], bv = all[col * 4 + tp][4 + hi];
                    acc = fmaf(h(av, i), h(bv, i), acc);
                }
        c[ci] = acc;
    }
}
#else
__device__ __forceinline__ void mma16816_f16f32(float * c, const unsigned * a, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#endif

constexpr int GM_PAD = 32;     // halves of padding per activation row in shared memory: consecutive rows start 16 banks apart, so the 8 lanes of a quarter warp
                               // (rows g, g+1 x 4 lanes x 16 B) cover all 32 banks once per 128-bit load (8 halves = 4 banks made every such load a 2-way conflict)
constexpr int GM_KC  = 2048;   // activations are staged through shared memory in K chunks of this many columns (16 rows x 2048 x fp16 = 64 KB; twice that for SPLIT)
constexpr float GM_LO_SCALE = 2048.0f;   // SPLIT: the low halves are carried scaled by 2^11 so that they stay in fp16's normal range
static inline bool gemv_mma_ok(int K, int N, int R) { (void) R; return K % 256 == 0 && N % 8 == 0; }
// MT = m16 tiles per block (batch rows / 16): the weight fragment of a k-step is reused for all of them, so a batch of 64 still streams W once
static inline int gemv_mma_kc(int K, int mt) { const int kc = GM_KC / mt; return K < kc ? K : kc; }
static inline size_t gemv_mma_smem(int K, bool split, int mt) { return (size_t) 16 * mt * (gemv_mma_kc(K, mt) + GM_PAD) * 2 * (split ? 2 : 1) + (size_t) 8 * 16 * mt * 8 * 4; }
// F16 matrices: on by default since it reproduced the reference's tokens on a B200 (round 2: 9.0 -> 4.0 ms per Parler-Mini step); B2TTS_AR_MMA=0 selects the plain kernels.
// The split (fp32-faithful) form for F32 matrices (Orpheus) doubles the resident weight bytes and stays opt-in: B2TTS_AR_MMA=1.
static inline bool gemv_mma_enabled() { static const bool on = [] { const char * e = getenv("B2TTS_AR_MMA"); return !(e && e[0] == '0'); }(); return on; }
static inline bool gemv_split_mma_enabled() { static const bool on = [] { const char * e = getenv("B2TTS_AR_MMA"); return e && e[0] == '1'; }(); return on; }

// D k-steps of 32: all D weight loads are issued before the first mma so that a lane keeps D x 16 B (SPLIT: 2 x D x 16 B) of the weight stream in flight
// (read once: streaming hint).  SPLIT: the fp32-faithful product of an F32 matrix -- x = xh + xl, W = Wh + Wl in fp16 pairs, x.W ~ xh.Wh + (xl.Wh + xh.Wl)
// with the two cross terms in their own accumulator (scaled by 2^11); the dropped xl.Wl term and the rounding of the low halves are ~2^-22 relative.
// tile: halves between the activation rows of consecutive m16 tiles in shared memory.
template <int D, bool SPLIT, int MT> __device__ __forceinline__ void gm_chunk(float (*c)[4], float (*cl)[4], const __half * w, const __half * wl, const __half * xa, const __half * xb,
                                                                              const __half * xla, const __half * xlb, size_t tile) {
    uint4 wv[D], wlv[SPLIT ? D : 1];
#pragma unroll
    for (int j = 0; j < D; j++) { wv[j] = __ldcs(reinterpret_cast<const uint4 *>(w + 32 * j)); if constexpr (SPLIT) wlv[j] = __ldcs(reinterpret_cast<const uint4 *>(wl + 32 * j)); }
#pragma unroll
    for (int j = 0; j < D; j++) {
#pragma unroll
        for (int m = 0; m < MT; m++) {
            const uint4 a = *reinterpret_cast<const uint4 *>(xa + m * tile + 32 * j), b = *reinterpret_cast<const uint4 *>(xb + m * tile + 32 * j);
            const unsigned f0[4] = {a.x, b.x, a.y, b.y}, f1[4] = {a.z, b.z, a.w, b.w};
            mma16816_f16f32(c[m], f0, wv[j].x, wv[j].y);
            mma16816_f16f32(c[m], f1, wv[j].z, wv[j].w);
            if constexpr (SPLIT) {
                const uint4 la = *reinterpret_cast<const uint4 *>(xla + m * tile + 32 * j), lb = *reinterpret_cast<const uint4 *>(xlb + m * tile + 32 * j);
                const unsigned l0[4] = {la.x, lb.x, la.y, lb.y}, l1[4] = {la.z, lb.z, la.w, lb.w};
                mma16816_f16f32(cl[m], l0, wv[j].x, wv[j].y);      // xl . Wh
                mma16816_f16f32(cl[m], l1, wv[j].z, wv[j].w);
                mma16816_f16f32(cl[m], f0, wlv[j].x, wlv[j].y);    // xh . Wl
                mma16816_f16f32(cl[m], f1, wlv[j].z, wlv[j].w);
            }
        }
    }
}

template <bool SPLIT, int MT>
__device__ __forceinline__ void gemv_mma_body(const float * __restrict__ X, int ldx, const __half * __restrict__ W, const __half * __restrict__ Wl, int K, int N, int R,
                                              const GemvOut out, int blk) {
    extern __shared__ __align__(16) float gm_smem[];
    constexpr int ROWS = 16 * MT;
    const int kc = K < GM_KC / MT ? K : GM_KC / MT, pitch = kc + GM_PAD;
    __half * sX = reinterpret_cast<__half *>(gm_smem);                                   // [ROWS][kc + GM_PAD] fp16-rounded activations of the current K chunk (rows >= R are zero)
    __half * sXl = sX + (SPLIT ? (size_t) ROWS * pitch : 0);                              // SPLIT: their low halves, scaled by 2^11
    float * red = reinterpret_cast<float *>(sX + (size_t) ROWS * pitch * (SPLIT ? 2 : 1));  // [8 warps][ROWS][8] partial tiles
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n0 = blk * 8;
    float c[MT][4], cl[MT][4];
#pragma unroll
    for (int m = 0; m < MT; m++) { c[m][0] = c[m][1] = c[m][2] = c[m][3] = 0.f; cl[m][0] = cl[m][1] = cl[m][2] = cl[m][3] = 0.f; }
    const __half * wrow = W + (size_t) (n0 + g) * K + t * 8;
    const __half * wlrow = SPLIT ? Wl + (size_t) (n0 + g) * K + t * 8 : nullptr;
    const __half * xa = sX + (size_t) g * pitch + t * 8, * xb = sX + (size_t) (g + 8) * pitch + t * 8;
    const __half * xla = sXl + (size_t) g * pitch + t * 8, * xlb = sXl + (size_t) (g + 8) * pitch + t * 8;
    const size_t tile = (size_t) 16 * pitch;
    for (int k0 = 0; k0 < K; k0 += kc) {
        const int kn = K - k0 < kc ? K - k0 : kc;                                         // a multiple of 256, like K and GM_KC / MT
        if (k0) __syncthreads();                                                         // every warp is done with the previous chunk
        // 8 independent 16-byte loads per thread in flight (measured on a B200: one load per iteration made this staging pass a chain of 16 dependent L2 round
        // trips, ~9 of the kernel's ~14 us at batch 16)
        for (int i0 = tid * 4; i0 < ROWS * kn; i0 += 256 * 4 * 8) {
            float4 vv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * 256 * 4, r = i / kn, k = i - r * kn;
                vv[u] = (i < ROWS * kn && r < R) ? *reinterpret_cast<const float4 *>(X + (size_t) r * ldx + k0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * 256 * 4, r = i / kn, k = i - r * kn;
                if (i >= ROWS * kn) continue;
                const float4 v = vv[u];
                const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
                __half2 * d = reinterpret_cast<__half2 *>(sX + (size_t) r * pitch + k);
                d[0] = h01; d[1] = h23;
                if constexpr (SPLIT) {
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    __half2 * dl = reinterpret_cast<__half2 *>(sXl + (size_t) r * pitch + k);
                    dl[0] = __floats2half2_rn((v.x - f01.x) * GM_LO_SCALE, (v.y - f01.y) * GM_LO_SCALE);
                    dl[1] = __floats2half2_rn((v.z - f23.x) * GM_LO_SCALE, (v.w - f23.y) * GM_LO_SCALE);
                }
            }
        }
        __syncthreads();
        const int ks = kn >> 3, kbeg = warp * ks, kend = kbeg + ks;                        // this warp's slice of the chunk (a multiple of 32)
        int k = kbeg;
        if constexpr (!SPLIT && MT == 1) for (; k + 256 <= kend; k += 256) gm_chunk<8, false, 1>(c, cl, wrow + k0 + k, nullptr, xa + k, xb + k, nullptr, nullptr, tile);   // 8 x 16 B in flight per lane
        if constexpr (MT <= 2) for (; k + 128 <= kend; k += 128) gm_chunk<4, SPLIT, MT>(c, cl, wrow + k0 + k, SPLIT ? wlrow + k0 + k : nullptr, xa + k, xb + k, xla + k, xlb + k, tile);
        for (; k + 64 <= kend; k += 64) gm_chunk<2, SPLIT, MT>(c, cl, wrow + k0 + k, SPLIT ? wlrow + k0 + k : nullptr, xa + k, xb + k, xla + k, xlb + k, tile);
        for (; k < kend; k += 32) gm_chunk<1, SPLIT, MT>(c, cl, wrow + k0 + k, SPLIT ? wlrow + k0 + k : nullptr, xa + k, xb + k, xla + k, xlb + k, tile);
    }
    float * my = red + (size_t) warp * ROWS * 8;                                         // c0,c1: row g, cols 2t,2t+1;  c2,c3: row g+8 (of tile m)
#pragma unroll
    for (int m = 0; m < MT; m++) {
        if constexpr (SPLIT) { for (int i = 0; i < 4; i++) c[m][i] += cl[m][i] * (1.0f / GM_LO_SCALE); }
        float * mm = my + m * 128;
        mm[g * 8 + 2 * t] = c[m][0]; mm[g * 8 + 2 * t + 1] = c[m][1]; mm[(g + 8) * 8 + 2 * t] = c[m][2]; mm[(g + 8) * 8 + 2 * t + 1] = c[m][3];
    }
    __syncthreads();
    for (int i = tid; i < ROWS * 8; i += 256) {
        const int r = i >> 3, col = i & 7;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) a += red[(size_t) w * ROWS * 8 + i];
        if (r < R && n0 + col < N) gemv_store(out, r, n0 + col, a);
    }
}
template <bool SPLIT, int MT>
__global__ void __launch_bounds__(256) gemv_mma_kernel(const float * __restrict__ X, int ldx, const __half * __restrict__ W, const __half * __restrict__ Wl, int K, int N, int R,
                                                       const float * res, float * Y, int ldy) {
    gemv_mma_body<SPLIT, MT>(X, ldx, W, Wl, K, N, R, GemvOut{res, Y, nullptr, ldy, 0}, (int) blockIdx.x);
}
template <bool SPLIT, int MT>
__global__ void __launch_bounds__(256) gemv_mma_group_kernel(const float * __restrict__ X, int ldx, int K, int R, const GemvGroup g) {
    const GemvSeg s = gemv_group_pick(g, (int) blockIdx.x);
    gemv_mma_body<SPLIT, MT>(X, ldx, (const __half *) s.W, (const __half *) s.W2, K, s.N, R, s.o, (int) blockIdx.x - s.blk0);
}

template <bool SPLIT, int MT>
static inline int gemv_mma_launch_t(Ctx * ctx, cudaStream_t st, size_t & smem_set, const float * X, int ldx, const __half * W, const __half * Wl, int K, int N, int R, const float * res, float * Y,
                                    int ldy) {
    const size_t smem = gemv_mma_smem(K, SPLIT, MT);
    if (smem > smem_set) { B2_CUDA(cudaFuncSetAttribute(gemv_mma_kernel<SPLIT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); smem_set = smem; }
    gemv_mma_kernel<SPLIT, MT><<<N / 8, 256, smem, st>>>(X, ldx, W, Wl, K, N, R, res, Y, ldy);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

// R rows (any number: chunks of up to 64) of X against an F16 matrix (Wl == nullptr) or the fp16 (hi, scaled lo) split of an F32 matrix; 0 launched, 1 error.
// smem_set: six counters of the caller (one per kernel instantiation) remembering the dynamic shared memory already opted into.
static inline int gemv_mma_launch(Ctx * ctx, cudaStream_t st, size_t * smem_set, const float * X, int ldx, const __half * W, const __half * Wl, int K, int N, int R,
                                  const float * res, float * Y, int ldy) {
    const bool split = Wl != nullptr;
    for (int r0 = 0; r0 < R; r0 += 64) {
        const int rn = R - r0 < 64 ? R - r0 : 64;
        const float * x = X + (size_t) r0 * ldx, * rs = res ? res + (size_t) r0 * ldy : nullptr;
        float * y = Y + (size_t) r0 * ldy;
        int rc;
        if (rn <= 16)      rc = split ? gemv_mma_launch_t<true, 1>(ctx, st, smem_set[0], x, ldx, W, Wl, K, N, rn, rs, y, ldy) : gemv_mma_launch_t<false, 1>(ctx, st, smem_set[1], x, ldx, W, Wl, K, N, rn, rs, y, ldy);
        else if (rn <= 32) rc = split ? gemv_mma_launch_t<true, 2>(ctx, st, smem_set[2], x, ldx, W, Wl, K, N, rn, rs, y, ldy) : gemv_mma_launch_t<false, 2>(ctx, st, smem_set[3], x, ldx, W, Wl, K, N, rn, rs, y, ldy);
        else               rc = split ? gemv_mma_launch_t<true, 4>(ctx, st, smem_set[4], x, ldx, W, Wl, K, N, rn, rs, y, ldy) : gemv_mma_launch_t<false, 4>(ctx, st, smem_set[5], x, ldx, W, Wl, K, N, rn, rs, y, ldy);
        if (rc) return 1;
    }
    return 0;
}

// NeoX RoPE over the whole head with per-pair frequency factors (ggml_rope_ext mode 2, theta base 5e5, ggml-cpu rope cache: theta starts at
// the position and is multiplied by theta_scale pair after pair), applied to q in place and to k on its way into the cache; v is copied
// (orpheus_build_kv_store, model.cpp:196-228 -- here the cache is compact: the 3x head expansion is done by indexing in the attention)
// ff may be null (no frequency factors: Dia); the cache row of r is row_dst[r] when given, else row_seq[r] * Tmax + row_pos[r].
__global__ void rope_append_kernel(float * q, const float * __restrict__ k, const float * __restrict__ v, const float * __restrict__ ff, const int * __restrict__ row_seq,
                                   const int * __restrict__ row_pos, int heads, int kv_heads, int hd, float theta_scale, float * Kc, float * Vc, int Tmax,
                                   const int * __restrict__ row_dst) {
    const int r = blockIdx.x, h = blockIdx.y;                  // h < heads: a query head; h >= heads: kv head h - heads
    const int pos = row_pos[r], half = hd >> 1;
    const size_t dst = row_dst ? (size_t) row_dst[r] : (size_t) row_seq[r] * Tmax + pos;
    const int KV = kv_heads * hd, H = heads * hd;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float theta = (float) pos;
        for (int j = 0; j < i; j++) theta *= theta_scale;
        const float th = ff ? theta / ff[i] : theta;
        const float c = cosf(th), s = sinf(th);
        if (h < heads) {
            float * p = q + (size_t) r * H + (size_t) h * hd;
            const float x0 = p[i], x1 = p[i + half];
            p[i] = x0 * c - x1 * s; p[i + half] = x0 * s + x1 * c;
        } else {
            const int kh = h - heads;
            const float * p = k + (size_t) r * KV + (size_t) kh * hd;
            float * d = Kc + dst * KV + (size_t) kh * hd;
            const float x0 = p[i], x1 = p[i + half];
            d[i] = x0 * c - x1 * s; d[i + half] = x0 * s + x1 * c;
            const float * pv = v + (size_t) r * KV + (size_t) kh * hd;
            float * dv = Vc + dst * KV + (size_t) kh * hd;
            dv[i] = pv[i]; dv[i + half] = pv[i + half];
        }
    }
}

// attention of one query row over cache positions [row_base[r], row_base[r] + row_len[r]) of Kc / Vc (rows of kv_heads * hd floats): softmax(q.K^T * scale) V
// with ggml_soft_max's double-accumulated sum (ggml-cpu.c soft_max: max, expf, ggml_float sum, scale by (float)(1/sum)).  A causal or block mask of
// -inf entries is the same as restricting the range: exp(-inf) adds an exact zero.  Query head h reads kv head h / (heads / kv_heads) -- the
// reference's repeat-interleaved GQA cache (orpheus model.cpp:196-228, dia model.cpp:426-437) without the copies.  Tcap >= max row_len sizes the scores.
__global__ void __launch_bounds__(128) attention_kernel(const float * __restrict__ q, const float * __restrict__ Kc, const float * __restrict__ Vc,
                                                        const int * __restrict__ row_base, const int * __restrict__ row_len, int heads, int kv_heads, int hd,
                                                        int Tcap, float scale, float * __restrict__ out) {
    extern __shared__ float sc[];          // [T] scores, then 128 floats + 128 doubles of reduction scratch behind them
    const int r = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    const size_t base = (size_t) row_base[r];
    const int T = row_len[r];
    const int KV = kv_heads * hd, H = heads * hd, kh = h / (heads / kv_heads);
    float * redf = sc + ((Tcap + 1) & ~1);      // keeps the double scratch behind it 8-byte aligned
    double * redd = reinterpret_cast<double *>(redf + 128);
    const float * qv = q + (size_t) r * H + (size_t) h * hd;
    float mx = -INFINITY;
    for (int t = tid; t < T; t += 128) {
        const float * kr = Kc + (base + t) * KV + (size_t) kh * hd;
        float a = 0.f;
        for (int d = 0; d < hd; d += 4) {        // 16-B loads: each thread walks its own cache row, so wide loads are what keeps the L1 sector efficiency up
            const float4 k4 = *reinterpret_cast<const float4 *>(kr + d);
            a = fmaf(qv[d + 3], k4.w, fmaf(qv[d + 2], k4.z, fmaf(qv[d + 1], k4.y, fmaf(qv[d], k4.x, a))));
        }
        a *= scale;
        sc[t] = a;
        mx = fmaxf(mx, a);
    }
    redf[tid] = mx;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (tid < o) redf[tid] = fmaxf(redf[tid], redf[tid + o]); __syncthreads(); }
    mx = redf[0];
    double sum = 0.0;
    for (int t = tid; t < T; t += 128) { const float e = expf(sc[t] - mx); sc[t] = e; sum += (double) e; }
    redd[tid] = sum;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (tid < o) redd[tid] += redd[tid + o]; __syncthreads(); }
    const float inv = (float) (1.0 / redd[0]);
    for (int d = tid; d < hd; d += 128) {
        float a = 0.f;
        for (int t = 0; t < T; t++) a = fmaf(sc[t] * inv, Vc[(base + t) * KV + (size_t) kh * hd + d], a);
        out[(size_t) r * H + (size_t) h * hd + d] = a;
    }
}
// The same attention for ALL the query heads that share one kv head (grid: rows x kv_heads): every K and V row of the compact cache is read from HBM once per
// kv head instead of once per query head (3x less cache traffic for Orpheus' 24 / 8 heads, 4x for Dia), and the P.V product uses the whole block (16 threads x
// 16 B cover a V row, 8 slices of positions reduced in a fixed order).  Scores, maxima and the double-accumulated softmax sums are computed exactly as in
// attention_kernel (same per-thread order, same trees); only the P.V summation order differs.  REP = heads / kv_heads <= 8, hd % 4 == 0, hd <= 128.
constexpr int ATT_MAX_REP = 8;
__global__ void __launch_bounds__(128) attention_gqa_kernel(const float * __restrict__ q, const float * __restrict__ Kc, const float * __restrict__ Vc,
                                                            const int * __restrict__ row_base, const int * __restrict__ row_len, int heads, int kv_heads, int hd,
                                                            int Tcap, float scale, float * __restrict__ out) {
    extern __shared__ __align__(16) float ag_smem[];
    const int rep = heads / kv_heads, Tp = (Tcap + 3) & ~3;
    float * sc = ag_smem;                                      // [rep][Tp] scores, then probabilities
    float * sq = sc + (size_t) rep * Tp;                        // [rep][hd] the query heads of this group
    float * redf = sq + (size_t) rep * hd;                      // [128] floats
    double * redd = reinterpret_cast<double *>(redf + 128);     // [128] doubles (8-byte aligned: every segment above is a multiple of 4 floats, and rep*Tp, rep*hd even)
    float * pv = reinterpret_cast<float *>(redd + 128);         // [8 slices][rep][hd] partial outputs
    const int r = blockIdx.x, kh = blockIdx.y, tid = threadIdx.x;
    const size_t base = (size_t) row_base[r];
    const int T = row_len[r];
    const int KV = kv_heads * hd, H = heads * hd;
    for (int i = tid; i < rep * hd; i += 128) sq[i] = q[(size_t) r * H + (size_t) kh * rep * hd + i];      // query heads kh*rep .. kh*rep+rep-1 are contiguous
    __syncthreads();
    for (int t = tid; t < T; t += 128) {
        const float * kr = Kc + (base + t) * KV + (size_t) kh * hd;
        float a[ATT_MAX_REP];
#pragma unroll
        for (int j = 0; j < ATT_MAX_REP; j++) a[j] = 0.f;
        for (int d = 0; d < hd; d += 4) {
            const float4 k4 = *reinterpret_cast<const float4 *>(kr + d);
#pragma unroll
            for (int j = 0; j < ATT_MAX_REP; j++)
                if (j < rep) { const float * qj = sq + j * hd + d; a[j] = fmaf(qj[3], k4.w, fmaf(qj[2], k4.z, fmaf(qj[1], k4.y, fmaf(qj[0], k4.x, a[j])))); }
        }
#pragma unroll
        for (int j = 0; j < ATT_MAX_REP; j++) if (j < rep) sc[(size_t) j * Tp + t] = a[j] * scale;
    }
    for (int j = 0; j < rep; j++) {                             // per query head: max, exp, double sum -- the reductions of attention_kernel
        float mloc = -INFINITY;
        for (int t = tid; t < T; t += 128) mloc = fmaxf(mloc, sc[(size_t) j * Tp + t]);      // a thread re-reads the scores it wrote itself
        redf[tid] = mloc;
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) { if (tid < o) redf[tid] = fmaxf(redf[tid], redf[tid + o]); __syncthreads(); }
        const float m = redf[0];
        double sum = 0.0;
        for (int t = tid; t < T; t += 128) { const float e = expf(sc[(size_t) j * Tp + t] - m); sc[(size_t) j * Tp + t] = e; sum += (double) e; }
        redd[tid] = sum;
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) { if (tid < o) redd[tid] += redd[tid + o]; __syncthreads(); }
        const float inv = (float) (1.0 / redd[0]);
        __syncthreads();
        for (int t = tid; t < T; t += 128) sc[(size_t) j * Tp + t] *= inv;
    }
    __syncthreads();
    // P.V: thread (slice, d4) walks positions slice, slice + 8, ... and 4 consecutive channels; a V row is read once for all rep heads
    const int nd4 = hd >> 2, d4 = tid % 16, slice = tid / 16;   // 16 x 8 threads; channels beyond 64 are covered by looping d4 in steps of 16
    for (int dd = d4; dd < nd4; dd += 16) {
        float acc[ATT_MAX_REP][4];
#pragma unroll
        for (int j = 0; j < ATT_MAX_REP; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
        for (int t = slice; t < T; t += 8) {
            const float4 v4 = *reinterpret_cast<const float4 *>(Vc + (base + t) * KV + (size_t) kh * hd + dd * 4);
#pragma unroll
            for (int j = 0; j < ATT_MAX_REP; j++)
                if (j < rep) {
                    const float p = sc[(size_t) j * Tp + t];
                    acc[j][0] = fmaf(p, v4.x, acc[j][0]); acc[j][1] = fmaf(p, v4.y, acc[j][1]); acc[j][2] = fmaf(p, v4.z, acc[j][2]); acc[j][3] = fmaf(p, v4.w, acc[j][3]);
                }
        }
#pragma unroll
        for (int j = 0; j < ATT_MAX_REP; j++)
            if (j < rep) { float * d = pv + ((size_t) slice * rep + j) * hd + dd * 4; d[0] = acc[j][0]; d[1] = acc[j][1]; d[2] = acc[j][2]; d[3] = acc[j][3]; }
    }
    __syncthreads();
    for (int i = tid; i < rep * hd; i += 128) {
        float a = 0.f;
#pragma unroll
        for (int sl = 0; sl < 8; sl++) a += pv[(size_t) sl * rep * hd + i];
        out[(size_t) r * H + (size_t) kh * rep * hd + i] = a;
    }
}
static inline size_t attention_gqa_smem_bytes(int Tcap, int rep, int hd) { return ((size_t) rep * ((Tcap + 3) & ~3) + (size_t) rep * hd + 128) * 4 + 128 * 8 + (size_t) 8 * rep * hd * 4; }
static inline bool attention_gqa_ok(int heads, int kv_heads, int hd, int Tcap) {
    return kv_heads > 0 && heads % kv_heads == 0 && heads / kv_heads <= ATT_MAX_REP && hd % 4 == 0 && hd <= 128 && ((heads / kv_heads) * hd) % 2 == 0 &&
           attention_gqa_smem_bytes(Tcap, heads / kv_heads, hd) <= 200 * 1024;
}

// B2TTS_AR_ATT=plain selects attention_kernel (one block per query head) for A/B runs; the grouped kernel is the default
static inline bool attention_gqa_enabled() { static const bool on = [] { const char * e = getenv("B2TTS_AR_ATT"); return !(e && e[0] == 'p'); }(); return on; }

static inline size_t attention_smem_bytes(int Tcap) { return (size_t) ((Tcap + 1) & ~1) * 4 + 128 * 4 + 128 * 8; }

__global__ void silu_mul_kernel(float * g, const float * __restrict__ u, size_t n) {      // ggml_silu (x / (1 + expf(-x))) * up
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float x = g[i]; g[i] = (x / (1.0f + expf(-x))) * u[i]; }
}

__global__ void gather_rows_f32_kernel(const float * __restrict__ x, const int * __restrict__ idx, int H, float * __restrict__ y) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < H; c += blockDim.x) y[(size_t) b * H + c] = x[(size_t) idx[b] * H + c];
}

// sampler::max: the first maximum wins
__global__ void __launch_bounds__(256) argmax_kernel(const float * __restrict__ logits, int V, int * cur_tok, int * out_tokens, int n_steps, const int * __restrict__ d_step) {
    __shared__ float sv[256]; __shared__ int si[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float * lg = logits + (size_t) b * V;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 256) { const float v = lg[i]; if (v > best) { best = v; bi = i; } }
    sv[tid] = best; si[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { if (sv[tid + o] > sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; } }
        __syncthreads();
    }
    if (tid == 0) { const int t = si[0] == 0x7fffffff ? 0 : si[0]; cur_tok[b] = t; out_tokens[(size_t) b * n_steps + *d_step] = t; }      // (all-NaN logits: token 0, never an out-of-range id)
}


// ggml_norm (ggml-cpu.c:7114-7163: mean and variance of the centred row accumulated in double, scale = 1/sqrtf(var + eps)) followed by the
// weight multiply and bias add (parler model.cpp build_norm)
__global__ void layernorm_kernel(const float * __restrict__ x, const float * __restrict__ w, const float * __restrict__ bias, int H, int R, float eps, float * __restrict__ y) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float * row = x + (size_t) r * H;
    const bool vec = (H & 3) == 0;
    const float4 * row4 = reinterpret_cast<const float4 *>(row);
    const int n4 = H >> 2;
    double s = 0.0;
    if (vec) {                                                  // 16-byte loads, 8 in flight per lane
        for (int j0 = lane; j0 < n4; j0 += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = j0 + 32 * u < n4 ? row4[j0 + 32 * u] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; u++) s += (double) v[u].x + (double) v[u].y + (double) v[u].z + (double) v[u].w;
        }
    } else {
        for (int c = lane; c < H; c += 32) s += (double) row[c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = (float) (s / (double) H);
    double s2 = 0.0;
    if (vec) {
        for (int j0 = lane; j0 < n4; j0 += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = j0 + 32 * u < n4 ? row4[j0 + 32 * u] : make_float4(mean, mean, mean, mean);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float a = v[u].x - mean, b = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
                s2 += (double) (a * a) + (double) (b * b) + (double) (c * c) + (double) (d * d);
            }
        }
    } else {
        for (int c = lane; c < H; c += 32) { const float v = row[c] - mean; s2 += (double) (v * v); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    const float var = (float) (s2 / (double) H);
    const float scale = 1.0f / sqrtf(var + eps);
    for (int c = lane; c < H; c += 32) y[(size_t) r * H + c] = ((row[c] - mean) * scale) * w[c] + bias[c];
}

// x[r] = table[tok[r]] + pos_table[pos[r]] (prompt rows: ggml_get_rows + ggml_add of the positional rows)
__global__ void embed_pos_kernel(const int * __restrict__ row_tok, const int * __restrict__ row_pos, const float * __restrict__ table, const float * __restrict__ pos_table, int H,
                                 float * __restrict__ x) {
    const int r = blockIdx.x;
    const float * src = table + (size_t) row_tok[r] * H;
    const float * ps = pos_table + (size_t) row_pos[r] * H;
    for (int c = threadIdx.x; c < H; c += blockDim.x) x[(size_t) r * H + c] = src[c] + ps[c];
}

// the audio-token input of a decode step: the rows of the n_out codebook tables summed in head order ((e0 + e1) + e2 ...: parler_build_inp_embd,
// model.cpp:320-340; dia model.cpp build_dia_decoder_inp_embd), plus the positional row when there is one.  tables[i] rows are tab_rows apart.
__global__ void codebook_embed_kernel(const int * __restrict__ ids, int n_out, const float * __restrict__ tables, size_t tab_stride, const float * __restrict__ pos_table,
                                      const int * __restrict__ row_pos, int H, float * __restrict__ x) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = tables[(size_t) ids[(size_t) r * n_out] * H + c];
        for (int i = 1; i < n_out; i++) a = tables[(size_t) i * tab_stride + (size_t) ids[(size_t) r * n_out + i] * H + c] + a;
        if (pos_table) a = a + pos_table[(size_t) row_pos[r] * H + c];
        x[(size_t) r * H + c] = a;
    }
}

// NeoX RoPE in place on rows of nh heads (ggml_rope_ext mode 2 without frequency factors; theta advanced by repeated multiplication like the ggml-cpu cache)
__global__ void rope_rows_kernel(float * x, const int * __restrict__ row_pos, int nh, int hd, float theta_scale) {
    const int r = blockIdx.x, h = blockIdx.y, half = hd >> 1;
    const int pos = row_pos[r];
    float * p = x + (size_t) r * nh * hd + (size_t) h * hd;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float theta = (float) pos;
        for (int j = 0; j < i; j++) theta *= theta_scale;
        const float c = cosf(theta), s = sinf(theta);
        const float x0 = p[i], x1 = p[i + half];
        p[i] = x0 * c - x1 * s; p[i + half] = x0 * s + x1 * c;
    }
}

// copy the new k / v rows to their cache slots (row_dst[r] in rows of KV floats): ggml_cpy into the cache views
__global__ void store_kv_kernel(const float * __restrict__ k, const float * __restrict__ v, const int * __restrict__ row_dst, int KV, float * Kc, float * Vc) {
    const int r = blockIdx.x;
    const size_t d = (size_t) row_dst[r] * KV;
    for (int c = threadIdx.x; c < KV; c += blockDim.x) { Kc[d + c] = k[(size_t) r * KV + c]; if (v) Vc[d + c] = v[(size_t) r * KV + c]; }
}

// ggml_gelu in place (the stand-alone pass; the decode paths apply it in fc1's GEMV epilogue, GemvOut::act)
__global__ void gelu_f16lut_kernel(float * g, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = gelu_f16lut(g[i]);
}

// sampler::max over rows of V logits: the first maximum wins.  row -> out[step * gridDim.x + row], step = *d_step (0 when d_step is null)
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float * __restrict__ logits, int V, int * __restrict__ out, const int * __restrict__ d_step) {
    __shared__ float sv[256]; __shared__ int si[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float * lg = logits + (size_t) b * V;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 256) { const float v = lg[i]; if (v > best) { best = v; bi = i; } }
    sv[tid] = best; si[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { if (sv[tid + o] > sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; } }
        __syncthreads();
    }
    if (tid == 0) out[(size_t) (d_step ? *d_step : 0) * gridDim.x + b] = si[0] == 0x7fffffff ? 0 : si[0];                                 // (all-NaN logits: token 0)
}

// rows of an audio decode step under the delay pattern (parler generate_audio_tokens, model.cpp:795-832): output head i is fed BOS until step i + 1, then the
// token it produced in the previous step (d_out [steps][B][n_out]) -- or EOS for good once an EOS of that head is two or more steps old (eos_seen is
// updated by check_stopping at the top of an iteration, after the next batch was already built).  One row per sequence at position first_pos[b] + step.
// seen / stopped (may be null: fixed-length generation): parler_context::eos_seen per head and the step at which check_stopping (model.cpp:715-732: position
// >= max_generation, or every head has produced EOS) would have ended the reference's loop for sequence b.
// The step number lives in device memory (d_step, advanced by step_advance_kernel) so that one captured CUDA graph of a step can be replayed for every step.
__global__ void delay_rows_kernel(const int * __restrict__ d_out, const int * __restrict__ first_pos, int B, int n_out, const int * __restrict__ d_step, int bos, int eos, int max_gen,
                                  int Tmax, int * seen, int * stopped, int * ids, int * row_pos, int * row_base, int * row_len, int * row_dst) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int step = *d_step;
    const int pos = (first_pos ? first_pos[b] : 0) + step;
    const int * last = d_out + ((size_t) (step > 0 ? step - 1 : 0) * B + b) * n_out;
    if (seen && step >= 1 && stopped[b] < 0) {                       // check_stopping at the top of this iteration
        bool stop = pos >= max_gen;
        if (!stop) { stop = true; for (int i = 0; i < n_out; i++) stop = stop && (seen[b * n_out + i] || last[i] == eos); }
        if (stop) stopped[b] = step;
    }
    for (int i = 0; i < n_out; i++) {
        const bool s = seen && seen[b * n_out + i];                  // as of the outputs up to step - 2
        ids[b * n_out + i] = step > i ? (s ? eos : last[i]) : bos;
        if (seen && step >= 1 && last[i] == eos) seen[b * n_out + i] = 1;
    }
    row_pos[b] = pos; row_base[b] = b * Tmax; row_len[b] = pos + 1; row_dst[b] = b * Tmax + pos;
}

__global__ void step_advance_kernel(int * d_step) { if (threadIdx.x == 0 && blockIdx.x == 0) *d_step += 1; }

// ---- host helpers for block-quantised tensors (shared by parler.cu and dia.cu)

// a GGUF tensor of Q4_0 / Q5_0 / Q8_0 blocks -> HostTensor: keeps the blocks in `raw` and dequantises them like dequantize_row_q{4,5,8}_0 (ggml-quants.c) into `v`
// (what ggml_get_rows hands to the embedding users).  0 ok, 1 error (message set).
static inline int host_tensor_from_blocks(HostTensor & t, const char * name, int type, int64_t n, const void * data, size_t nbytes) {
    const size_t blk = type == 2 ? 18 : type == 6 ? 22 : 34;
    if (n % 32 || nbytes < (size_t) n / 32 * blk) { set_error("tensor %s: short or ragged quantised data", name); return 1; }
    t.qtype = type;
    t.raw.assign((const uint8_t *) data, (const uint8_t *) data + (size_t) n / 32 * blk);
    t.v.resize((size_t) n);
    for (int64_t b = 0; b < n / 32; b++) {
        const uint8_t * p = t.raw.data() + (size_t) b * blk;
        __half_raw hr; memcpy(&hr.x, p, 2);
        const float d = __half2float(__half(hr));
        float * y = t.v.data() + (size_t) b * 32;
        if (type == 8) { for (int j = 0; j < 32; j++) y[j] = (float) (int8_t) p[2 + j] * d; }
        else {
            uint32_t qh = 0; if (type == 6) memcpy(&qh, p + 2, 4);
            const uint8_t * qs = p + (type == 6 ? 6 : 2);
            for (int j = 0; j < 16; j++) {
                int x0 = qs[j] & 0x0F, x1 = qs[j] >> 4;
                if (type == 6) { x0 = (x0 | (int) (((qh >> j) & 1u) << 4)) - 16; x1 = (x1 | (int) (((qh >> (j + 16)) & 1u) << 4)) - 16; }
                else { x0 -= 8; x1 -= 8; }
                y[j] = (float) x0 * d; y[j + 16] = (float) x1 * d;
            }
        }
    }
    return 0;
}

// one GGUF tensor (ggml type 0 = F32, 1 = F16; with allow_quant also 2 / 6 / 8 = Q4_0 / Q5_0 / Q8_0 blocks) -> HostTensor: fp32 values in `v` (F16 widened exactly,
// blocks dequantised like ggml), outermost dimension first in `shape`.  The assign_weight of all three decode paths.  0 ok, 1 error (message set).
static inline int host_tensor_from_gguf(HostTensor & t, const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes, bool allow_quant) {
    int64_t n = 1;
    for (int i = n_dims - 1; i >= 0; i--) { t.shape.push_back(ne[i]); n *= ne[i]; }
    t.v.resize((size_t) n);
    if (type == 0) {
        if (nbytes < (size_t) n * 4) { set_error("tensor %s: short data", name); return 1; }
        memcpy(t.v.data(), data, (size_t) n * 4);
    } else if (type == 1) {
        if (nbytes < (size_t) n * 2) { set_error("tensor %s: short data", name); return 1; }
        const uint16_t * s = (const uint16_t *) data;
        for (int64_t i = 0; i < n; i++) { __half_raw r; r.x = s[i]; t.v[(size_t) i] = __half2float(__half(r)); }
        t.f16 = true;
    } else if (allow_quant && (type == 2 || type == 6 || type == 8)) {
        if (host_tensor_from_blocks(t, name, type, n, data, nbytes)) return 1;
    } else {
        set_error(allow_quant ? "tensor %s: ggml type %d not supported (F32, F16, Q4_0, Q5_0, Q8_0)" : "tensor %s: ggml type %d not supported (F32/F16 only)", name, type);
        return 1;
    }
    return 0;
}

// the planes of a block-quantised matrix in HBM: the ggml blocks (fp16 scale | [4 bytes of fifth bits] | 16 or 32 bytes of values, 18 / 22 / 34 bytes, unaligned) split
// into values, scales and fifth bits so that a lane reads a block's values with one aligned 16-byte (two for Q8_0) load; same bytes in total.  false on cudaMalloc failure.
static inline bool upload_quant_planes(const HostTensor & t, ArW & w, std::vector<void *> & dev_allocs, size_t & weight_bytes) {
    w = ArW(); w.qtype = t.qtype;
    const size_t nblk = t.v.size() / 32, blk = t.qtype == 2 ? 18 : t.qtype == 6 ? 22 : 34, vb = t.qtype == 8 ? 32 : 16;
    std::vector<uint8_t> vals(nblk * vb); std::vector<uint16_t> sc(nblk); std::vector<uint32_t> hb(t.qtype == 6 ? nblk : 0);
    for (size_t b = 0; b < nblk; b++) {
        const uint8_t * p = t.raw.data() + b * blk;
        memcpy(&sc[b], p, 2);
        if (t.qtype == 6) memcpy(&hb[b], p + 2, 4);
        memcpy(&vals[b * vb], p + (t.qtype == 6 ? 6 : 2), vb);
    }
    bool ok = true;
    auto put = [&](const void * src, size_t bytes) -> void * {
        void * d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of %zu bytes failed", bytes); ok = false; return nullptr; }
        cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
        dev_allocs.push_back(d); weight_bytes += bytes;
        return d;
    };
    w.p = put(vals.data(), vals.size()); w.scales = put(sc.data(), sc.size() * 2);
    if (t.qtype == 6) w.qh = put(hb.data(), hb.size() * 4);
    return ok;
}

// Y = X . W^T (+ res) for a block-quantised W: 0 launched, 1 error
static inline int gemv_q_launch(Ctx * ctx, cudaStream_t st, size_t & smem_set, const float * X, int ldx, const ArW & W, int K, int N, int R, const float * res, float * Y, int ldy) {
    if (K % 32 || gemv_q_smem(K) > 200 * 1024) { set_error("quantised matrix with K = %d is not supported", K); return 1; }
    const size_t smem = gemv_q_smem(K);
    if (smem > smem_set) { B2_CUDA(cudaFuncSetAttribute(gemv_rows_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); smem_set = smem; }
    gemv_rows_q_kernel<<<cdiv(N, 8), 256, smem, st>>>(X, ldx, (const uint8_t *) W.p, (const __half *) W.scales, (const unsigned *) W.qh, W.qtype, K, N, R, res, Y, ldy);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

// ---- grouped launches: up to three matrices of ONE storage kind against the same activation rows (q / k / v, gate / up), each with its own epilogue (GemvOut).
// The per-output arithmetic is the single-matrix kernels' (same bodies): results are bit-identical to separate launches.
enum GemvKind { GEMV_F32 = 0, GEMV_F16 = 1, GEMV_F16_MMA = 2, GEMV_SPLIT_MMA = 3, GEMV_QUANT = 4 };
struct GemvItem { const void * W; const void * W2; const void * W3; int N; GemvOut o; };      // W2 / W3 as in GemvSeg
struct GemvGroupSmem { size_t mma[6] = {0, 0, 0, 0, 0, 0}; size_t q = 0; };                 // dynamic shared memory already opted into, per group-kernel instantiation

template <bool SPLIT, int MT>
static inline int gemv_mma_group_launch_t(Ctx * ctx, cudaStream_t st, size_t & smem_set, const float * X, int ldx, int K, int R, const GemvGroup & g, int blocks) {
    const size_t smem = gemv_mma_smem(K, SPLIT, MT);
    if (smem > smem_set) { B2_CUDA(cudaFuncSetAttribute(gemv_mma_group_kernel<SPLIT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); smem_set = smem; }
    gemv_mma_group_kernel<SPLIT, MT><<<blocks, 256, smem, st>>>(X, ldx, K, R, g);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}
template <typename WT, bool ROUND_X>
static inline void gemv_rows_group_launch_t(cudaStream_t st, int gn, int blocks, const float * X, int ldx, int K, int R, const GemvGroup & g) {
    if (gn == 4) gemv_rows_group_kernel<WT, ROUND_X, 4><<<blocks, 256, 0, st>>>(X, ldx, K, R, g);
    else if (gn == 2) gemv_rows_group_kernel<WT, ROUND_X, 2><<<blocks, 256, 0, st>>>(X, ldx, K, R, g);
    else gemv_rows_group_kernel<WT, ROUND_X, 1><<<blocks, 256, 0, st>>>(X, ldx, K, R, g);
}

// B2TTS_AR_FUSE=0 turns the fused launches of the decode paths off (separate q / k / v launches, stand-alone KV store and activation passes) for A/B runs
static inline bool ar_fuse_enabled() { static const bool on = [] { const char * e = getenv("B2TTS_AR_FUSE"); return !(e && e[0] == '0'); }(); return on; }

// kind: the storage / kernel family of ALL n items (the caller checks that they agree); qtype for GEMV_QUANT.  0 launched, 1 error.
static inline int gemv_group_launch(Ctx * ctx, cudaStream_t st, GemvGroupSmem & sm, int kind, int qtype, const float * X, int ldx, int K, int R, const GemvItem * it, int n) {
    if (n < 1 || n > 3) { set_error("gemv group of %d matrices", n); return 1; }
    GemvGroup g; g.n = n;
    int total_n = 0;
    for (int i = 0; i < n; i++) total_n += it[i].N;
    auto fill = [&](int rows_per_block, int r0) {             // segments for the row chunk starting at r0; returns the number of blocks
        int blk = 0;
        for (int i = 0; i < n; i++) {
            GemvOut o = it[i].o;
            if (r0) { if (o.res) o.res += (size_t) r0 * o.ldy; if (o.row_dst) o.row_dst += r0; else o.Y += (size_t) r0 * o.ldy; }
            g.s[i] = GemvSeg{it[i].W, it[i].W2, it[i].W3, o, it[i].N, blk};
            blk += cdiv(it[i].N, rows_per_block);
        }
        for (int i = n; i < 3; i++) g.s[i] = g.s[0];
        return blk;
    };
    if (kind == GEMV_QUANT) {
        if (K % 32 || gemv_q_smem(K) > 200 * 1024) { set_error("quantised matrix with K = %d is not supported", K); return 1; }
        const size_t smem = gemv_q_smem(K);
        if (smem > sm.q) { B2_CUDA(cudaFuncSetAttribute(gemv_rows_q_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); sm.q = smem; }
        const int blocks = fill(8, 0);
        gemv_rows_q_group_kernel<<<blocks, 256, smem, st>>>(X, ldx, qtype, K, R, g);
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    if (kind == GEMV_F16_MMA || kind == GEMV_SPLIT_MMA) {
        const bool split = kind == GEMV_SPLIT_MMA;
        for (int r0 = 0; r0 < R; r0 += 64) {                  // row chunks of up to 64 (4 m16 tiles), like gemv_mma_launch
            const int rn = R - r0 < 64 ? R - r0 : 64, blocks = fill(8, r0);
            const float * x = X + (size_t) r0 * ldx;
            int rc;
            if (rn <= 16)      rc = split ? gemv_mma_group_launch_t<true, 1>(ctx, st, sm.mma[0], x, ldx, K, rn, g, blocks) : gemv_mma_group_launch_t<false, 1>(ctx, st, sm.mma[1], x, ldx, K, rn, g, blocks);
            else if (rn <= 32) rc = split ? gemv_mma_group_launch_t<true, 2>(ctx, st, sm.mma[2], x, ldx, K, rn, g, blocks) : gemv_mma_group_launch_t<false, 2>(ctx, st, sm.mma[3], x, ldx, K, rn, g, blocks);
            else               rc = split ? gemv_mma_group_launch_t<true, 4>(ctx, st, sm.mma[4], x, ldx, K, rn, g, blocks) : gemv_mma_group_launch_t<false, 4>(ctx, st, sm.mma[5], x, ldx, K, rn, g, blocks);
            if (rc) return 1;
        }
        return 0;
    }
    const int gn = gemv_rows_gn(total_n), blocks = fill(8 * gn, 0);
    if (kind == GEMV_F16) gemv_rows_group_launch_t<__half, true>(st, gn, blocks, X, ldx, K, R, g);
    else gemv_rows_group_launch_t<float, false>(st, gn, blocks, X, ldx, K, R, g);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

// up to 3 ArW matrices against the same rows: one launch when they share a storage kind, else one launch each
static inline int gemv_group_arw(Ctx * ctx, cudaStream_t st, GemvGroupSmem & sm, const float * X, int ldx, int K, int R, const ArW * const * W, const int * N, const GemvOut * o, int n) {
    auto kind_of = [&](const ArW & w, int Nw) -> int { return w.qtype ? GEMV_QUANT : (w.f16 ? ((gemv_mma_enabled() && gemv_mma_ok(K, Nw, 16)) ? GEMV_F16_MMA : GEMV_F16) : GEMV_F32); };
    GemvItem it[3];
    bool same = true;
    const int k0 = kind_of(*W[0], N[0]);
    for (int i = 0; i < n; i++) {
        it[i] = GemvItem{W[i]->p, W[i]->scales, W[i]->qh, N[i], o[i]};
        same = same && kind_of(*W[i], N[i]) == k0 && W[i]->qtype == W[0]->qtype;
    }
    if (same) return gemv_group_launch(ctx, st, sm, k0, W[0]->qtype, X, ldx, K, R, it, n);
    for (int i = 0; i < n; i++) if (gemv_group_launch(ctx, st, sm, kind_of(*W[i], N[i]), W[i]->qtype, X, ldx, K, R, it + i, 1)) return 1;
    return 0;
}

// ---- the launch helpers the three decode paths share (each model's forward context derives from this)
struct ArLaunch {
    Ctx * ctx = nullptr; cudaStream_t st = nullptr;
    size_t mma_smem_set[6] = {0, 0, 0, 0, 0, 0}, q_smem_set = 0, att_smem_set = 0, gqa_smem_set = 0;      // dynamic shared memory already opted into, per kernel instantiation
    GemvGroupSmem group_smem;
    // softmax(q K^T * scale) V for R rows over their cache ranges: grouped by kv head when the shape allows (K / V read once per kv head), else one block per query head
    int attend(const float * q, const float * Kc, const float * Vc, const int * row_base, const int * row_len, int R, int heads, int kv_heads, int hd, int Tcap, float scale, float * out) {
        if (attention_gqa_enabled() && attention_gqa_ok(heads, kv_heads, hd, Tcap)) {
            const size_t smem = attention_gqa_smem_bytes(Tcap, heads / kv_heads, hd);
            if (smem > gqa_smem_set) { B2_CUDA(cudaFuncSetAttribute(attention_gqa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); gqa_smem_set = smem; }
            dim3 grid(R, kv_heads);
            attention_gqa_kernel<<<grid, 128, smem, st>>>(q, Kc, Vc, row_base, row_len, heads, kv_heads, hd, Tcap, scale, out);
        } else {
            const size_t smem = attention_smem_bytes(Tcap);
            if (smem > 200 * 1024) { set_error("context of %d positions exceeds the attention kernel's shared memory", Tcap); return 1; }
            if (smem > att_smem_set) { B2_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); att_smem_set = smem; }
            dim3 grid(R, heads);
            attention_kernel<<<grid, 128, smem, st>>>(q, Kc, Vc, row_base, row_len, heads, kv_heads, hd, Tcap, scale, out);
        }
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    // Y = X . W^T (+ res) for a matrix in any of its storage kinds
    int gemv(const float * X, int ldx, const ArW & W, int K, int N, int R, const float * res, float * Y, int ldy) {
        if (W.qtype) return gemv_q_launch(ctx, st, q_smem_set, X, ldx, W, K, N, R, res, Y, ldy);      // Q4_0 / Q5_0 / Q8_0: Q8_0-requantised activations, dp4a per block
        if (W.f16 && gemv_mma_enabled() && gemv_mma_ok(K, N, 16))         // tensor-core path: chunks of 16 rows (a decode step of <= 16 sequences is one chunk)
            return gemv_mma_launch(ctx, st, mma_smem_set, X, ldx, (const __half *) W.p, nullptr, K, N, R, res, Y, ldy);
        gemv_rows_launch(st, X, ldx, W.p, W.f16, K, N, R, res, Y, ldy);
        B2_LAUNCH_CHECK(ctx);
        return 0;
    }
    int gemv_group(const float * X, int ldx, int K, int R, const ArW * const * W, const int * N, const GemvOut * o, int n) { return gemv_group_arw(ctx, st, group_smem, X, ldx, K, R, W, N, o, n); }
};

}  // namespace
}  // namespace b2
