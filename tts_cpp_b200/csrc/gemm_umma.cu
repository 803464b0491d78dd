// gemm_umma.cu -- persistent, warp-specialised tcgen05 + TMA implicit-GEMM Conv1d / Linear for sm_100a.
//
// Same contract as conv_gemm_kernel (gemm_conv.cu): out[b][t][co] = epilogue( sum_{k,ci} A[b][t + k*dil - pad][ci] * W[co][k][ci] ),
// fp16 operands, fp32 accumulation -- the reference's ggml_conv_1d / ggml_mul_mat with F16 weights (ggml/src/ggml.c:3870-3894,
// ggml/src/ggml-cpu/ggml-cpu.c:262-267) -- for stride-1 layers with Cin % 64 == 0 and Cout % 128 == 0 (all the heavy ones).
//
// Blackwell mapping:
//   * one CTA per SM, persistent over (utterance, 128-row time tile, Cout tile) work items;
//   * warp 0  = TMA producer: per 64-channel chunk it loads the activation rows ONCE, with the conv halo
//               (128 + (K-1)*dil rows), as 8 un-swizzled [rows][16 B] core-matrix columns; every tap then reads the same
//               shared-memory chunk at a row offset (descriptor start address + tap*dil*16 B) -- no im2col, no re-load per tap.
//               Weight tiles [Cout_tile][64] stream through a 4-stage 128B-swizzled ring, one per (chunk, tap);
//   * warp 1  = MMA issuer: one elected thread issues tcgen05.mma (M=128, N=128|256, K=16) into a TMEM accumulator;
//               tcgen05.commit releases smem stages / publishes the accumulator through mbarriers;
//   * warps 2-5 = epilogue: tcgen05.ld the accumulator (double-buffered in TMEM, so the next tile's MMAs overlap), apply
//               bias / residual adds / divide / activation, store fp32 and/or fp16 rows.
//   * activations outside [0, len_b) must read as zero: TMA zero-fills out-of-range rows, and zero_tail_rows() clears the
//     first rows past each utterance's end for ragged batches.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace b2 {
namespace {

constexpr int UM = 128, BKC = 64, RA_MAX = 192, NSTAGE = 4;
constexpr int A_BUF_BYTES = 8 * RA_MAX * 16;   // 8 core-matrix columns x RA_MAX rows x 16 B
constexpr int UMMA_THREADS = 192;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t * bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void * dst, const CUtensorMap * tm, uint64_t * bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void * dst, const CUtensorMap * tm, uint64_t * bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t * bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t * r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// shared-memory matrix descriptor (K-major operand).  layout_type: 0 = no swizzle, 2 = 128B swizzle; version field = 1 on sm_100
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t) ((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t) 1 << 46;
    d |= (uint64_t) layout_type << 61;
    return d;
}

__device__ __forceinline__ float gelu_f16lut_u(float x) {   // ggml-cpu.c:1816-1830 (same as gemm_conv.cu)
    if (x <= -10.0f) return 0.0f;
    if (x >= 10.0f) return x;
    float xh = __half2float(__float2half_rn(x));
    float y  = 0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh)));
    return __half2float(__float2half_rn(y));
}

struct UmmaExtra {
    int RA;            // activation rows staged per chunk (128 + (KW-1)*dil rounded up to 8)
    int n_mt;          // 128-row tiles per utterance
    int n_nt;          // Cout tiles
    int total_tiles;
    int vec4;          // epilogue may use 16-byte accesses
};

template <int NT>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_umma_kernel(const ConvGemmParams p, const UmmaExtra e, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    constexpr int B_STAGE_BYTES = NT * 128;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char * sB = smem;                                        // [NSTAGE][NT][128 B], 128B-swizzled by TMA
    unsigned char * sA = smem + NSTAGE * B_STAGE_BYTES;               // [2][8][RA][16 B]
    uint64_t * bars = reinterpret_cast<uint64_t *>(sA + 2 * A_BUF_BYTES);
    uint64_t * a_full = bars, * a_empty = bars + 2, * b_full = bars + 4, * b_empty = bars + 4 + NSTAGE;
    uint64_t * acc_full = bars + 4 + 2 * NSTAGE, * acc_empty = acc_full + 2;
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    float * red = reinterpret_cast<float *>(sA + 2 * A_BUF_BYTES + 256);   // [4 epilogue warps][NT][2]: per-tile column sums for the fused InstanceNorm statistics

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_cc = p.CinPad / BKC;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < NSTAGE; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 2 accumulators of NT fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * NT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer
        if (lane == 0) {
            uint32_t a_it = 0, b_it = 0;
            for (int tile = blockIdx.x; tile < e.total_tiles; tile += gridDim.x) {
                const int nt = tile % e.n_nt, rest = tile / e.n_nt, mt = rest % e.n_mt, b = rest / e.n_mt;
                const int t0 = mt * UM, n0 = nt * NT;
                for (int cc = 0; cc < n_cc; cc++) {
                    const int ab = a_it & 1;
                    mbar_wait(&a_empty[ab], ((a_it >> 1) & 1) ^ 1);
                    mbar_expect_tx(&a_full[ab], (uint32_t) e.RA * 128u);
#pragma unroll
                    for (int kc = 0; kc < 8; kc++)
                        tma_load_3d(sA + ab * A_BUF_BYTES + kc * e.RA * 16, &tmA, &a_full[ab], cc * BKC + kc * 8, t0 - p.pad, b);
                    a_it++;
                    for (int tap = 0; tap < p.KW; tap++) {
                        const int s = b_it % NSTAGE;
                        mbar_wait(&b_empty[s], ((b_it / NSTAGE) & 1) ^ 1);
                        mbar_expect_tx(&b_full[s], (uint32_t) B_STAGE_BYTES);
                        tma_load_2d(sB + s * B_STAGE_BYTES, &tmB, &b_full[s], tap * p.CinPad + cc * BKC, n0);
                        b_it++;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        const uint32_t idesc = (1u << 4) | ((uint32_t) (NT >> 3) << 17) | ((uint32_t) (UM >> 4) << 24);   // F32 accum, F16 x F16, K-major A and B
        uint32_t a_it = 0, b_it = 0, acc_it = 0;
        for (int tile = blockIdx.x; tile < e.total_tiles; tile += gridDim.x) {
            const int acb = acc_it & 1;
            mbar_wait(&acc_empty[acb], ((acc_it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t) (acb * NT);
            uint32_t accumulate = 0;
            for (int cc = 0; cc < n_cc; cc++) {
                const int ab = a_it & 1;
                mbar_wait(&a_full[ab], (a_it >> 1) & 1);
                const uint32_t a_base = smem_u32(sA + ab * A_BUF_BYTES);
                for (int tap = 0; tap < p.KW; tap++) {
                    const int s = b_it % NSTAGE;
                    mbar_wait(&b_full[s], (b_it / NSTAGE) & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t b_base = smem_u32(sB + s * B_STAGE_BYTES);
                        const uint32_t a_tap = a_base + (uint32_t) (tap * p.dil) * 16u;
#pragma unroll
                        for (int j = 0; j < BKC / 16; j++) {
                            const uint64_t ad = smem_desc(a_tap + (uint32_t) (2 * j * e.RA) * 16u, (uint32_t) e.RA * 16u, 128u, 0u);
                            const uint64_t bd = smem_desc(b_base + (uint32_t) j * 32u, 16u, 1024u, 2u);
                            umma_f16(d_tmem, ad, bd, idesc, accumulate);
                            accumulate = 1;
                        }
                        umma_commit(&b_empty[s]);
                    }
                    __syncwarp();
                    b_it++;
                }
                if (lane == 0) umma_commit(&a_empty[ab]);
                __syncwarp();
                a_it++;
            }
            if (lane == 0) umma_commit(&acc_full[acb]);
            __syncwarp();
            acc_it++;
        }
    } else {
        // ================================================================= epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t acc_it = 0;
        for (int tile = blockIdx.x; tile < e.total_tiles; tile += gridDim.x) {
            const int nt = tile % e.n_nt, rest = tile / e.n_nt, mt = rest % e.n_mt, b = rest / e.n_mt;
            const int t = mt * UM + row, n0 = nt * NT;
            const int lo = p.lenOut ? p.lenOut[b] : p.LmaxOut;
            const bool valid = t < lo && t < p.LmaxOut;
            const size_t r = (size_t) b * p.LmaxOut + t;
            const int acb = acc_it & 1;
            mbar_wait(&acc_full[acb], (acc_it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (acb * NT);
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t) c0, v);
                const int cb = n0 + c0;
                const bool on = valid && cb < p.N;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
                const bool vec = e.vec4 && cb + 32 <= p.N;
                if (on) {
                    if (vec) {
                        if (p.bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) { const float4 bb = *reinterpret_cast<const float4 *>(p.bias + cb + j); f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w; }
                        }
                        if (p.add1) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) { const float4 a = *reinterpret_cast<const float4 *>(p.add1 + r * p.ldadd1 + cb + j); f[j] = a.x + f[j]; f[j + 1] = a.y + f[j + 1]; f[j + 2] = a.z + f[j + 2]; f[j + 3] = a.w + f[j + 3]; }
                        }
                        if (p.add2) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) { const float4 a = *reinterpret_cast<const float4 *>(p.add2 + r * p.ldadd2 + cb + j); f[j] = a.x + f[j]; f[j + 1] = a.y + f[j + 1]; f[j + 2] = a.z + f[j + 2]; f[j + 3] = a.w + f[j + 3]; }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int c = cb + j;
                            if (c < p.N) {
                                if (p.bias) f[j] = f[j] + p.bias[c];
                                if (p.add1) f[j] = p.add1[r * p.ldadd1 + c] + f[j];
                                if (p.add2) f[j] = p.add2[r * p.ldadd2 + c] + f[j];
                            }
                        }
                    }
                    if (p.div != 0.f) {
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = __fdiv_rn(f[j], p.div);
                    }
                    if (p.act == ACT_GELU_F16LUT) {
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = gelu_f16lut_u(f[j]);
                    } else if (p.act == ACT_LRELU_02) {
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = (f[j] > 0.f ? f[j] : 0.f) + 0.2f * (f[j] < 0.f ? f[j] : 0.f);
                    } else if (p.act == ACT_EXP_SIN_11) {
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = (cb + j < 11) ? expf(f[j]) : sinf(f[j]);
                    }
                    if (vec) {
                        if (p.outF) {
                            float * o = p.outF + r * p.ldo + p.coff + cb;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        }
                        if (p.outH) {
                            __half * o = p.outH + r * p.ldoh + p.coffh + cb;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                __half2 h0 = __floats2half2_rn(f[j], f[j + 1]), h1 = __floats2half2_rn(f[j + 2], f[j + 3]);
                                __half2 h2 = __floats2half2_rn(f[j + 4], f[j + 5]), h3 = __floats2half2_rn(f[j + 6], f[j + 7]);
                                uint4 u;
                                u.x = *reinterpret_cast<uint32_t *>(&h0); u.y = *reinterpret_cast<uint32_t *>(&h1);
                                u.z = *reinterpret_cast<uint32_t *>(&h2); u.w = *reinterpret_cast<uint32_t *>(&h3);
                                *reinterpret_cast<uint4 *>(o + j) = u;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int c = cb + j;
                            if (c < p.N) {
                                if (p.outF) p.outF[r * p.ldo + p.coff + c] = f[j];
                                if (p.outH) p.outH[r * p.ldoh + p.coffh + c] = __float2half_rn(f[j]);
                            }
                        }
                    }
                }
                if (p.statsPart) {
                    // fused InstanceNorm statistics: column sums of the stored values over this warp's 32 rows (butterfly
                    // reduce-scatter: 31 shuffles leave lane l with column l), staged per warp for the cross-warp combine below
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) { if (!on || cb + j >= p.N) f[j] = 0.f; sq[j] = f[j] * f[j]; }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; i++) {
                            const float send = up ? f[i] : f[i + off], keep = up ? f[i + off] : f[i];
                            f[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            const float send2 = up ? sq[i] : sq[i + off], keep2 = up ? sq[i + off] : sq[i];
                            sq[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
                        }
                    }
                    red[((warp - 2) * NT + c0 + lane) * 2 + 0] = f[0];
                    red[((warp - 2) * NT + c0 + lane) * 2 + 1] = sq[0];
                }
            }
            if (p.statsPart) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int te = threadIdx.x - 64;
                for (int c = te; c < NT; c += 128) {
                    if (n0 + c < p.N) {
                        const float s0 = (red[(0 * NT + c) * 2] + red[(1 * NT + c) * 2]) + (red[(2 * NT + c) * 2] + red[(3 * NT + c) * 2]);
                        const float s1 = (red[(0 * NT + c) * 2 + 1] + red[(1 * NT + c) * 2 + 1]) + (red[(2 * NT + c) * 2 + 1] + red[(3 * NT + c) * 2 + 1]);
                        float * dst = p.statsPart + (((size_t) b * e.n_mt + mt) * p.N + n0 + c) * 2;
                        dst[0] = s0; dst[1] = s1;
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acb]);
            acc_it++;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * NT) : "memory");
    }
}

// rows [len_b, len_b + halo) of every utterance cleared (ragged batches: the conv must see zeros past the end)
__global__ void zero_tail_rows_kernel(__half * A, int lda, int C, int Lmax, const int * __restrict__ len, int halo) {
    const int b = blockIdx.y;
    const int t = len[b] + blockIdx.x;
    if (blockIdx.x >= halo || t >= Lmax) return;
    __half * row = A + ((size_t) b * Lmax + t) * lda;
    for (int c = threadIdx.x; c < C; c += blockDim.x) row[c] = __float2half_rn(0.f);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_umma_state = 0;   // 0 = unknown, 1 = usable, -1 = disabled
int g_num_sms = 148;

int umma_init() {
    if (g_umma_state) return g_umma_state;
    g_umma_state = -1;
    if (getenv("B2TTS_NO_UMMA")) return g_umma_state;
    void * fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return g_umma_state;
    }
    g_encode = (EncodeTiledFn) fn;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    constexpr int smem256 = NSTAGE * 256 * 128 + 2 * A_BUF_BYTES + 256 + 4 * 256 * 8, smem128 = NSTAGE * 128 * 128 + 2 * A_BUF_BYTES + 256 + 4 * 128 * 8;
    if (cudaFuncSetAttribute(conv_umma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem256) != cudaSuccess ||
        cudaFuncSetAttribute(conv_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128) != cudaSuccess) {
        cudaGetLastError();
        return g_umma_state;
    }
    g_umma_state = 1;
    return g_umma_state;
}

}  // namespace

// returns 0 = launched, 1 = error, 2 = shape not supported here (caller falls back to the mma.sync kernel)
int conv_umma(Ctx * ctx, const ConvGemmParams & p) {
    if (p.stride != 1 || p.CinPad % BKC != 0 || p.lda % 8 != 0 || p.N < 128) return 2;
    const int NT = (p.Npad % 256 == 0 && p.N > 128) ? 256 : (p.Npad % 128 == 0 ? 128 : 0);
    if (!NT) return 2;
    const int RA = round_up(UM + (p.KW - 1) * p.dil, 8);
    if (RA > RA_MAX) return 2;
    if (p.LmaxIn != p.LmaxOut && p.KW > 1) return 2;
    if (umma_init() != 1) return 2;
    if (((uintptr_t) p.A & 15) || ((uintptr_t) p.W & 15)) return 2;

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[3] = {(cuuint64_t) p.CinPad, (cuuint64_t) p.LmaxIn, (cuuint64_t) p.B};
        cuuint64_t strides[2] = {(cuuint64_t) p.lda * 2, (cuuint64_t) p.LmaxIn * p.lda * 2};
        cuuint32_t box[3] = {8, (cuuint32_t) RA, 1};
        cuuint32_t es[3] = {1, 1, 1};
        if (g_encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void *) p.A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 2;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t) p.KW * p.CinPad, (cuuint64_t) p.Npad};
        cuuint64_t strides[1] = {(cuuint64_t) p.KW * p.CinPad * 2};
        cuuint32_t box[2] = {BKC, (cuuint32_t) NT};
        cuuint32_t es[2] = {1, 1};
        if (g_encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *) p.W, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 2;
    }
    if (p.KW > 1 && p.lenIn) {   // ragged batches: clear the halo rows past each utterance's end
        dim3 grid(32, p.B);
        zero_tail_rows_kernel<<<grid, 128, 0, ctx->stream>>>(const_cast<__half *>(p.A), p.lda, p.CinPad, p.LmaxIn, p.lenIn, 32);
        B2_LAUNCH_CHECK(ctx);
    }
    UmmaExtra e;
    e.RA = RA;
    e.n_mt = cdiv(p.LmaxOut, UM);
    e.n_nt = cdiv(p.N, NT);
    e.total_tiles = p.B * e.n_mt * e.n_nt;
    auto al4 = [](const void * q, int ld, int off) { return q == nullptr || ((((uintptr_t) q) & 15) == 0 && ld % 4 == 0 && off % 4 == 0); };
    e.vec4 = al4(p.outF, p.ldo, p.coff) && al4(p.add1, p.ldadd1, 0) && al4(p.add2, p.ldadd2, 0) && al4(p.bias, 4, 0) &&
             (p.outH == nullptr || ((((uintptr_t) p.outH) & 15) == 0 && p.ldoh % 8 == 0 && p.coffh % 8 == 0));
    const int grid = e.total_tiles < g_num_sms ? e.total_tiles : g_num_sms;
    {
        const double rows = (double) (p.validRows ? p.validRows : (int64_t) p.B * p.LmaxOut), cin = (double) (p.CinTrue ? p.CinTrue : p.CinPad);
        ctx->prof_begin(PROF_GEMM, 2.0 * rows * p.N * p.KW * cin, rows * cin * 2.0 + (double) p.N * p.KW * cin * 2.0 + rows * p.N * 4.0);
    }
    if (NT == 256) {
        constexpr int smem = NSTAGE * 256 * 128 + 2 * A_BUF_BYTES + 256 + 4 * 256 * 8;
        conv_umma_kernel<256><<<grid, UMMA_THREADS, smem, ctx->stream>>>(p, e, tmA, tmB);
    } else {
        constexpr int smem = NSTAGE * 128 * 128 + 2 * A_BUF_BYTES + 256 + 4 * 128 * 8;
        conv_umma_kernel<128><<<grid, UMMA_THREADS, smem, ctx->stream>>>(p, e, tmA, tmB);
    }
    ctx->prof_end();
    ctx->umma_launches++;
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b2
