// gemm_umma.cu -- persistent, warp-specialised tcgen05 + TMA implicit-GEMM Conv1d / Linear for sm_100a.
//
// Same contract as conv_gemm_kernel (gemm_conv.cu): out[b][t][co] = epilogue( sum_{k,ci} A[b][t + k*dil - pad][ci] * W[co][k][ci] ),
// fp16 operands, fp32 accumulation -- the reference's ggml_conv_1d / ggml_mul_mat with F16 weights (ggml/src/ggml.c:3870-3894,
// ggml/src/ggml-cpu/ggml-cpu.c:262-267) -- for stride-1 layers with Cin % 64 == 0 (all the heavy ones).
//
// Blackwell mapping (one CTA per SM, persistent over (utterance, time tile, Cout tile) work items; two tile shapes, see below):
//   * warp 0  = TMA producer.  Per 64-channel chunk the activation rows are loaded ONCE with the conv halo
//               (tile rows + (K-1)*dil) as a 128B-swizzled [rows][128 B] chunk; every tap then reads the same shared-memory
//               chunk from a start address shifted by tap*dil rows: no im2col, no reload per tap.
//               Weight tiles [NT][64] (128B-swizzled) stream through a ring, one per (chunk, tap).
//   * warp 1  = MMA issuer: one elected thread issues tcgen05.mma M=128 N=NT K=16; tcgen05.commit releases smem stages and
//               publishes accumulators (double-buffered in TMEM, 2 x 256 columns, so the next tile's MMAs overlap the epilogue).
//   * warps 2-9 = epilogue.  tcgen05.ld gives a thread one accumulator ROW; rows are >= 512 B apart in global memory, so the
//               32x32 chunk is transposed through a per-warp shared-memory tile and everything else -- bias, residual adds,
//               divide, activation, fp32/fp16 stores, fused InstanceNorm column sums -- runs row-contiguous (one warp
//               instruction = 4 rows x 128 B).  Measured: with ~1240 instructions per chunk the 8 epilogue warps (2 per
//               scheduler) were the bottleneck of the big generator layers; this form needs ~4x fewer.
//   * activations outside [0, len_b) must read as zero: TMA zero-fills out-of-range rows, and zero_tail_rows() clears the
//     first rows past each utterance's end for ragged batches.
#include "common.cuh"
#include <cstdio>
#include <cuda.h>
#include <cstdlib>
#include <type_traits>

namespace b2 {
namespace {

constexpr int UM = 128, BKC = 64;
constexpr int RA_MAX = 320;                            // tile rows + (K-1)*dil, rounded up to 16
// Two tile configurations (template <NT, NH>):
//   Cout % 256 == 0 : NT = 256, NH = 1 -> 128 rows x 256 columns per work item, one accumulator (largest MMA, N = 256)
//   otherwise       : NT = 128, NH = 2 -> 256 rows x 128 columns: each 16 KB weight stage feeds two 128-row accumulators
// Either way a weight byte fetched from L2 is used for 128*256 MACs per K element and TMEM holds 2 x 256 columns (double buffer).
__host__ __device__ constexpr int nstage_for(int nt) { return nt == 256 ? 4 : 6; }   // weight-tile ring (4 x 32 KB or 6 x 16 KB)
constexpr int STAGE_PITCH = 36;                        // floats per staged row (32 + 4: conflict-free 16 B accesses both ways)
constexpr int STAGE_WARP_BYTES = 32 * STAGE_PITCH * 4;
constexpr int UMMA_THREADS = 320;                      // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int RED_BYTES = 8 * 128 * 8;                 // [slabs][NT][2] floats: 8 x 128 or 4 x 256 columns
__host__ __device__ constexpr int fixed_bytes(int nt) { return (nstage_for(nt) * nt * 128 + 256 + RED_BYTES + 8 * STAGE_WARP_BYTES + 1023) / 1024 * 1024; }
constexpr int UMMA_SMEM_TOTAL = 232448;                // 227 KB: all the opt-in shared memory of an SM
constexpr int MAX_ABUF = 4;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// exactly one lane of a fully converged warp: the compiler then knows the guarded code runs on a single lane and keeps the
// tcgen05 operands in uniform registers (an `if (lane == 0)` guard made it wrap every MMA in an R2UR + ELECT + BRA.U.ANY
// "uniformisation" loop, ~150 cycles per instruction: the issuing thread, not the tensor core, was the bottleneck)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
    if (elect_one())
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
    if (elect_one())
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void * dst, const CUtensorMap * tm, uint64_t * bar, int c0, int c1, int c2) {
    if (elect_one())
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (elect_one())
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t * bar) {
    if (elect_one())
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
// shared-memory matrix descriptor (K-major operand).  layout_type: 0 = no swizzle, 2 = 128B swizzle; version field = 1 on sm_100.
// (Row-shifted 128B-swizzled descriptors via base_offset were tried for the activation operand and give wrong tap shifts
//  for dil = 3 on sm_100a, so the activation chunk uses the un-swizzled core-matrix layout, where a row shift is start + 16 B.)
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
    int RA;            // activation rows staged per chunk: TM + (KW-1)*dil rounded up to 16 (loaded as two TMA boxes of RA/2 rows)
    int n_abuf;        // activation chunk buffers in the ring
    int n_mt;          // TM-row tiles per utterance
    int n_nt;          // NT-column Cout tiles
    int total_tiles;
    int vec4;          // epilogue may use 16-byte accesses
    int dbg;           // B2TTS_UMMA_DBG ablation bits (timing experiments only; results are wrong when set)
};

template <int NT, int NH>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_umma_kernel(const ConvGemmParams p, const UmmaExtra e, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    constexpr int TM = UM * NH, NSTAGE = nstage_for(NT), B_STAGE_BYTES = NT * 128, FIXED_BYTES = fixed_bytes(NT);
    constexpr int NSLAB = 4 * NH;                 // row slabs of 32 rows that own a distinct set of tile rows
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char * sB = smem;                                        // [NSTAGE][NT][128 B], 128B-swizzled by TMA
    uint64_t * bars = reinterpret_cast<uint64_t *>(smem + NSTAGE * B_STAGE_BYTES);
    uint64_t * a_full = bars, * a_empty = bars + MAX_ABUF, * b_full = bars + 2 * MAX_ABUF, * b_empty = b_full + NSTAGE;
    uint64_t * acc_full = b_empty + NSTAGE, * acc_empty = acc_full + 2;
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    float * red = reinterpret_cast<float *>(smem + NSTAGE * B_STAGE_BYTES + 256);       // [8 slabs][NT][2]
    float * stage_all = red + RED_BYTES / 4;                                               // [8 epilogue warps][32][STAGE_PITCH]
    unsigned char * sA = smem + FIXED_BYTES;                                            // [n_abuf][RA][128 B], 128B-swizzled by TMA
    const int a_buf_bytes = e.RA * 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_cc = p.CinPad / BKC;

    if (threadIdx.x == 0) {
        for (int i = 0; i < MAX_ABUF; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NSTAGE; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 2 (double buffer) x NH accumulators x NT fp32 columns = all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer (whole warp runs the loop; one elected lane issues)
        {
            uint32_t a_it = 0, b_it = 0;
            const int RAh = e.RA / 2;
            for (int tile = blockIdx.x; tile < e.total_tiles; tile += gridDim.x) {
                const int nt = tile % e.n_nt, rest = tile / e.n_nt, mt = rest % e.n_mt, b = rest / e.n_mt;
                const int t0 = mt * TM, n0 = nt * NT;
                for (int cc = 0; cc < n_cc; cc++) {
                    const int ab = a_it % e.n_abuf;
                    mbar_wait(&a_empty[ab], ((a_it / e.n_abuf) & 1) ^ 1);
                    mbar_expect_tx(&a_full[ab], (e.dbg & 1) ? 0u : (uint32_t) e.RA * 128u);
                    if (!(e.dbg & 1)) {   // one 128B-swizzled [RA rows][64 channels] chunk, as two boxes of RA/2 rows
                        unsigned char * dst = sA + ab * a_buf_bytes;
                        tma_load_3d(dst, &tmA, &a_full[ab], cc * BKC, t0 - p.pad, b);
                        tma_load_3d(dst + RAh * 128, &tmA, &a_full[ab], cc * BKC, t0 - p.pad + RAh, b);
                    }
                    a_it++;
                    for (int tap = 0; tap < p.KW; tap++) {
                        const int s = b_it % NSTAGE;
                        mbar_wait(&b_empty[s], ((b_it / NSTAGE) & 1) ^ 1);
                        mbar_expect_tx(&b_full[s], (e.dbg & 2) ? 0u : (uint32_t) B_STAGE_BYTES);
                        if (!(e.dbg & 2)) tma_load_2d(sB + s * B_STAGE_BYTES, &tmB, &b_full[s], tap * p.CinPad + cc * BKC, n0);
                        b_it++;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        const uint32_t idesc = (1u << 4) | ((uint32_t) (NT >> 3) << 17) | ((uint32_t) (UM >> 4) << 24);   // F32 accum, F16 x F16, K-major A and B
        // descriptor templates (everything but the 14-bit start address, which is added per MMA in 16-byte units)
        const uint64_t bdesc_t = ((uint64_t) (16u >> 4) << 16) | ((uint64_t) (1024u >> 4) << 32) | ((uint64_t) 1 << 46) | ((uint64_t) 2 << 61);   // 128B swizzle
        // A uses the same canonical layout (128 B rows, 8-row swizzle groups 1024 B apart).  A conv tap is the SAME staged chunk read
        // from a start address shifted by tap*dil whole rows: the swizzle XOR is a function of the absolute shared-memory address
        // (chunk buffers are 1024 B aligned), so a row-shifted start needs no base offset -- verified on hardware for dil 1, 3, 5.
        const uint64_t adesc_t = bdesc_t;
        constexpr uint32_t a_row16 = 8u, a_k16 = 2u;   // descriptor address steps (16 B units): one row, one K=16 slice
        const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
        const uint32_t abuf16 = (uint32_t) a_buf_bytes >> 4;
        uint32_t a_it = 0, b_it = 0, acc_it = 0;
        for (int tile = blockIdx.x; tile < e.total_tiles; tile += gridDim.x) {
            const int acb = acc_it & 1;
            mbar_wait(&acc_empty[acb], ((acc_it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t) (acb * 256);
            uint32_t accumulate = 0;
            for (int cc = 0; cc < n_cc; cc++) {
                const int ab = a_it % e.n_abuf;
                mbar_wait(&a_full[ab], (a_it / e.n_abuf) & 1);
                const uint32_t a16 = sA16 + (uint32_t) ab * abuf16;
                for (int tap = 0; tap < p.KW; tap++) {
                    const int s = b_it % NSTAGE;
                    mbar_wait(&b_full[s], (b_it / NSTAGE) & 1);
                    tc_fence_after();
                    {
                        const uint32_t b16 = sB16 + (uint32_t) s * (B_STAGE_BYTES >> 4);
                        const uint32_t at16 = a16 + (uint32_t) (tap * p.dil) * a_row16;
#pragma unroll
                        for (int j = 0; j < BKC / 16; j++) {
                            const uint64_t bd = bdesc_t | (uint64_t) (b16 + 2u * j);
#pragma unroll
                            for (int h = 0; h < NH; h++) {   // the same weight stage feeds every 128-row accumulator of the tile
                                const uint64_t ad = adesc_t | (uint64_t) (at16 + (uint32_t) j * a_k16 + (uint32_t) (h * UM) * a_row16);
                                if (!(e.dbg & 4)) umma_f16(d_tmem + (uint32_t) (h * NT), ad, bd, idesc, accumulate);
                            }
                            accumulate = 1;
                        }
                        umma_commit(&b_empty[s]);
                    }
                    __syncwarp();
                    b_it++;
                }
                umma_commit(&a_empty[ab]);
                __syncwarp();
                a_it++;
            }
            umma_commit(&acc_full[acb]);
            __syncwarp();
            acc_it++;
        }
    } else {
        // ================================================================= epilogue: warp -> (row half mh, TMEM lane quarter q)
        // A thread owns one accumulator ROW, but rows are >= 512 B apart in global memory, so global accesses are staged through
        // a per-warp shared-memory tile and issued row-contiguous (one instruction = 4 rows x 128 B).  The residual chunk for the
        // next iteration is loaded one chunk ahead.
        const int q = warp & 3, g = (warp - 2) >> 2;
        const int mh = NH == 2 ? g : 0;                       // NH == 2: the two warp groups take the two row halves, all columns
        const int c_lo = NH == 2 ? 0 : g * (NT / 2), c_hi = NH == 2 ? NT : (g + 1) * (NT / 2);   // NH == 1: they split the columns
        const int slab = mh * 4 + q;
        const int rloc = mh * UM + q * 32;                    // first row (within the tile) of this warp's 32-row slab
        float * stg = stage_all + (warp - 2) * (32 * STAGE_PITCH);
        const int cr = lane >> 3, cc4 = (lane & 7) * 4;      // row-contiguous pattern: step i covers rows 4i + cr, columns cc4..cc4+3
        const int emode = (p.div == 0.f && p.act == ACT_NONE) ? 0 : (p.act == ACT_NONE ? 1 : ((p.div == 0.f && p.act == ACT_GELU_F16LUT) ? 2 : 3));
        const bool efast = emode <= 1 && p.bias != nullptr && p.outF != nullptr && p.outH == nullptr && !(e.dbg & 16);
        const bool do_f = p.outF != nullptr, do_h = p.outH != nullptr, has1 = p.add1 != nullptr, has2 = p.add2 != nullptr;
        uint32_t acc_it = 0;
        for (int tile = blockIdx.x; tile < e.total_tiles; tile += gridDim.x) {
            const int nt = tile % e.n_nt, rest = tile / e.n_nt, mt = rest % e.n_mt, b = rest / e.n_mt;
            const int n0 = nt * NT;
            const int tq = mt * TM + rloc;
            const int lo = min(p.lenOut ? p.lenOut[b] : p.LmaxOut, p.LmaxOut);
            const int nrows = lo - tq - cr;                          // step i is a valid output row iff 4 * i < nrows
            const size_t rq = (size_t) b * p.LmaxOut + tq;
            // per-tile row pointers of the row-contiguous pattern (advanced by 4 rows per step, by 32 columns per chunk)
            const float * a1p = has1 ? p.add1 + (rq + cr) * p.ldadd1 + n0 + cc4 : nullptr;
            const float * a2p = has2 ? p.add2 + (rq + cr) * p.ldadd2 + n0 + cc4 : nullptr;
            float *       ofp = do_f ? p.outF + (rq + cr) * p.ldo + p.coff + n0 + cc4 : nullptr;
            __half *      ohp = do_h ? p.outH + (rq + cr) * p.ldoh + p.coffh + n0 + cc4 : nullptr;
            const size_t s1 = 4 * (size_t) p.ldadd1, s2 = 4 * (size_t) p.ldadd2, sf = 4 * (size_t) p.ldo, sh = 4 * (size_t) p.ldoh;
            const int acb = acc_it & 1;
            bool waited = false;
            const uint32_t taddr = tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (acb * 256 + mh * NT);
#pragma unroll 1
            for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                const int cb = n0 + c0;
                if ((e.dbg & 8) || cb >= p.N) break;
                if (e.vec4 && cb + 32 <= p.N) {
                    // ---- residual rows of this chunk: issued first, they do not depend on the accumulator
                    float4 r1[8], r2[8];
                    if (has1) {
#pragma unroll
                        for (int i = 0; i < 8; i++) r1[i] = (4 * i < nrows) ? __ldg(reinterpret_cast<const float4 *>(a1p + i * s1 + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (has2) {
#pragma unroll
                        for (int i = 0; i < 8; i++) r2[i] = (4 * i < nrows) ? __ldg(reinterpret_cast<const float4 *>(a2p + i * s2 + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4 *>(p.bias + cb + cc4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!waited) { mbar_wait(&acc_full[acb], (acc_it >> 1) & 1); tc_fence_after(); waited = true; }
                    // ---- accumulator rows (one per thread) -> shared tile, then everything below works row-contiguous:
                    //      one warp instruction = 4 rows x 128 B
                    {
                        uint32_t v[32];
                        if (!(e.dbg & 32)) tmem_ld32(taddr + (uint32_t) c0, v);
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; j++) *reinterpret_cast<uint4 *>(stg + lane * STAGE_PITCH + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        __syncwarp();
                    }
                    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
                    auto steps = [&](auto mode_tag) {
                    // compile-time epilogue flavour: 0 plain (no division, no activation: the resblock convs), 1 division only
                    // (AdaIN decoder blocks), 2 GELU only (ALBERT FFN), 3 anything
                    constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float4 x4 = *reinterpret_cast<const float4 *>(stg + (4 * i + cr) * STAGE_PITCH + cc4);
                        float x[4] = {x4.x, x4.y, x4.z, x4.w};
                        if (p.bias) { x[0] += b4.x; x[1] += b4.y; x[2] += b4.z; x[3] += b4.w; }
                        if (has1) { x[0] = r1[i].x + x[0]; x[1] = r1[i].y + x[1]; x[2] = r1[i].z + x[2]; x[3] = r1[i].w + x[3]; }
                        if (has2) { x[0] = r2[i].x + x[0]; x[1] = r2[i].y + x[1]; x[2] = r2[i].z + x[2]; x[3] = r2[i].w + x[3]; }
                        if constexpr (MODE == 1) {
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = __fdiv_rn(x[k], p.div);
                        } else if constexpr (MODE == 2) {
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = gelu_f16lut_u(x[k]);
                        } else if constexpr (MODE == 3) {
                            if (p.div != 0.f) {
#pragma unroll
                                for (int k = 0; k < 4; k++) x[k] = __fdiv_rn(x[k], p.div);
                            }
                            if (p.act == ACT_GELU_F16LUT) {
#pragma unroll
                                for (int k = 0; k < 4; k++) x[k] = gelu_f16lut_u(x[k]);
                            } else if (p.act == ACT_LRELU_02) {
#pragma unroll
                                for (int k = 0; k < 4; k++) x[k] = (x[k] > 0.f ? x[k] : 0.f) + 0.2f * (x[k] < 0.f ? x[k] : 0.f);
                            } else if (p.act == ACT_EXP_SIN_11) {
#pragma unroll
                                for (int k = 0; k < 4; k++) x[k] = (cb + cc4 + k < 11) ? expf(x[k]) : sinf(x[k]);
                            } else if (p.act == ACT_TANH) {
#pragma unroll
                                for (int k = 0; k < 4; k++) x[k] = tanhf(x[k]);
                            }
                        }
                        if (4 * i < nrows) {
                            if (do_f && !(e.dbg & 16)) *reinterpret_cast<float4 *>(ofp + i * sf + c0) = make_float4(x[0], x[1], x[2], x[3]);
                            if (do_h) {
                                const __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
                                uint2 u;
                                u.x = *reinterpret_cast<const uint32_t *>(&h0); u.y = *reinterpret_cast<const uint32_t *>(&h1);
                                *reinterpret_cast<uint2 *>(ohp + i * sh + c0) = u;
                            }
#pragma unroll
                            for (int k = 0; k < 4; k++) { cs[k] += x[k]; cq[k] += x[k] * x[k]; }
                        }
                    }
                    };
                    // branch-free body for the generator's resblock convs (bias, fp32 result, <= 1 residual, no fp16 copy): with every
                    // launch-uniform flag a template constant the 8 unrolled steps are ONE basic block, so their independent
                    // LDS -> FADD -> STG chains interleave (each runtime-uniform `if` used to end a block and serialise them)
                    auto fast = [&](auto has1_tag, auto has2_tag, auto div_tag, auto stats_tag) {
                        constexpr bool HAS1 = decltype(has1_tag)::value, HAS2 = decltype(has2_tag)::value, DIV = decltype(div_tag)::value, STATS = decltype(stats_tag)::value;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const float4 x4 = *reinterpret_cast<const float4 *>(stg + (4 * i + cr) * STAGE_PITCH + cc4);
                            float x[4] = {x4.x + b4.x, x4.y + b4.y, x4.z + b4.z, x4.w + b4.w};
                            if constexpr (HAS1) { x[0] = r1[i].x + x[0]; x[1] = r1[i].y + x[1]; x[2] = r1[i].z + x[2]; x[3] = r1[i].w + x[3]; }
                            if constexpr (HAS2) { x[0] = r2[i].x + x[0]; x[1] = r2[i].y + x[1]; x[2] = r2[i].z + x[2]; x[3] = r2[i].w + x[3]; }
                            if constexpr (DIV) {
#pragma unroll
                                for (int k = 0; k < 4; k++) x[k] = __fdiv_rn(x[k], p.div);
                            }
                            const bool rv = 4 * i < nrows;
                            if (rv) *reinterpret_cast<float4 *>(ofp + i * sf + c0) = make_float4(x[0], x[1], x[2], x[3]);
                            if constexpr (STATS) {
#pragma unroll
                                for (int k = 0; k < 4; k++) { const float xm = rv ? x[k] : 0.f; cs[k] += xm; cq[k] = fmaf(xm, xm, cq[k]); }
                            }
                        }
                    };
                    if (efast) {
                        using T = std::true_type; using F = std::false_type;
                        auto d4 = [&](auto a, auto b, auto c) { if (p.statsPart) fast(a, b, c, T{}); else fast(a, b, c, F{}); };
                        auto d3 = [&](auto a, auto b) { if (emode == 1) d4(a, b, T{}); else d4(a, b, F{}); };
                        auto d2 = [&](auto a) { if (has2) d3(a, T{}); else d3(a, F{}); };
                        if (has1) d2(T{}); else d2(F{});
                    }
                    else if (emode == 0) steps(std::integral_constant<int, 0>{});
                    else if (emode == 1) steps(std::integral_constant<int, 1>{});
                    else if (emode == 2) steps(std::integral_constant<int, 2>{});
                    else steps(std::integral_constant<int, 3>{});
                    if (p.statsPart && !(e.dbg & 64)) {
                        // fused InstanceNorm statistics: per-thread column sums over its 8 rows, then across the 4 row groups of the warp
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);  cq[k] += __shfl_xor_sync(0xffffffffu, cq[k], 8);
                            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16); cq[k] += __shfl_xor_sync(0xffffffffu, cq[k], 16);
                        }
                        if (lane < 8) {
                            float4 * dst = reinterpret_cast<float4 *>(red + (slab * NT + c0 + cc4) * 2);
                            dst[0] = make_float4(cs[0], cq[0], cs[1], cq[1]);
                            dst[1] = make_float4(cs[2], cq[2], cs[3], cq[3]);
                        }
                    }
                } else {
                    // ---- generic path (unaligned rows or a ragged last column chunk): lane = column, loop over the 32 rows
                    if (!waited) { mbar_wait(&acc_full[acb], (acc_it >> 1) & 1); tc_fence_after(); waited = true; }
                    {
                        uint32_t v[32];
                        tmem_ld32(taddr + (uint32_t) c0, v);
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; j++) *reinterpret_cast<uint4 *>(stg + lane * STAGE_PITCH + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        __syncwarp();
                    }
                    const int c = cb + lane;
                    const bool cok = c < p.N;
                    const float bs = (p.bias && cok) ? p.bias[c] : 0.f;
                    float s0 = 0.f, q0 = 0.f;
                    const int rmax = min(32, lo - tq);
#pragma unroll 4
                    for (int r = 0; r < rmax; r++) {
                        float x = stg[r * STAGE_PITCH + lane];
                        if (cok) {
                            const size_t row = rq + r;
                            x = x + bs;
                            if (has1) x = p.add1[row * p.ldadd1 + c] + x;
                            if (has2) x = p.add2[row * p.ldadd2 + c] + x;
                            if (p.div != 0.f) x = __fdiv_rn(x, p.div);
                            if (p.act == ACT_GELU_F16LUT) x = gelu_f16lut_u(x);
                            else if (p.act == ACT_LRELU_02) x = (x > 0.f ? x : 0.f) + 0.2f * (x < 0.f ? x : 0.f);
                            else if (p.act == ACT_EXP_SIN_11) x = (c < 11) ? expf(x) : sinf(x);
                            else if (p.act == ACT_TANH) x = tanhf(x);
                            if (do_f) p.outF[row * p.ldo + p.coff + c] = x;
                            if (do_h) p.outH[row * p.ldoh + p.coffh + c] = __float2half_rn(x);
                            s0 += x; q0 += x * x;
                        }
                    }
                    if (p.statsPart) *reinterpret_cast<float2 *>(red + (slab * NT + c0 + lane) * 2) = make_float2(s0, q0);
                }
            }
            if (!waited) { mbar_wait(&acc_full[acb], (acc_it >> 1) & 1); tc_fence_after(); }
            if (p.statsPart) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int c = threadIdx.x - 64;
                if (c < NT && n0 + c < p.N) {
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int sl = 0; sl < NSLAB; sl++) { const float2 a = *reinterpret_cast<const float2 *>(red + (sl * NT + c) * 2); s0 += a.x; s1 += a.y; }
                    *reinterpret_cast<float2 *>(p.statsPart + (((size_t) b * e.n_mt + mt) * p.N + n0 + c) * 2) = make_float2(s0, s1);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
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
    if (cudaFuncSetAttribute(conv_umma_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UMMA_SMEM_TOTAL) != cudaSuccess ||
        cudaFuncSetAttribute(conv_umma_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UMMA_SMEM_TOTAL) != cudaSuccess ||
        cudaFuncSetAttribute(conv_umma_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UMMA_SMEM_TOTAL) != cudaSuccess) {
        cudaGetLastError();
        return g_umma_state;
    }
    g_umma_state = 1;
    return g_umma_state;
}

}  // namespace

// rows per work item the kernel will use for these params (layout of statsPart); 0 if the shape is not supported
// Cout >= 128 (weights padded to 128 rows), or a narrow head (Cout < 128, weights padded to 64 rows: the TMA box reads the
// missing weight rows as zeros) when there are enough output rows to fill the machine
static bool umma_shape_ok(const ConvGemmParams & p) {
    if (p.stride != 1 || p.CinPad % BKC != 0) return false;
    if (p.N >= 128 && p.Npad % 128 == 0) return true;
    return p.N < 128 && p.Npad % 64 == 0 && (int64_t) p.B * p.LmaxOut >= 32768;
}
// tile shape: Cout % 256 == 0 -> 128 x 256 (or 128 x 128 when that would leave more than half of the SMs without a tile: small
// pointwise layers); otherwise 256 x 128
static void umma_config(const ConvGemmParams & p, int & NT, int & NH) {
    if (p.Npad % 256 == 0) {
        const int64_t tiles256 = (int64_t) p.B * cdiv(p.LmaxOut, 128) * cdiv(p.N, 256);
        NT = (2 * tiles256 <= g_num_sms && !p.statsPart) ? 128 : 256; NH = 1;
    } else { NT = 128; NH = 2; }
}
int conv_umma_tile_m(const ConvGemmParams & p) {
    if (!umma_shape_ok(p)) return 0;
    return (p.Npad % 256 == 0) ? 128 : 256;
}

// returns 0 = launched, 1 = error, 2 = shape not supported here (caller falls back to the mma.sync kernel)
int conv_umma(Ctx * ctx, const ConvGemmParams & p_in) {
    ConvGemmParams p = p_in;
    if (!umma_shape_ok(p) || p.lda % 8 != 0) return 2;
    if (p.KW == 1 && p.LmaxIn == p.LmaxOut && p.B > 1 && !p.statsPart) {
        // pointwise layers (Linear): the batch is one flat row range [0, B * Lmax).  Rows past an utterance's length are computed
        // and stored too -- they are padding no consumer reads (kernels.cuh) -- which removes the per-utterance partial tiles.
        const int tm = conv_umma_tile_m(p);
        if (cdiv((int64_t) p.B * p.LmaxOut, tm) < p.B * cdiv(p.LmaxOut, tm)) {
            p.LmaxIn = p.LmaxOut = p.B * p.LmaxOut; p.B = 1; p.lenIn = p.lenOut = nullptr;
        }
    }
    int NT, NH;
    umma_config(p, NT, NH);
    const int TM = UM * NH;
    const int FIXED_BYTES = fixed_bytes(NT);
    const int RA = round_up(TM + (p.KW - 1) * p.dil, 16);
    if (RA > RA_MAX) return 2;
    if (umma_init() != 1) return 2;
    if (((uintptr_t) p.A & 15) || ((uintptr_t) p.W & 15)) return 2;
    UmmaExtra e;
    e.RA = RA;
    e.n_abuf = (UMMA_SMEM_TOTAL - FIXED_BYTES - 1024) / (RA * 128);
    if (e.n_abuf > MAX_ABUF) e.n_abuf = MAX_ABUF;
    if (e.n_abuf < 2) return 2;

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[3] = {(cuuint64_t) p.CinPad, (cuuint64_t) p.LmaxIn, (cuuint64_t) p.B};
        cuuint64_t strides[2] = {(cuuint64_t) p.lda * 2, (cuuint64_t) p.LmaxIn * p.lda * 2};
        cuuint32_t box[3] = {BKC, (cuuint32_t) (RA / 2), 1};
        cuuint32_t es[3] = {1, 1, 1};
        if (g_encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void *) p.A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
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
    if (p.KW > 1 && p.lenIn && !p.tailClean) {   // ragged batches: clear the halo rows past each utterance's end
        dim3 grid(32, p.B);
        zero_tail_rows_kernel<<<grid, 128, 0, ctx->stream>>>(const_cast<__half *>(p.A), p.lda, p.CinPad, p.LmaxIn, p.lenIn, 32);
        B2_LAUNCH_CHECK(ctx);
    }
    { static const int dbg_env = getenv("B2TTS_UMMA_DBG") ? atoi(getenv("B2TTS_UMMA_DBG")) : 0; static const int dbg_lo = getenv("B2TTS_UMMA_DBG_LO") ? atoi(getenv("B2TTS_UMMA_DBG_LO")) : 4096; static const int dbg_hi = getenv("B2TTS_UMMA_DBG_HI") ? atoi(getenv("B2TTS_UMMA_DBG_HI")) : (1 << 30); e.dbg = (p.LmaxOut >= dbg_lo && p.LmaxOut < dbg_hi) ? dbg_env : 0; }
    e.n_mt = cdiv(p.LmaxOut, TM);
    e.n_nt = cdiv(p.N, NT);
    e.total_tiles = p.B * e.n_mt * e.n_nt;
    auto al4 = [](const void * q, int ld, int off) { return q == nullptr || ((((uintptr_t) q) & 15) == 0 && ld % 4 == 0 && off % 4 == 0); };
    e.vec4 = al4(p.outF, p.ldo, p.coff) && al4(p.add1, p.ldadd1, 0) && al4(p.add2, p.ldadd2, 0) && al4(p.bias, 4, 0) &&
             (p.outH == nullptr || ((((uintptr_t) p.outH) & 15) == 0 && p.ldoh % 8 == 0 && p.coffh % 8 == 0));
    const int grid = e.total_tiles < g_num_sms ? e.total_tiles : g_num_sms;
    {
        const double rows = (double) (p.validRows ? p.validRows : (int64_t) p.B * p.LmaxOut), cin = (double) (p.CinTrue ? p.CinTrue : p.CinPad);
        snprintf(ctx->tag, sizeof(ctx->tag), "%s N%d K%d C%d L%d B%d d%d", "umma", p.N, p.KW, p.CinPad, p.LmaxOut, p.B, p.dil);
        ctx->prof_begin(PROF_GEMM, 2.0 * rows * p.N * p.KW * cin, rows * cin * 2.0 + (double) p.N * p.KW * cin * 2.0 + rows * p.N * 4.0);
    }
    if (NT == 256)    conv_umma_kernel<256, 1><<<grid, UMMA_THREADS, UMMA_SMEM_TOTAL, ctx->stream>>>(p, e, tmA, tmB);
    else if (NH == 1) conv_umma_kernel<128, 1><<<grid, UMMA_THREADS, UMMA_SMEM_TOTAL, ctx->stream>>>(p, e, tmA, tmB);
    else              conv_umma_kernel<128, 2><<<grid, UMMA_THREADS, UMMA_SMEM_TOTAL, ctx->stream>>>(p, e, tmA, tmB);
    ctx->prof_end();
    ctx->umma_launches++;
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b2
