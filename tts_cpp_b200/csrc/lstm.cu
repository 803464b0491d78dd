// lstm.cu -- persistent thread-block-cluster bi-LSTM for sm_100a (hidden size 256).
//
// Replaces build_lstm / build_lstm_run (reference src/models/kokoro/model.cpp:35-86), which unrolls every
// time step into 26 GGML nodes per direction (each followed by a thread-pool barrier).  Here one cluster of
// 8 CTAs per (direction, batch tile of 16 utterances) runs the whole sequence:
//   * the hidden-side weights W_hh (fp16, 1024x256) stay resident in shared memory, 128 gate rows per CTA
//     (4 gates x 32 hidden units);
//   * every step is one tensor-core contraction  G[128 x 16] = W_hh_slice[128 x 256] . h^T[256 x 16]
//     (h re-rounded to fp16 exactly like ggml does for an F16 weight, ggml-cpu.c:262-267);
//   * gate math (sigmoid/tanh, c = f*c + i*g, h = o*tanh(c), model.cpp:63-76) happens in registers, the cell
//     state never leaves them;
//   * the new h slice is broadcast to the 8 CTAs' shared memory through DSMEM as st.async stores that complete a transaction count
//     on each destination's mbarrier: a CTA waits only for its own buffer to fill, there is no cluster-wide barrier per step.
// The input-side projections W_ih x + b_ih for all steps come from conv_gemm (they are one big GEMM), laid out
// [b][t][dir][unit][gate] so a thread fetches its 4 gate pre-activations with one 16-byte load.
// Ragged batches: utterance b is active for len[b] steps; the reverse direction walks t = len[b]-1-s.
#include "common.cuh"
#include <cstdlib>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace b2 {
namespace {

// utterances per cluster: 16 (not 32) -- the step is a latency chain (HMMA accumulator chain, MUFU-heavy gate math, cluster barrier), so
// halving the per-thread work per step and running twice as many clusters shortens it; one n8 MMA tile per warp
constexpr int H = 256, CL = 8, UNITS = 32, NBT = 16, NI = NBT / 16;  // hidden, cluster size, units per CTA, utterances per cluster, n8 tiles per warp
constexpr int LDW = 264;                              // smem row stride (halves): 528 B -> conflict-free ldmatrix
constexpr int SW_BYTES = 128 * LDW * 2;
constexpr int SH_BYTES = 2 * NBT * LDW * 2;
constexpr int ST_BYTES = NBT * UNITS * 2;
constexpr int LSTM_SMEM = SW_BYTES + SH_BYTES + ST_BYTES + 16;   // + two mbarriers (the asynchronous h exchange)
constexpr unsigned STEP_BYTES = CL * NBT * UNITS * 2;         // bytes a CTA receives per step: a [NBT][UNITS] fp16 slice from each of the CL CTAs

__device__ __forceinline__ void ldsm_x4(unsigned & r0, unsigned & r1, unsigned & r2, unsigned & r3, const void * p) {
    unsigned s = (unsigned) __cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void ldsm_x2(unsigned & r0, unsigned & r1, const void * p) {
    unsigned s = (unsigned) __cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(s));
}
__device__ __forceinline__ void mma16816(float * c, const unsigned * a, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }  // ggml_vec_sigmoid_f32

template <bool ASYNC>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256) bilstm_kernel(const LstmParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half * sW  = reinterpret_cast<__half *>(smem_raw);                       // [128][LDW]
    __half * sH  = reinterpret_cast<__half *>(smem_raw + SW_BYTES);            // [2][NBT][LDW]
    __half * sSt = reinterpret_cast<__half *>(smem_raw + SW_BYTES + SH_BYTES); // [NBT][UNITS]
    unsigned long long * hbar = reinterpret_cast<unsigned long long *>(smem_raw + SW_BYTES + SH_BYTES + ST_BYTES);   // [2]: one per h buffer

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int) cluster.block_rank();
    const int dir  = blockIdx.y;
    const int b0   = blockIdx.z * NBT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ug = warp & 3, nh = warp >> 2;

    // ---- resident weights: local row lr = ug*32 + tile*16 + r ; gate = tile*2 + (r>=8) ; unit = ug*8 + (r&7)
    const __half * Wd = p.whh + (size_t) dir * 4 * H * H;
    for (int idx = tid; idx < 128 * (H / 8); idx += 256) {
        const int lr = idx / (H / 8), ch = idx % (H / 8);
        const int g4 = lr >> 5, tile = (lr >> 4) & 1, r = lr & 15;
        const int gate = tile * 2 + (r >> 3);
        const int unit = rank * UNITS + g4 * 8 + (r & 7);
        const int4 v = *reinterpret_cast<const int4 *>(Wd + (size_t) (gate * H + unit) * H + ch * 8);
        *reinterpret_cast<int4 *>(sW + lr * LDW + ch * 8) = v;
    }
    for (int idx = tid; idx < 2 * NBT * LDW / 2; idx += 256) reinterpret_cast<unsigned *>(sH)[idx] = 0u;

    // ---- per-thread ownership: unit ul (local), utterances u[ni][e]
    const int ul = ug * 8 + (lane >> 2);
    const int unit_g = rank * UNITS + ul;
    int   ub[NI][2], ulen[NI][2];
    float cst[NI][2];
    int nsteps = 0;
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int u = nh * (8 * NI) + ni * 8 + (lane & 3) * 2 + e;
            ub[ni][e]   = b0 + u;
            ulen[ni][e] = (ub[ni][e] < p.B) ? p.len[ub[ni][e]] : 0;
            cst[ni][e]  = 0.f;
        }
    for (int u = 0; u < NBT; u++)
        if (b0 + u < p.B) nsteps = max(nsteps, p.len[b0 + u]);
    float bh[4];
#pragma unroll
    for (int g = 0; g < 4; g++) bh[g] = p.bhh[dir * 4 * H + g * H + unit_g];

    // ASYNC: the step's h slices travel as st.async (remote shared-memory stores that complete a transaction count on the DESTINATION
    // CTA's mbarrier); a CTA just waits for its own barrier to have received all CL slices.  No cluster-wide barrier and no release
    // fence per step: ncu showed 22 % of the stall samples of the barrier version on the ERRBAR of barrier.cluster.arrive.release.
    // Double buffering is enough: a CTA sends D(s+2) only after it received everybody's D(s+1), which each sender produced after
    // it had finished reading D(s) -- so nobody overwrites a buffer (or runs a barrier phase) that is still in use.
    const unsigned hbar_s = (unsigned) __cvta_generic_to_shared(hbar);
    if (ASYNC && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hbar_s));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hbar_s + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // arm both buffers: D(0) lands in buffer 1, D(1) in buffer 0
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hbar_s), "r"(STEP_BYTES) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hbar_s + 8), "r"(STEP_BYTES) : "memory");
    }
    unsigned hphase0 = 0u, hphase1 = 0u;
    __syncthreads();
    cluster.sync();

    // input-side pre-activations of a step (one float4 = gates i,f,g,o of (b, t, dir, unit)); loaded one step ahead, while the
    // cluster barrier of the previous step is in flight, so their L2 latency is off the recurrence's critical path
    float4 xp[NI][2];
    int    tt[NI][2];
    auto load_xp = [&](int s) {
#pragma unroll
        for (int ni = 0; ni < NI; ni++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const bool act = s < ulen[ni][e];
                tt[ni][e]      = dir == 0 ? s : ulen[ni][e] - 1 - s;
                if (act) {
                    const size_t row = (size_t) ub[ni][e] * p.Lmax + tt[ni][e];
                    xp[ni][e] = *reinterpret_cast<const float4 *>(p.xp + ((row * 2 + dir) * H + unit_g) * 4);
                } else {
                    xp[ni][e] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
    };
    load_xp(0);

    for (int s = 0; s < nsteps; s++) {
        const __half * hcur = sH + (size_t) (s & 1) * NBT * LDW;
        __half *       hnxt = sH + (size_t) ((s + 1) & 1) * NBT * LDW;

        // 2. G = W_hh_slice . h^T on tensor cores: 2 m16 tiles (i|f , g|o) x 2 n8 tiles per warp
        float acc[2][NI][4];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < NI; b++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[a][b][e] = 0.f;
#pragma unroll 4
        for (int ks = 0; ks < H / 16; ks++) {
            unsigned af[2][4], bf[4];
#pragma unroll
            for (int tl = 0; tl < 2; tl++) {
                const int r = ug * 32 + tl * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm_x4(af[tl][0], af[tl][1], af[tl][2], af[tl][3], sW + r * LDW + ks * 16 + (lane >> 4) * 8);
            }
            if constexpr (NI == 2) {
                const int r = nh * 16 + (lane & 7) + (lane >> 4) * 8;
                ldsm_x4(bf[0], bf[1], bf[2], bf[3], hcur + r * LDW + ks * 16 + ((lane >> 3) & 1) * 8);
            } else {
                const int r = nh * 8 + (lane & 7);
                ldsm_x2(bf[0], bf[1], hcur + r * LDW + ks * 16 + ((lane >> 3) & 1) * 8);
                bf[2] = bf[3] = 0u;
            }
#pragma unroll
            for (int tl = 0; tl < 2; tl++) {
                mma16816(acc[tl][0], af[tl], bf[0], bf[1]);
                if constexpr (NI == 2) mma16816(acc[tl][NI - 1], af[tl], bf[2], bf[3]);
            }
        }

        // 3. gates in registers.  acc[0][ni][e] = i, acc[0][ni][2+e] = f, acc[1][ni][e] = g, acc[1][ni][2+e] = o
#pragma unroll
        for (int ni = 0; ni < NI; ni++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int  u   = nh * (8 * NI) + ni * 8 + (lane & 3) * 2 + e;
                const bool act = s < ulen[ni][e];
                float hval = 0.f;
                if (act) {
                    const float4 x = xp[ni][e];
                    const float gi = sigmoidf_ref(x.x + (acc[0][ni][e] + bh[0]));       // model.cpp:64 association
                    const float gf = sigmoidf_ref(x.y + (acc[0][ni][2 + e] + bh[1]));
                    const float gg = tanhf(x.z + (acc[1][ni][e] + bh[2]));
                    const float go = sigmoidf_ref(x.w + (acc[1][ni][2 + e] + bh[3]));
                    const float c  = gf * cst[ni][e] + gi * gg;
                    cst[ni][e]     = c;
                    hval           = tanhf(c) * go;
                    const size_t row = (size_t) ub[ni][e] * p.Lmax + tt[ni][e];
                    p.out[row * p.ldo + p.coff + dir * H + unit_g] = hval;
                    if (p.outH) p.outH[row * p.ldoh + p.coffh + dir * H + unit_g] = __float2half_rn(hval);
                }
                sSt[u * UNITS + ul] = __float2half_rn(hval);
            }
        __syncthreads();

        // 4. broadcast the 32x32 fp16 slice into every CTA's next-step h buffer (DSMEM, 16 B stores)
#pragma unroll
        for (int j = 0; j < NBT / 8; j++) {
            const int q = tid + j * 256;
            const int dest = q / (NBT * 4), rem = q % (NBT * 4), u = rem >> 2, part = rem & 3;
            const int4 v = *reinterpret_cast<const int4 *>(sSt + u * UNITS + part * 8);
            __half * dst_local = hnxt + u * LDW + rank * UNITS + part * 8;
            if constexpr (ASYNC) {
                if (s + 1 < nsteps) {   // the last step's h is never consumed
                    unsigned ra, rb;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((unsigned) __cvta_generic_to_shared(dst_local)), "r"(dest));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(hbar_s + 8u * ((s + 1) & 1)), "r"(dest));
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                                 ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rb) : "memory");
                }
            } else {
                int4 * dst = reinterpret_cast<int4 *>(cluster.map_shared_rank(dst_local, dest));
                *dst = v;
            }
        }
        if constexpr (ASYNC) {
            if (s + 1 < nsteps) {
                load_xp(s + 1);
                const int bi = (s + 1) & 1;
                const unsigned bar = hbar_s + 8u * bi, par = bi ? hphase1 : hphase0;
                asm volatile(
                    "{\n\t"
                    ".reg .pred P1;\n\t"
                    "LAB_WAIT:\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                    "@P1 bra DONE;\n\t"
                    "bra LAB_WAIT;\n\t"
                    "DONE:\n\t"
                    "}" ::"r"(bar), "r"(par) : "memory");
                if (bi) hphase1 ^= 1u; else hphase0 ^= 1u;
                // re-arm this buffer's barrier for D(s+2) (its data cannot arrive before every thread here has passed the wait above)
                if (tid == 0 && s + 3 < nsteps + 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STEP_BYTES) : "memory");
            }
        } else {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            if (s + 1 < nsteps) load_xp(s + 1);
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
    if constexpr (ASYNC) cluster.sync();   // nobody leaves while a peer might still address its shared memory
}

}  // namespace

int bilstm(Ctx * ctx, const LstmParams & p) {
    if (p.B <= 0 || p.maxLen <= 0) return 0;
    static bool attr_done = false;
    if (!attr_done) {
        B2_CUDA(cudaFuncSetAttribute(bilstm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
        B2_CUDA(cudaFuncSetAttribute(bilstm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
        attr_done = true;
    }
    dim3 grid(CL, 2, cdiv(p.B, NBT));
    ctx->prof_begin(PROF_LSTM, 2.0 * p.B * p.maxLen * 2.0 * 1024.0 * 256.0, 0.0);
    static const bool use_async = getenv("B2TTS_LSTM_BARRIER") == nullptr;   // default: st.async exchange; the cluster-barrier form is kept for A/B runs
    if (use_async) bilstm_kernel<true><<<grid, 256, LSTM_SMEM, ctx->stream>>>(p);
    else           bilstm_kernel<false><<<grid, 256, LSTM_SMEM, ctx->stream>>>(p);
    ctx->prof_end();
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b2
