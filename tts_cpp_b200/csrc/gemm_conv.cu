// gemm_conv.cu -- implicit-GEMM Conv1d / Linear for sm_100a, fp16 operands, fp32 accumulate.
//
// Replaces, for every F16-weight contraction of the Kokoro graphs, the reference's
//   ggml_conv_1d = ggml_im2col (materialised IC*K x L unfold) + ggml_mul_mat   (ggml/src/ggml.c:3870-3894)
//   ggml_mul_mat with F16 src0 (activations re-rounded to fp16)               (ggml/src/ggml-cpu/ggml-cpu.c:262-267,7683-7720)
// and folds the bias add / residual adds / divide / activation nodes that follow them in
// src/models/kokoro/model.cpp (e.g. :120-132, :151-162, :991-1004, :232-238) into the epilogue.
//
// No im2col buffer exists: the A tile of tap k is the activation tile shifted by k*dil - pad rows
// (channels-last layout makes each tap a contiguous row segment), zero-filled outside the utterance.
//
// Tile 128(M = time) x BN(N = Cout) x 32(K), 8 warps, 4-stage cp.async pipeline, ldmatrix + mma.sync.m16n8k16.
// (Round-1 tensor path; the tcgen05/TMA version of this kernel is the round-2 item -- see DESIGN.md.)
#include "common.cuh"
#include <cstdio>

namespace b2 {

namespace {

constexpr int BM = 128, BK = 32, STAGES = 4, LDS = 40;  // LDS: smem row stride in halves (80 B -> conflict-free ldmatrix)

__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gsrc, int src_bytes) {
    unsigned s = (unsigned) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldsm_x4(unsigned & r0, unsigned & r1, unsigned & r2, unsigned & r3, const void * p) {
    unsigned s = (unsigned) __cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(s));
}
__device__ __forceinline__ void mma16816(float * c, const unsigned * a, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ggml's GELU for F32 tensors is an fp16 lookup table of the tanh approximation (ggml-cpu.c:1816-1830)
__device__ __forceinline__ float gelu_f16lut(float x) {
    if (x <= -10.0f) return 0.0f;
    if (x >= 10.0f) return x;
    float xh = __half2float(__float2half_rn(x));
    float y  = 0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh)));
    return __half2float(__float2half_rn(y));
}

template <int BN>
__global__ void __launch_bounds__(256) conv_gemm_kernel(const ConvGemmParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half * sA = reinterpret_cast<__half *>(smem_raw);           // [STAGES][BM][LDS]
    __half * sB = sA + STAGES * BM * LDS;                         // [STAGES][BN][LDS]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int M    = p.B * p.LmaxOut;
    const int Ktot = p.KW * p.CinPad;
    const int nkb  = Ktot / BK;

    // ---- loader bookkeeping: each thread always loads the same A rows (2 x 16 B chunks per stage)
    const __half * a_base[2];
    int a_t0[2], a_len[2];
    bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int row = (tid + i * 256) >> 2;
        const int r   = m0 + row;
        a_ok[i]       = r < M;
        const int b   = a_ok[i] ? r / p.LmaxOut : 0;
        const int t   = r - b * p.LmaxOut;
        a_len[i]      = p.lenIn ? p.lenIn[b] : p.LmaxIn;
        a_t0[i]       = t * p.stride - p.pad;
        a_base[i]     = p.A + (size_t) b * p.LmaxIn * p.lda;
    }
    const int kc = tid & 3;

    auto load_stage = [&](int stage, int kb) {
        const int kk  = kb * BK;
        const int tap = kk / p.CinPad;
        const int ci0 = kk - tap * p.CinPad + kc * 8;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int row = (tid + i * 256) >> 2;
            const int ts  = a_t0[i] + tap * p.dil;
            const bool ok = a_ok[i] && ts >= 0 && ts < a_len[i];
            const __half * src = ok ? a_base[i] + (size_t) ts * p.lda + ci0 : p.A;
            cp_async16(sA + ((size_t) stage * BM + row) * LDS + kc * 8, src, ok ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < BN / 64; i++) {
            const int row = (tid + i * 256) >> 2;
            const bool ok = (n0 + row) < p.Npad;
            const __half * src = ok ? p.W + (size_t) (n0 + row) * Ktot + kk + kc * 8 : p.W;
            cp_async16(sB + ((size_t) stage * BN + row) * LDS + kc * 8, src, ok ? 16 : 0);
        }
    };

    constexpr int WN = BN / 4;      // columns per warp
    constexpr int NI = WN / 8;      // n8 tiles per warp (4 or 2)
    const int wm = warp & 1, wn = warp >> 1;
    float acc[4][NI][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NI; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[i][j][e] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }

    for (int kb = 0; kb < nkb; kb++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kb + STAGES - 1;
            if (nk < nkb) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const int st = kb % STAGES;
        const __half * cA = sA + (size_t) st * BM * LDS;
        const __half * cB = sB + (size_t) st * BN * LDS;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ks++) {
            unsigned af[4][4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) {
                const int r = wm * 64 + mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = ks * 16 + (lane >> 4) * 8;
                ldsm_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], cA + r * LDS + c);
            }
            unsigned bf[NI][2];
#pragma unroll
            for (int nj = 0; nj < NI / 2; nj++) {
                const int r = wn * WN + nj * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = ks * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(bf[2 * nj][0], bf[2 * nj][1], bf[2 * nj + 1][0], bf[2 * nj + 1][1], cB + r * LDS + c);
            }
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) mma16816(acc[mi][ni], af[mi], bf[ni][0], bf[ni][1]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const int r = m0 + wm * 64 + mi * 16 + (lane >> 2) + hh * 8;
            if (r >= M) continue;
            const int b = r / p.LmaxOut;
            const int t = r - b * p.LmaxOut;
            const int lo = p.lenOut ? p.lenOut[b] : p.LmaxOut;
            if (t >= lo) continue;
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int c = n0 + wn * WN + ni * 8 + (lane & 3) * 2 + e;
                    if (c >= p.N) continue;
                    float v = acc[mi][ni][hh * 2 + e];
                    if (p.bias) v = v + p.bias[c];
                    if (p.add1) v = p.add1[(size_t) r * p.ldadd1 + c] + v;
                    if (p.add2) v = p.add2[(size_t) r * p.ldadd2 + c] + v;
                    if (p.div != 0.f) v = __fdiv_rn(v, p.div);
                    if (p.act == ACT_GELU_F16LUT) v = gelu_f16lut(v);
                    else if (p.act == ACT_EXP_SIN_11) v = (c < 11) ? expf(v) : sinf(v);
                    else if (p.act == ACT_TANH) v = tanhf(v);
                    else if (p.act == ACT_LRELU_02) v = (v > 0.f ? v : 0.f) + 0.2f * (v < 0.f ? v : 0.f);
                    if (p.outF) p.outF[(size_t) r * p.ldo + p.coff + c] = v;
                    if (p.outH) p.outH[(size_t) r * p.ldoh + p.coffh + c] = __float2half_rn(v);
                }
            }
        }
    }
}

}  // namespace

int conv_gemm(Ctx * ctx, const ConvGemmParams & p) {
    if (p.CinPad % BK != 0 || p.lda % 8 != 0 || p.Npad % 64 != 0 || p.N > p.Npad) {
        set_error("conv_gemm: bad shape CinPad=%d lda=%d N=%d Npad=%d", p.CinPad, p.lda, p.N, p.Npad);
        return 1;
    }
    const int64_t M = (int64_t) p.B * p.LmaxOut;
    if (M <= 0) return 0;
    {   // the tcgen05 + TMA kernel takes every shape it supports (stride 1, Cin % 64 == 0, Cout >= 128); the rest stays here
        const int r = conv_umma(ctx, p);
        if (r != 2) return r;
    }
    static bool attr_done = false;
    constexpr int smem128 = STAGES * (BM + 128) * LDS * 2, smem64 = STAGES * (BM + 64) * LDS * 2;
    if (!attr_done) {
        B2_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128));
        B2_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64));
        attr_done = true;
    }
    {
        const double rows = (double) (p.validRows ? p.validRows : M), cin = (double) (p.CinTrue ? p.CinTrue : p.CinPad);
        // algorithmic work of this launch: 2*rows*N*KW*Cin flops; bytes = fp16 operand once + weights once + fp32 result once
        snprintf(ctx->tag, sizeof(ctx->tag), "%s N%d K%d C%d L%d B%d d%d", "mmasync", p.N, p.KW, p.CinPad, p.LmaxOut, p.B, p.dil);
        ctx->prof_begin(PROF_GEMM, 2.0 * rows * p.N * p.KW * cin, rows * cin * 2.0 + (double) p.N * p.KW * cin * 2.0 + rows * p.N * 4.0);
    }
    if (p.N > 64) {
        dim3 grid(cdiv(M, BM), cdiv(p.Npad, 128));
        conv_gemm_kernel<128><<<grid, 256, smem128, ctx->stream>>>(p);
    } else {
        dim3 grid(cdiv(M, BM), 1);
        conv_gemm_kernel<64><<<grid, 256, smem64, ctx->stream>>>(p);
    }
    ctx->prof_end();
    ctx->mma_sync_launches++;
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b2
