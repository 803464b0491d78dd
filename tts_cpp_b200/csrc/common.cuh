// common.cuh -- shared declarations of the B200 (sm_100a) TTS hot-path library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstddef>
#include <vector>

namespace b2 {

void set_error(const char * fmt, ...);

// optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline accounting)
enum { PROF_GEMM = 0, PROF_LSTM = 1, PROF_NORM = 2, PROF_CONVT = 3, PROF_KINDS = 4 };
struct ProfRec { cudaEvent_t a, b; int kind; double flops, bytes; char tag[48]; };

struct Ctx {
    int          device   = 0;
    cudaStream_t stream   = nullptr;
    uint64_t     launches = 0;   // kernels launched by this library on this context
    uint64_t     umma_launches = 0, mma_sync_launches = 0;   // conv_gemm dispatch: tcgen05 kernel vs mma.sync fallback
    bool                     prof = false;
    char                     tag[48] = "";   // optional label of the next profiled launch (B2TTS_PROF_DUMP diagnostics)
    std::vector<ProfRec>     recs;
    std::vector<cudaEvent_t> pool;
    void prof_begin(int kind, double flops, double bytes) {
        if (!prof) return;
        ProfRec r; r.kind = kind; r.flops = flops; r.bytes = bytes;
        for (int i = 0; i < 48; i++) r.tag[i] = tag[i];
        tag[0] = 0;
        for (cudaEvent_t * e : { &r.a, &r.b }) {
            if (!pool.empty()) { *e = pool.back(); pool.pop_back(); } else cudaEventCreate(e);
        }
        cudaEventRecord(r.a, stream);
        recs.push_back(r);
    }
    void prof_end() { if (prof && !recs.empty()) cudaEventRecord(recs.back().b, stream); }
};

#define B2_CUDA(x)                                                                                          \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess) {                                                                            \
            b2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_));               \
            return 1;                                                                                       \
        }                                                                                                   \
    } while (0)

#define B2_LAUNCH_CHECK(ctx)                                                                                \
    do {                                                                                                    \
        (ctx)->launches++;                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                                                \
        if (e_ != cudaSuccess) {                                                                            \
            b2::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));        \
            return 1;                                                                                       \
        }                                                                                                   \
    } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int) ((a + b - 1) / b); }
static inline int round_up(int a, int m) { return (a + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM Conv1d / Linear on fp16 operands with fp32 accumulation (gemm_conv.cu).
//   out[b][t][co] = epilogue( sum_{k<KW} sum_{ci<CinPad} A[b][t*stride + k*dil - pad][ci] * W[co][k][ci] )
// A    : fp16 activations, channels-last, rows (b*LmaxIn + t), row stride lda (>= CinPad, multiple of 8);
//        rows outside [0, lenIn[b]) read as zero (this is the conv zero padding and the ragged-batch mask)
// W    : fp16 weights [Npad][KW*CinPad] (pad rows / pad channels are zero)
// epilogue: v = acc + bias[co]; v = add1 + v; v = add2 + v; v = v / div; v = act(v); store fp32 and/or fp16
// The reference computes these contractions as ggml_mul_mat / im2col+mul_mat with F16 weights: activations
// re-rounded to fp16, products accumulated in fp32 (ggml-cpu.c:262-267, ggml.c:3870-3894).
// ---------------------------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_GELU_F16LUT = 1, ACT_EXP_SIN_11 = 2, ACT_LRELU_02 = 3, ACT_TANH = 4 };

struct ConvGemmParams {
    const __half * A       = nullptr;
    const __half * W       = nullptr;
    const float *  bias    = nullptr;
    float *        outF    = nullptr;  int ldo = 0;  int coff = 0;
    __half *       outH    = nullptr;  int ldoh = 0; int coffh = 0;
    const float *  add1    = nullptr;  int ldadd1 = 0;
    const float *  add2    = nullptr;  int ldadd2 = 0;
    float          div     = 0.f;
    int            act     = ACT_NONE;
    int            B = 1, LmaxIn = 0, LmaxOut = 0;
    const int *    lenIn   = nullptr;   // device, per utterance; nullptr -> LmaxIn
    const int *    lenOut  = nullptr;   // device, per utterance; nullptr -> LmaxOut
    int            N = 0, Npad = 0, KW = 1, CinPad = 0, lda = 0, stride = 1, dil = 1, pad = 0;
    int            CinTrue = 0;         // un-padded input channels (roofline accounting only; 0 -> CinPad)
    int64_t        validRows = 0;       // sum of lenOut (roofline accounting only; 0 -> B*LmaxOut)
    bool           tailClean = false;   // the producer of A (adain_apply) already zeroed the rows past each utterance's end: skip zero_tail_rows
    float *        statsPart = nullptr; // optional [B][ceil(LmaxOut/conv_umma_tile_m)][N][2]: per-tile (sum, sum of squares) of the stored values over valid rows
                                        // (tcgen05 kernel only; the caller checks Ctx::umma_launches to know it was produced)
};
int conv_gemm(Ctx * ctx, const ConvGemmParams & p);
int conv_umma_tile_m(const ConvGemmParams & p);   // rows per tcgen05 work item for this shape (128 or 256; 0 = unsupported): statsPart is [B][ceil(LmaxOut / tile_m)][N][2]
int conv_umma(Ctx * ctx, const ConvGemmParams & p);   // gemm_umma.cu: 0 launched, 1 error, 2 shape unsupported (fallback)

// ---------------------------------------------------------------------------------------------
// Persistent cluster bi-LSTM (lstm.cu).  Hidden size 256.
//   xp  : [B][Lmax][2 dirs][256 units][4 gates] fp32 = W_ih x + b_ih (gate order i,f,g,o), from conv_gemm
//   whh : [2][1024][256] fp16, row = gate*256 + unit (the GGUF layout)
//   bhh : [2][1024] fp32
//   out : [B][Lmax][ldo] fp32, channels [coff + dir*256 + unit]; optional fp16 copy (ldoh/coffh) for the next GEMM
// ---------------------------------------------------------------------------------------------
struct LstmParams {
    const float *  xp = nullptr;
    const __half * whh = nullptr;
    const float *  bhh = nullptr;
    float *        out = nullptr;  int ldo = 512;  int coff = 0;
    __half *       outH = nullptr; int ldoh = 0;   int coffh = 0;
    const int *    len = nullptr;   // device [B]
    int            B = 0, Lmax = 0, maxLen = 0;
};
int bilstm(Ctx * ctx, const LstmParams & p);

}  // namespace b2
