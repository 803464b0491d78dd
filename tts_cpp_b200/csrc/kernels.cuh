// kernels.cuh -- launchers of the non-GEMM kernels (elementwise.cu, source.cu).  All activations are
// channels-last, batch-major, padded to the longest utterance: X[b][t][c] = X[(b*Lmax + t)*ld + c]; rows
// t >= len[b] are never read by any consumer.
#pragma once
#include "common.cuh"

namespace b2 {

// ---- InstanceNorm over time + AdaIN + activation (model.cpp:93-101,142-149; ggml_norm ggml-cpu.c:7114-7163)
// sums[b][c][2] (double): sum, sum of squares over t < len[b]   (zeroed by the launcher)
int inorm_stats(Ctx * ctx, const float * x, int ldx, int C, int B, int Lmax, const int * len, double * sums);

// per-tile partials written by the tcgen05 GEMM epilogue ([B][n_mt][C][2] fp32) -> sums[b][c][2] (double), same format as inorm_stats
int stats_finalize(Ctx * ctx, const float * part, int B, int n_mt, int C, double * sums);

enum { NACT_LRELU02 = 0, NACT_SNAKE = 1, NACT_NONE = 2 };
struct AdainParams {
    const float *  x = nullptr;   int ldx = 0;
    int            C = 0, B = 0, Lmax = 0;
    const int *    len = nullptr;
    const double * sums = nullptr;
    const float *  gb = nullptr;  int ldgb = 0, goff = 0, boff = 0;   // gamma/beta of utterance b: gb[b*ldgb + goff/boff + c]
    int            act = NACT_LRELU02;
    const float *  alpha = nullptr;                                    // snake: per channel
    __half *       outH = nullptr; int ldoh = 0, Cpad = 0;            // fp16 GEMM operand (pad channels zeroed)
    float *        outF = nullptr; int ldof = 0;                      // optional fp32 copy
};
int adain_apply(Ctx * ctx, const AdainParams & p);

// depthwise ConvTranspose1d k3 s2 p1 op1 + bias ("pool", model.cpp:103-106): x fp32 [B][L][C] -> fp16 [B][2L][Cpad]
int pool_convt(Ctx * ctx, const float * x, int ldx, int C, int B, int Lmax, const int * len, const float * w3, const float * bias,
               __half * outH, int ldoh, int Cpad);

// ---- row LayerNorm over channels (ggml_norm over ne0) with the three affine flavours the graphs use
enum { LN_AFFINE = 0, LN_ADA = 1 };
struct RowNormParams {
    const float * x = nullptr;  int ldx = 0;  int C = 0;
    int           B = 0, Lmax = 0;  const int * len = nullptr;
    float         eps = 1e-5f;
    int           mode = LN_AFFINE;
    const float * w = nullptr, * bias = nullptr;                        // LN_AFFINE: y = n*w + b
    const float * gb = nullptr; int ldgb = 0, goff = 0, boff = 0;       // LN_ADA:    y = (n + n*gamma) + beta
    int           lrelu02 = 0;
    float *       outF = nullptr; int ldof = 0, coff = 0;
    __half *      outH = nullptr; int ldoh = 0, coffh = 0;
};
int row_norm(Ctx * ctx, const RowNormParams & p);

// ---- small helpers
// fp32 -> fp16 operand copy with optional nearest x2 upsample along time (ggml_upscale_ext, model.cpp:127) and channel padding
int cast_rows(Ctx * ctx, const float * x, int ldx, int C, int B, int LmaxIn, const int * lenOut, int LmaxOut, int up2, float ns, __half * outH,
              int ldoh, int Cpad);   // ns: leaky-relu slope applied before the cast (1.0f = identity)
// split-fp16 operand for fp32-faithful tensor-core products: out[b][q][0:C] = hi(v), [C:2C] = lo(v), [2C:3C] = hi(v) with
// v = leaky_relu(x[b][q][c], ns), hi = fp16(v), lo = fp16(v - hi); rows q >= len[b] (up to Lq) are written as zeros
int split3_rows(Ctx * ctx, const float * x, int ldx, int C, int B, int LmaxIn, const int * len, float ns, __half * outH, int Lq);
// copy channel slice of fp32 rows: dst[b][t][dcoff + c] = src[b][t][scoff + c], c < C
int copy_cols(Ctx * ctx, const float * src, int lds, int scoff, float * dst, int ldd, int dcoff, int C, int B, int Lmax, const int * len);
// broadcast a per-utterance vector into channel slice: dstF/dstH[b][t][coff + c] = v[b*ldv + c]
int bcast_cols(Ctx * ctx, const float * v, int ldv, int C, int B, int Lmax, const int * len, float * dstF, int ldf, int cofff, __half * dstH,
               int ldh, int coffh);
// gather rows by alignment: dst[b][t][:] = src[b][idx[b][t]][:]  (the one-hot duration-mask matmuls, model.cpp:1163-1164,1206)
int gather_rows(Ctx * ctx, const float * src, int lds, int LmaxSrc, const int * idx, int C, int B, int Lmax, const int * len, float * dstF,
                int ldf, __half * dstH, int ldh, int Cpad);
// fp32 linear y[r][n] = sum_k x[r][k] W[n][k] + bias[n]  (F32 weights: albert.embd, AdaIN gamma/beta projections)
int linear_f32(Ctx * ctx, const float * x, int ldx, const float * W, const float * bias, int rows, int K, int N, float * y, int ldy);
// ALBERT embeddings: (tok_embd[tok] + pos_embd[pos]) + type -> LN(eps)*w + b      (model.cpp:10-21)
int albert_embed(Ctx * ctx, const int * tokens, const int * tok_off, const float * tok_embd, const float * pos_embd, const float * type_embd,
                 const float * nw, const float * nb, int B, int Lmax, const int * len, float * out, int ldo);
// text encoder embedding rows (fp16 table -> fp32 values re-rounded to fp16 for the conv operand)  (model.cpp:1196)
int embed_rows_h(Ctx * ctx, const int * tokens, const int * tok_off, const __half * table, int C, int B, int Lmax, const int * len,
                 __half * outH, int ldoh);
// ALBERT self-attention, fp32 like the reference (F32 x F32 mul_mat + soft_max_ext, model.cpp:974-990): qkv[b][t][3*768] -> fp16 [b][t][768]
int albert_attention(Ctx * ctx, const float * qkv, int B, int Lmax, const int * len, int heads, int hd, float scale, __half * outH, int ldoh);
// duration head tail: lens[b][t] = clamp(round(sum_j sigmoid(logit[b][t][j])), 1, 50)   (model.cpp:1037-1040); also per-utterance totals
int duration_tail(Ctx * ctx, const float * logits, int ldl, int n_bins, int B, int Lmax, const int * len, float * lens_out);
// alignment: from lens[b][n] build idx[b][t] (token of frame t), frame totals T[b]  (model.cpp:1265-1274, 1284-1287)
int build_alignment(Ctx * ctx, const float * lens, int B, int LmaxTok, const int * ntok, int LmaxFrames, int * idx, int * T);
// k3 s2 p1 conv of a 1-channel curve with F16 kernel: decoder f0_conv / n_conv (model.cpp:1216-1217): x[b][2T] -> dst[b][t][coff]
int curve_conv_s2(Ctx * ctx, const float * x, int ldx, int B, const int * lenOut, int LmaxOut, const int * lenIn, const float * w3,
                  const float * bias, float * dst, int ldd, int coff);

// ---- harmonic source / STFT / iSTFT (source.cu)
struct SourceParams {
    const float * f0 = nullptr;        // [B][L2max] f0 curve (2T per utterance)
    int           B = 0, L2max = 0;
    const int *   len2 = nullptr;      // 2T per utterance
    const unsigned long long * noise_skip = nullptr;  // device [B]
    float         w_src[9];            // m_source weight (fp16-representable) and bias
    float         b_src = 0.f;
    float *       phase = nullptr;     // scratch [B][9][L2max]
    float *       har = nullptr;       // [B][Smax]   (Smax = 300*L2max)
    int           Smax = 0;
    float *       sing = nullptr;      // optional tap [B][Smax][9]
};
int source_har(Ctx * ctx, const SourceParams & p);
// har[b][S] -> fp16 operand [b][frame][Cpad] (mag 0..10, phase 11..21) and optional fp32 tap; frames = S/5+1
// FpitchH: rows per utterance of the fp16 operand (>= Fmax; 0 -> Fmax): the strided noise conv wants a multiple of its stride
int stft20(Ctx * ctx, const float * har, int Smax, int B, const int * lenS, int Fmax, __half * outH, int ldoh, int Cpad, float * outF, int ldof,
           int FpitchH = 0);
// rows [len_b, Lmax) of every utterance of a fp16 operand cleared (what a conv must read as zero padding past the end)
int zero_rows_past_end(Ctx * ctx, __half * A, int lda, int C, int B, int Lmax, const int * len);
// spec/phase [b][frame][22] (exp'd magnitude, sin'd phase) -> pcm[b][S]
int istft20(Ctx * ctx, float * specph, int ld, int B, const int * lenF, int Fmax, float * pcm, int Smax);
// stand-alone pieces for the op-level ABI
int op_cumsum(Ctx * ctx, const float * x, int L, int rows, float * y);
int op_unary(Ctx * ctx, int which, const float * x, int64_t n, float arg, float * y);   // 0 mod, 1 round, 2 reciprocal
int op_upscale_linear(Ctx * ctx, const float * x, int L, int rows, int factor, float * y);
int op_snake(Ctx * ctx, const float * alpha, int C, const float * x, int L, float * y);
int op_uniform(Ctx * ctx, unsigned long long skip, int64_t count, float * y);
int op_conv_transpose_1d(Ctx * ctx, const float * w, int K, int coutg, int cin, const float * x, int L, int s, int p, int op, int g, float * y,
                         int Lout);
// ConvTranspose1d for the generator up-convs, channels-last, fp32 math (F32 kernel, ggml-cpu.c:10104-10200):
// x[b][t][Cin] (lrelu applied on load with slope `ns`), w[K][Cin][Cout] (repacked), -> y[b][o][Cout] + bias; optional left reflect pad of 1
int convt_cl(Ctx * ctx, const float * x, int ldx, int Cin, int B, int LmaxIn, const int * lenIn, const float * w, const float * bias, int K,
             int Cout, int stride, int pad, float ns, int reflect1, float * y, int ldy, int LmaxOut, const int * lenOut);

// vad.cu -- apply_energy_voice_inactivity_detection (reference examples/cli/vad.cpp:11-68) for B utterances resident in HBM, back to back in d_pcm.
// d_off[B + 1]: sample offsets; d_eoff[B + 1]: offsets into d_energies (n_b / spf whole frames each); max_frames = the largest frame count; spf, early_frames: the
// reference's samples_per_frame / early_cuttoff_frames (computed by the caller as vad.cpp:20,22 does); d_n_out[b] = the trimmed n_outputs
int vad_trim_rows(Ctx * ctx, const float * d_pcm, const long long * d_off, const long long * d_eoff, int B, int max_frames, int spf, int frame_threshold,
                  float norm_threshold, int trailing, int early_frames, float early_threshold, float * d_energies, long long * d_n_out);

// sampler.cu -- the reference sampler (src/sampler.cpp) on the device for `rows` independent heads: repetition penalty, temperature, top-k, top-p, one draw each.
// temperature <= 0 or do_sample == 0 is sampler::max.  The uniform of a row comes from (seed, row, *d_step) by a counter-based hash, so a captured graph of a
// decode step can be replayed; out[*d_step * rows + row] receives the token (d_step == nullptr: step 0).
struct SampleParams {
    const float * logits = nullptr;   // [rows][V]
    int rows = 0, V = 0;
    int do_sample = 0, top_k = 0;
    float temperature = 1.f, top_p = 1.f, repetition_penalty = 1.f;
    int * last_ids = nullptr;         // [rows], -1 initially (sampler::reset); only used when repetition_penalty != 1
    int * rep_counts = nullptr;       // [rows], 0 initially
    float * scratch = nullptr;        // [rows][V] workspace, needed when top_p < 1 (the full-vocabulary softmax)
    unsigned long long seed = 0;
    const int * d_step = nullptr;
    int * out = nullptr;
    int * cur_tok = nullptr;          // optional [rows]: also receives the token (Orpheus feeds it straight back)
    int out_stride_steps = 0;         // 0: out[step * rows + row]; else out[row * out_stride_steps + step] (Orpheus' [B][n_steps] layout)
};
int sample_rows(Ctx * ctx, const SampleParams & p);
float sample_uniform_host(unsigned long long seed, unsigned long long row, unsigned long long step);   // the uniform row `row` draws at step `step`
constexpr int SAMPLE_MAX_TOP_K = 1024;
// the sampler settings of one generate call (generation_configuration's temperature / top_k / top_p / repetition_penalty / sample, include/common.h:45-66)
struct ArSampling { int do_sample = 0, top_k = 0; float top_p = 1.f, temperature = 1.f, repetition_penalty = 1.f; unsigned long long seed = 0; };
// fills SampleParams' sampling fields and state pointers for `rows` heads of V logits; state / scratch come from the caller's arena (may be null when unused)
static inline SampleParams make_sample_params(const ArSampling & a, const float * logits, int rows, int V, int * last_ids, int * rep_counts, float * scratch, const int * d_step, int * out) {
    SampleParams p;
    p.logits = logits; p.rows = rows; p.V = V; p.do_sample = a.do_sample; p.top_k = a.top_k; p.top_p = a.top_p; p.temperature = a.temperature;
    p.repetition_penalty = a.repetition_penalty; p.seed = a.seed; p.last_ids = last_ids; p.rep_counts = rep_counts; p.scratch = scratch; p.d_step = d_step; p.out = out;
    return p;
}
static inline bool sampling_needs_scratch(const ArSampling & a, int V) { return a.do_sample && (a.top_p < 1.f || !(a.top_k > 0 && a.top_k < V)); }

}  // namespace b2
