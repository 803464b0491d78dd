/* b2tts.h -- C-ABI of the B200-native TTS hot path (libb2tts.so).
 *
 * The reference (mmwillet/TTS.cpp) has no FFI: its seam is the C++ API of static lib `tts`
 * (include/common.h:68-94, src/models/loaders.h:7-20).  This header is the thin C boundary the
 * reference's runners would call instead of building + computing GGML graphs; every entry point
 * names the reference function it replaces (file:line relative to the TTS.cpp tree).  Plain C
 * types only; all pointers are HOST memory owned by the caller unless stated; device memory
 * never crosses this ABI.  Every function returns 0 on success, non-zero on error
 * (b2tts_last_error() gives the text); the C++ shim (tts_cpp_b200/host) maps errors to the
 * reference's abort() convention (src/util.cpp:14-22).
 *
 * A context is one CUDA device + stream; one in-flight call per context (the reference's
 * runners are not re-entrant either: examples/server/server.cpp:316-321 gives each worker its own).
 */
#ifndef B2TTS_H
#define B2TTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2tts_ctx    b2tts_ctx;
typedef struct b2tts_kokoro b2tts_kokoro;

/* ggml type ids used by the weight hand-off (ggml/include/ggml.h enum ggml_type) */
enum { B2TTS_TYPE_F32 = 0, B2TTS_TYPE_F16 = 1 };

/* ---- context (replaces runner_context: src/tts_model.h:16-44, src/tts_model.cpp:38-67) ---- */
int          b2tts_ctx_create(int device, b2tts_ctx ** out);
void         b2tts_ctx_destroy(b2tts_ctx * ctx);
const char * b2tts_last_error(void);
/* number of kernels this library launched on the context since creation (bench.py "gpu_launches") */
uint64_t     b2tts_launch_count(const b2tts_ctx * ctx);
/* raw CUDA stream handle (cudaStream_t) so a host harness can record events on the launching stream */
void *       b2tts_stream(const b2tts_ctx * ctx);
/* per-kernel-class timing with CUDA events on the launching stream (bench.py's live roofline accounting; off by default).
 * kinds: 0 = conv_gemm (tensor-core contractions), 1 = cluster bi-LSTM, 2 = InstanceNorm/AdaIN passes, 3 = ConvTranspose1d.
 * b2tts_prof_enable(ctx, on) clears the records; b2tts_prof_read sums device ms / algorithmic flops / bytes / launches of a kind. */
/* conv/linear dispatch counters: which = 0 -> launches of the tcgen05 + TMA kernel, 1 -> launches of the mma.sync fallback */
uint64_t     b2tts_gemm_launches(const b2tts_ctx * ctx, int which);
int          b2tts_prof_enable(b2tts_ctx * ctx, int on);
int          b2tts_prof_read(b2tts_ctx * ctx, int kind, double * total_ms, double * flops, double * bytes, uint64_t * launches);

/* ---- Kokoro model: weight hand-off --------------------------------------------------------
 * b2tts_kokoro_create       <- kokoro_model::setup_from_file / prep_constants (src/models/kokoro/model.h:297-314, model.cpp:841-930)
 *                              `kv_keys/kv_vals` are the uint32 GGUF metadata (generator paddings/dilations/strides, recurrence ...)
 * b2tts_kokoro_assign_weight<- tts_generation_runner::assign_weight -> kokoro_model::assign_weight (model.cpp:1327-1332, 413-427);
 *                              `name` is the GGUF tensor name (with the "kokoro." prefix), `ne` the ggml dims (ne[0] fastest)
 * b2tts_kokoro_prepare      <- kokoro_runner::prepare_post_load (model.cpp:1244-1252, 310-392): repacks weights for the
 *                              tensor-core kernels, builds the constants (Hann window, harmonic factors, window^2 table)
 * b2tts_kokoro_load_gguf    <- runner_from_file (src/models/loaders.cpp:34-95) for arch "kokoro": does all three from a file
 */
int  b2tts_kokoro_create(b2tts_ctx * ctx, int n_kv, const char * const * kv_keys, const uint32_t * kv_vals, b2tts_kokoro ** out);
int  b2tts_kokoro_assign_weight(b2tts_kokoro * m, const char * name, int ggml_type, int n_dims, const int64_t * ne,
                                const void * data, size_t nbytes);
int  b2tts_kokoro_prepare(b2tts_kokoro * m);
int  b2tts_kokoro_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_kokoro ** out);
void b2tts_kokoro_free(b2tts_kokoro * m);
/* voices (tts_generation_runner::list_voices, model.cpp:1452-1455): returns count; names[i] valid for the model's life */
int  b2tts_kokoro_n_voices(const b2tts_kokoro * m);
const char * b2tts_kokoro_voice_name(const b2tts_kokoro * m, int i);
/* bytes of weights resident in HBM (for the roofline accounting) */
size_t b2tts_kokoro_weight_bytes(const b2tts_kokoro * m);

/* ---- Kokoro forward ------------------------------------------------------------------------
 * b2tts_kokoro_run_batch <- kokoro_runner::run (model.cpp:1277-1325), batched over independent utterances:
 *   duration pass (kokoro_duration_runner::run, model.cpp:1069-1123) -> host sum of durations -> generation pass.
 *   tokens      : concatenated token ids of all utterances (each already wrapped with BOS/EOS, as tokenize_chunks does)
 *   n_tokens[b] : tokens per utterance (3 <= n <= 512)
 *   voice       : voice name (NULL -> "af_heart", model.cpp:1390-1393); style row = n_tokens-3 (model.cpp:1013,1213)
 *   noise_skip[b]: how many draws of the reference's process-wide uniform engine (src/util.cpp:66-72) precede this
 *                 utterance's 9*600*T noise values (0 = first generate() of a fresh process); NULL -> all 0
 *   Outputs are borrowed pointers into context-owned pinned host buffers, valid until the next call on this model
 *   (same lifetime rule as tts_response.data, model.cpp:1299):
 *   pcm[b]      : float32 PCM @24 kHz, n_samples[b] = 600 * sum(durations)
 *   durations   : (optional, may be NULL) receives a pointer to sum(n_tokens) floats: the integral frame counts
 */
int b2tts_kokoro_run_batch(b2tts_kokoro * m, int batch, const uint32_t * tokens, const int32_t * n_tokens, const char * voice,
                           const uint64_t * noise_skip, const float ** pcm, int64_t * n_samples, const float ** durations);
/* The same forward with the PCM LEFT IN DEVICE MEMORY: *pcm_device points at a [batch][*row_stride] float32 block on the context's device (utterance b occupies the
 * first n_samples[b] floats of row b), owned by the model and valid until its next call; the forward is complete when the call returns.  The one entry point where a
 * device pointer crosses the ABI: it exists for the multi-GPU gather ("NCCL/NVLink used only to scatter prompts and gather PCM"), which sends the audio to
 * the gathering rank straight from device memory instead of through two host copies. */
int b2tts_kokoro_run_batch_device(b2tts_kokoro * m, int batch, const uint32_t * tokens, const int32_t * n_tokens, const char * voice, const uint64_t * noise_skip,
                                  const float ** pcm_device, int64_t * row_stride, int64_t * n_samples);
/* The same forward for utterances that are CONSECUTIVE generate() calls of one reference process -- the chunks kokoro_runner::generate makes of a long prompt
 * (model.cpp:1430-1447, run one after another there) or a drained request queue: utterance b's noise continues the process-wide uniform stream where utterance
 * b-1 left it (9 * 600 * T_b draws each), the first at noise_skip_first.  The offsets depend on the durations, so the library sets them between its two passes.
 * Everything else as b2tts_kokoro_run_batch. */
int b2tts_kokoro_run_chunks(b2tts_kokoro * m, int batch, const uint32_t * tokens, const int32_t * n_tokens, const char * voice, uint64_t noise_skip_first,
                            const float ** pcm, int64_t * n_samples, const float ** durations);

/* Stage timings of the last run_batch in milliseconds (CUDA events on the launching stream):
 * [0] duration pass, [1] generation pass (device), [2] whole call incl. H2D/D2H.  */
int b2tts_kokoro_last_timings(const b2tts_kokoro * m, float ms[3]);

/* ---- test taps (teacher forcing / stage parity; no effect on the product path unless used) ----
 * Named stage buffers of the LAST run (batch-major, padded to the longest utterance, channels-last fp32):
 * "albert","d","lens","en","shared","f0","n","t_en","asr","dec","har","har_spec","gen_out0","gen_out1","spec","pcm", ...
 * b2tts_kokoro_tap_info  : rows = batch*padded_len, cols = channels
 * b2tts_kokoro_tap_read  : copy the buffer to host
 * b2tts_kokoro_override  : for the NEXT runs, overwrite the named buffer with host data right after it is produced
 *                          (count floats must match); count==0 clears the override
 */
int b2tts_kokoro_set_taps(b2tts_kokoro * m, int enable);
int b2tts_kokoro_tap_info(b2tts_kokoro * m, const char * name, int64_t * rows, int64_t * cols, int64_t * padded_len);
int b2tts_kokoro_tap_read(b2tts_kokoro * m, const char * name, float * dst, size_t count);
int b2tts_kokoro_override(b2tts_kokoro * m, const char * name, const float * src, size_t count);

/* ---- the patched ggml ops as stand-alone device kernels (host buffers in/out) -------------------
 * These replace the ops the reference adds to its ggml fork (constructors ggml/src/ggml.c:2204-2345,3847-3945,4223-4233;
 * CPU kernels ggml/src/ggml-cpu/ggml-cpu.c:5526-5720,8476-8860,10004-10220,10849-10984) and the util.cpp wrappers
 * (src/util.cpp:86-137).  Layouts follow ggml: x[c][l] = data[c*L + l] ("ne = [L, C]").
 */
/* ggml_conv_transpose_1d(kernel[K,Cout/g,Cin], x[L,Cin], s, p, d=1, op, g): y[Lout,Cout], Lout=(L-1)s-2p+(K-1)+op+1; fp32 math */
int b2tts_op_conv_transpose_1d(b2tts_ctx * ctx, const float * kernel, int K, int cout_per_group, int cin, const float * x, int L,
                               int stride, int pad, int out_pad, int groups, float * y);
/* ggml_conv_1d (im2col + mul_mat): kernel[K,Cin,Cout]; f16_kernel!=0 re-rounds kernel and activations to fp16 like the reference */
int b2tts_op_conv_1d(b2tts_ctx * ctx, const float * kernel, int K, int cin, int cout, const float * x, int L, int stride, int pad,
                     int dil, int f16_kernel, float * y);
int b2tts_op_cumsum(b2tts_ctx * ctx, const float * x, int L, int rows, float * y);            /* ggml_cumsum along ne0 */
/* sampler::sample / sampler::max (reference src/sampler.cpp:3-70,187-204) for `rows` independent heads of `vocab` logits: repetition penalty, temperature,
 * top-k (<= 1024), top-p, one draw each.  do_sample == 0 is sampler::max.  The reference seeds a fresh generator from std::random_device per call; here row r
 * draws the uniform b2tts_sample_uniform(seed, r, step).  last_ids / rep_counts (may be NULL when repetition_penalty == 1): sampler::last_token_ids /
 * repetition_counts, updated in place (-1 / 0 after sampler::reset). */
int   b2tts_op_sample(b2tts_ctx * ctx, const float * logits, int rows, int vocab, int do_sample, int top_k, float top_p, float temperature, float repetition_penalty,
                      int32_t * last_ids, int32_t * rep_counts, uint64_t seed, int step, int32_t * tokens);
float b2tts_sample_uniform(uint64_t seed, uint64_t row, uint64_t step);
/* apply_energy_voice_inactivity_detection (reference examples/cli/vad.cpp:11-68, the `tts-cli --vad` trailing-silence trim) for a batch of B utterances laid back to
 * back in `pcm` (n_samples[b] floats each): n_out[b] = the data.n_outputs the reference leaves behind (its size_t as int64 two's complement).  The argument list is
 * the reference function's (vad.h:13-22, same defaults: 44100, 10, 20, 0.01, 5, 3, 0.1).  energies_out (may be NULL): the frame energies (vad.cpp:3-9), per
 * utterance n_samples[b] / samples_per_frame floats, back to back.  Errors where the reference divides by zero (ms_per_frame <= 0, samples_per_frame < 1). */
int b2tts_op_vad_trim(b2tts_ctx * ctx, const float * pcm, const int64_t * n_samples, int B, float sample_rate, int ms_per_frame, int frame_threshold,
                      float normalized_energy_threshold, int trailing_silent_frames, int early_cutoff_seconds_threshold, float early_cutoff_energy_threshold,
                      int64_t * n_out, float * energies_out);
int b2tts_op_mod(b2tts_ctx * ctx, const float * x, int64_t n, float mod_val, float * y);      /* ggml_mod: fmod */
int b2tts_op_round(b2tts_ctx * ctx, const float * x, int64_t n, float * y);                   /* ggml_round: (float)(int)(x+0.5f) */
int b2tts_op_reciprocal(b2tts_ctx * ctx, const float * x, int64_t n, float * y);              /* ggml_reciprocal */
int b2tts_op_upscale_linear(b2tts_ctx * ctx, const float * x, int L, int rows, int factor, float * y); /* ggml_upscale_linear */
int b2tts_op_snake(b2tts_ctx * ctx, const float * alpha, int C, const float * x, int L, float * y);    /* snake_1d */
/* stft(x[L], hann(n_fft), n_fft, hop, abs_and_angle=1, one_sided=1): mag,phase [frames][n_fft/2+1], frames = L/hop+1 */
int b2tts_op_stft(b2tts_ctx * ctx, const float * x, int L, int n_fft, int hop, float * mag, float * phase);
/* istft(mag,phase [frames][bins]) / window_sq_sum  -> y[(frames-1)*hop] */
int b2tts_op_istft(b2tts_ctx * ctx, const float * mag, const float * phase, int frames, int n_fft, int hop, float * y);
/* the reference's static uniform generator (src/util.cpp:66-72): draws skip+1 .. skip+count */
int b2tts_op_uniform(b2tts_ctx * ctx, uint64_t skip, int64_t count, float * y);
/* bidirectional LSTM (src/models/kokoro/model.cpp:35-86) on a ragged batch, used by the stage parity tests:
 * w_ih[2][4*H][In], w_hh[2][4*H][H], b_ih[2][4*H], b_hh[2][4*H] (dir 0 = forward; gate order i,f,g,o; fp16-representable values),
 * x[B][Lmax][In], len[B] -> y[B][Lmax][2H];  H must be 256 */
int b2tts_op_bilstm(b2tts_ctx * ctx, const float * w_ih, const float * w_hh, const float * b_ih, const float * b_hh, int In, int H,
                    const float * x, int B, int Lmax, const int32_t * len, float * y);

/* ------------------------------------------------------------------------------------------------------------------
 * DAC neural audio codec decoder: codebook indices -> PCM, batched over independent utterances.
 *   b2tts_dac_load_gguf  : dac_model::setup_from_file + the assign_weight loop over the "audio_encoder.*" tensors + prepare_post_load
 *                          (reference src/decoder/dac_model.h:40-44, src/decoder/dac_model.cpp:14-98,139-144; src/models/parler/loader.cpp:12-20,
 *                          src/models/loaders.cpp:79-89).  Works on a Parler / Dia GGUF or on a file holding only the codec.
 *   b2tts_dac_decode_batch : dac_runner::run (src/decoder/dac_model.cpp:172-212) for n_utterances at once.  codes[b] = frames[b] * n_heads
 *                          indices, frame-major (index of head h at frame t is codes[b][t*n_heads + h], the layout the reference takes);
 *                          pcm[b] = frames[b] * up_sampling_factor samples in a runner-owned pinned host buffer, valid until the next call
 *                          on this model (the lifetime of tts_response.data, dac_model.cpp:191).
 * All return 0 on success; on failure b2tts_last_error() holds the message (the reference would TTS_ABORT). */
typedef struct b2tts_dac b2tts_dac;
int   b2tts_dac_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_dac ** out);
void  b2tts_dac_free(b2tts_dac * m);
int   b2tts_dac_info(const b2tts_dac * m, int * n_heads, int * up_sampling_factor, int * codebook_size);
int   b2tts_dac_decode_batch(b2tts_dac * m, int n_utterances, const uint32_t * const * codes, const int32_t * frames, const float ** pcm, int64_t * n_samples);
float b2tts_dac_last_ms(const b2tts_dac * m);   /* device time of the last decode_batch (CUDA events), for bench / tests */

/* ------------------------------------------------------------------------------------------------------------------
 * SNAC neural audio codec decoder, batched (PCM 3.6e-6 RMS against the reference on a B200, tests/test_snac_gpu.py).
 *   b2tts_snac_load_gguf   : snac_model::setup_from_file + assign_weight loop over the "snac.*" tensors + prepare_post_load
 *                            (reference src/decoder/snac_model.h:41-46, snac_model.cpp:3-84,124-127; src/models/orpheus/loader.cpp:12-18)
 *   b2tts_snac_decode_batch: snac_runner::run (snac_model.cpp:181-208).  codes[b] = the three vectors the reference takes, concatenated:
 *                            L/4 coarse, L/2 medium, L fine indices, L = fine_frames[b] (a multiple of 4); pcm[b] = L * 512 samples in a
 *                            runner-owned pinned buffer.  The injected noise is the reference's process-wide normal generator
 *                            (src/util.cpp:74-80): utterance b of a call continues the stream where utterance b-1 (or the previous call) left it;
 *   b2tts_snac_reset_noise : rewinds that stream to the state of a fresh process. */
typedef struct b2tts_snac b2tts_snac;
int   b2tts_snac_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_snac ** out);
void  b2tts_snac_free(b2tts_snac * m);
int   b2tts_snac_info(const b2tts_snac * m, int * up_sampling_factor, int * codebook_size);
int   b2tts_snac_decode_batch(b2tts_snac * m, int n_utterances, const uint32_t * const * codes, const int32_t * fine_frames, const float ** pcm, int64_t * n_samples);
float b2tts_snac_last_ms(const b2tts_snac * m);   /* device time of the last decode_batch (CUDA events on the context's stream), like b2tts_dac_last_ms */
int   b2tts_snac_reset_noise(b2tts_snac * m);

/* ------------------------------------------------------------------------------------------------------------------
 * Orpheus autoregressive decode.  Two paths behind the same calls: F16 or Q8_0 matrices (all of one kind), greedy (calls with more than 16 sequences run as groups of 16), hidden <= 3 072 run decode steps 1 .. n-1 inside
 * the PERSISTENT DECODE KERNEL (csrc/pdk.cuh: RMSNorm folded into the staging, NeoX RoPE + cache append and SwiGLU in the GEMV epilogues, paged fp16 GQA cache, chunked
 * argmax; Q8_0: activations quantised per 32-block once per phase, int8 MMA, fp32 scale products = ggml_vec_dot_q8_0_q8_0; on a B200 the reference's tokens over 72
 * steps, logits 7.5e-3 (F16); Orpheus-3B shape F16 3.3-5.0 ms per step at 1-16 sequences); everything else runs the launch-per-op
 * path: batched GEMVs over F32 (the reference's only Orpheus dtype), F16 or Q4_0 / Q5_0 / Q8_0 matrices,
 * compact GQA KV cache, device argmax / sampler / stop rule, CUDA-graph replay of a decode step.  Hardware status (B200, tests/test_orpheus_gpu.py, all green): F32 -- the
 * reference's token ids exactly, logits 5e-6, plain and fp32-faithful tensor-core (B2TTS_AR_MMA=1) variants; F16 / Q8_0 -- no reference output exists (its runtime is
 * F32-only), they track the reference's F32 run within the storage format's noise (4e-4 / 1.7e-2 of the logit std, same greedy tokens).  Orpheus-3B-shaped q8_0
 * (BASELINE config 5) measured by `bench.py --workload orpheus`.  The same .cu file is also checked under the CPU emulation of tests/emu.
 *   b2tts_orpheus_load_gguf      : orpheus_model::setup_from_file + assign_weight loop over "orpheus.*" (reference
 *                                  src/models/orpheus/model.h:59-63, model.cpp:11-120; loader.cpp:8-23)
 *   b2tts_orpheus_generate_greedy: generate_from_batch's decode + sampler loop (model.cpp:230-353,389-398; sampler::max) for n_sequences
 *                                  independent prompts of token ids, n_steps tokens each, without the stop condition.
 *                                  out_tokens [n_sequences][n_steps]; out_logits (may be NULL) [n_sequences][n_steps][vocab]. */
/* sampler settings of one generate call: generation_configuration's sample / top_k / top_p / temperature / repetition_penalty (reference include/common.h:45-66).
 * do_sample == 0 is sampler::max.  seed: head r of step s draws the uniform b2tts_sample_uniform(seed, r, s) (see b2tts_op_sample). */
typedef struct b2tts_sampling { int32_t do_sample; int32_t top_k; float top_p; float temperature; float repetition_penalty; uint64_t seed; } b2tts_sampling;

typedef struct b2tts_orpheus b2tts_orpheus;
int   b2tts_orpheus_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_orpheus ** out);
void  b2tts_orpheus_free(b2tts_orpheus * m);
int   b2tts_orpheus_info(const b2tts_orpheus * m, int * vocab_size, int * n_layers, int * hidden_size);
int   b2tts_orpheus_generate_greedy(b2tts_orpheus * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps,
                                    int32_t * out_tokens, float * out_logits);
/* the same loop under the reference sampler's settings (sampler.cpp on the device, sampler.cu); sampling == NULL is the greedy call above */
int   b2tts_orpheus_generate(b2tts_orpheus * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const b2tts_sampling * sampling,
                             int32_t * out_tokens, float * out_logits);
/* generate_from_batch's loop WITH its stop condition (model.cpp:389-398: stop at orpheus.stopping_token_id or after max_steps tokens): per sequence n_generated[b]
 * tokens were produced, the last of them the stopping token when it came (tokens past it are zero); the batch stops stepping once every sequence has ended (checked
 * every 32 steps).  Matrices may be F32 (the reference's only Orpheus dtype), F16 or Q4_0 / Q5_0 / Q8_0 blocks (BASELINE config 5 runs q8_0). */
int   b2tts_orpheus_generate_until_stop(b2tts_orpheus * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int max_steps,
                                        const b2tts_sampling * sampling, int32_t * out_tokens, int32_t * n_generated);
int   b2tts_orpheus_set_stopping_token(b2tts_orpheus * m, int token_id);   /* overrides orpheus.stopping_token_id of the GGUF (model.h:43: 128258) */
size_t b2tts_orpheus_step_weight_bytes(const b2tts_orpheus * m);   /* W_step of SURVEY 8(d): bytes of the weight tensors one decode step touches, each once, stored dtype */
/* F16 or Q8_0 GGUFs, greedy (groups of 16 sequences), hidden <= 3 072: decode steps 1 .. n-1 run inside the persistent decode kernel (csrc/pdk.cuh; B2TTS_AR_PDK=0: launch per op);
 * -> cooperative launches so far and the decode steps they covered */
void  b2tts_orpheus_pdk_stats(const b2tts_orpheus * m, uint64_t * launches, uint64_t * steps);
size_t b2tts_orpheus_weight_bytes(const b2tts_orpheus * m);   /* bytes resident in HBM (B2TTS_AR_MMA=1 adds the fp16 split copies of the matrices) */
float b2tts_orpheus_last_ms(const b2tts_orpheus * m);

/* ------------------------------------------------------------------------------------------------------------------
 * Parler-TTS autoregressive decode.  Two paths behind the same calls: F16 GGUFs (BASELINE config 3), greedy or teacher-forced (calls with more than 16 sequences run as groups of 16) run
 * the PERSISTENT DECODE KERNEL (csrc/pdk.cuh: one cooperative launch per 32 decode steps, TMA weight ring, paged fp16 KV cache, LayerNorm / GELU / residual / KV append
 * folded into the GEMV phases); everything else (F32 / Q8_0 / Q5_0 / Q4_0 matrices, sampling, larger batches, B2TTS_AR_PDK=0) the launch-per-op path (tensor-core GEMV,
 * CUDA-graph replay).  Hardware status (B200, tests/test_parler_gpu.py + test_ar_fullsize_gpu.py, all green): the reference's token ids exactly on every F32 / F16
 * variant of either path (logits 1.5e-3 / 9e-3 at a logit std of 4 -- ggml's fp16 GELU table and F16 activation rounding), at the Parler-Mini size too
 * (tests/golden/parler_mini_vectors.npz); quantised GGUFs teacher-forced within 0.1 RMS; the stop rule against reference runs to completion.
 *   b2tts_parler_load_gguf      : parler_tts_model::setup_from_file + assign_weight loop over "decoder.*" + prep_cross_key_values (reference
 *                                 src/models/parler/model.cpp:3-28,110-173,271-318; parler/loader.cpp)
 *   b2tts_parler_generate_greedy: generate_from_batch's prompt decode, then the audio decode + sampler loop with the delay pattern
 *                                 (model.cpp:387-470,520-614,762-786; sampler::max per output head) for n_sequences independent prompts of token
 *                                 ids sharing the model's stored conditional-prompt encoding, n_steps frames each, without check_stopping.
 *                                 out_tokens [n_sequences][n_steps][n_heads]; out_logits (may be NULL) [n_sequences][n_steps][n_heads][out_vocab]. */
/* ---- T5 conditional-prompt encoder (reference src/models/parler/t5/model.cpp): the pass parler_tts_runner::update_conditional_prompt makes before
 * prep_cross_key_values (src/models/parler/model.cpp:510-518), on the GPU.
 *   b2tts_t5_load_gguf : text_encoder_from_file (t5/model.cpp:365-400) below the tokenizer: hyper-parameters (prep_constants, :118-158), the assign_weight loop over
 *                        the "t5encoder.*" tensors (F32, F16, Q8_0 / Q5_0 / Q4_0 matrices), prepare_post_load
 *   b2tts_t5_encode    : t5_runner::run (t5/model.cpp:336-363) = build_t5_graph + set_inputs (:214-334) for a ragged batch of prompts: tokens[b][0 .. n_tokens[b])
 *                        (the caller appends EOS = eos_token_id, as t5_runner::generate does, :365-371) -> encodings = the prompts' [n_tokens[b]][output_size] rows
 *                        back to back; one prompt's rows are what b2tts_parler_set_text_encoding takes.  Errors: n_tokens outside 1 .. context_length, id >= vocab. */
typedef struct b2tts_t5 b2tts_t5;
int    b2tts_t5_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_t5 ** out);
void   b2tts_t5_free(b2tts_t5 * m);
int    b2tts_t5_info(const b2tts_t5 * m, int * n_layers, int * hidden_size, int * output_size, int * vocab_size, int * context_length, int * eos_token_id);
int    b2tts_t5_encode(b2tts_t5 * m, int n_prompts, const uint32_t * const * tokens, const int32_t * n_tokens, float * encodings);
float  b2tts_t5_last_ms(const b2tts_t5 * m);          /* device time of the last encode (CUDA events around the forward) */
int    b2tts_t5_last_used_gemm(const b2tts_t5 * m);   /* 1: the last encode had more than 32 rows and sent its F16 matrices through the tensor-core GEMM (B2TTS_T5_GEMM=0 disables) */
size_t b2tts_t5_weight_bytes(const b2tts_t5 * m);

typedef struct b2tts_parler b2tts_parler;
int   b2tts_parler_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_parler ** out);
void  b2tts_parler_free(b2tts_parler * m);
int   b2tts_parler_info(const b2tts_parler * m, int * n_heads, int * out_vocab, int * n_layers, int * hidden_size);
int   b2tts_parler_generate_greedy(b2tts_parler * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps,
                                   int32_t * out_tokens, float * out_logits);
/* n_generated != NULL turns on the reference's stop rule (eos_seen feeding + check_stopping, model.cpp:715-732,795-832): n_generated[b] = frames before the
 * reference's loop would have ended (position >= max_generation, or every head has produced EOS), rows past it are zero.  NULL: fixed-length generation. */
int   b2tts_parler_generate(b2tts_parler * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const b2tts_sampling * sampling,
                            int32_t * out_tokens, float * out_logits, int32_t * n_generated);
/* parler_tts_runner::update_conditional_prompt's second half (model.cpp:510-518 -> prep_cross_key_values, model.cpp:110-173): replace the stored conditional-prompt
 * encoding by `encoding` [n_rows][hidden] (the T5 encoder's output for a new voice description: b2tts_t5_encode above, or any other caller-side encoder)
 * and recompute the cross-attention K / V of every layer on the device. */
int   b2tts_parler_set_text_encoding(b2tts_parler * m, const float * encoding, int n_rows);
/* parity helper: greedy, fixed length, but the tokens fed back through the delay pattern are `teacher` ([n_sequences][n_steps][n_heads], e.g. the reference's own
 * output) instead of the produced ones; out_tokens / out_logits are still what this library produced at every step.  Block-quantised models (Q4_0 / Q5_0 / Q8_0
 * matrices: activations re-quantised per 32 columns) flip near-tied tokens on summation-order noise and then diverge; teacher forcing keeps every step comparable. */
int   b2tts_parler_generate_teacher_forced(b2tts_parler * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps,
                                           const int32_t * teacher, int32_t * out_tokens, float * out_logits);
float b2tts_parler_last_ms(const b2tts_parler * m);
size_t b2tts_parler_weight_bytes(const b2tts_parler * m);   /* bytes of the matrices, tables and norms resident in HBM (F16 matrices count 2 bytes) */
/* W_step of SURVEY 8(d): bytes of every weight tensor ONE decode step touches, each once, in its stored dtype (decoder matrices, norms, the stored cross K / V,
 * the output heads; not the embedding tables, of which a step reads nine rows) -- the weight term of the decode-step roofline */
size_t b2tts_parler_step_weight_bytes(const b2tts_parler * m);
/* how many launches of the persistent decode kernel (csrc/pdk.cuh) this model has issued and how many decode steps they covered; 0 / 0 = every step so far took the
 * launch-per-op path (sampling, > 16 sequences, F32 or block-quantised matrices, B2TTS_AR_PDK=0) */
void  b2tts_parler_pdk_stats(const b2tts_parler * m, uint64_t * launches, uint64_t * steps);

/* ------------------------------------------------------------------------------------------------------------------
 * Dia autoregressive decode.  The encoder pass runs launch per op; for F16 GGUFs (greedy / teacher-forced, <= 8 utterances) the whole CFG decoder loop
 * runs inside the PERSISTENT DECODE KERNEL (csrc/pdk.cuh: delay pattern + check_stopping in the rows phase, RoPE'd self / cross queries, fp32 GQA pages, cross-attention
 * over each row's encoding, cfg_scale + argmax phase; Dia-1.6B shape on a B200: 4.4 ms per step against 9.6 on the launch-per-op path); other dtypes / sampling use the
 * launch-per-op path (tensor-core GEMV for F16 matrices, CUDA-graph replay).  Hardware status (B200,
 * tests/test_dia_gpu.py, all green): F32 -- the reference's token ids exactly, CFG-combined logits 3.5e-3 at a logit std of 13, check_stopping's 63-frame run; F16
 * (BASELINE config 4's dtype) -- teacher-forced logits within 0.15 RMS and identical tokens wherever the reference's top-2 gap exceeds twice the step's logit
 * difference (the one free-running difference on a B200 sits on a 0.099 gap: rounding-boundary noise x the CFG gain, see the test's header); Q8_0 teacher-forced.
 *   b2tts_dia_load_gguf      : dia_model::setup_from_file + assign_weight loop over "dia.*" (reference src/models/dia/model.cpp:3-132,200-262;
 *                              dia/loader.cpp:8-22)
 *   b2tts_dia_generate_greedy: dia_runner::decode -- the encoder pass over the conditional and the all-zero unconditional sequence, the cross K/V
 *                              store, then the CFG-paired decoder step with cfg_scale (model.cpp:324-637,705-737; src/util.cpp:175-200) -- inside
 *                              generate_from_batch's loop with check_stopping (model.cpp:806-864; sampler::max per output head), for n_sequences
 *                              independent prompts of byte tokens, at most n_steps frames each.  n_generated[b] (may be NULL) is where
 *                              check_stopping ended utterance b; rows past it are zero.
 *                              out_tokens [n_sequences][n_steps][n_heads]; out_logits (may be NULL) [n_sequences][n_steps][n_heads][out_vocab]. */
typedef struct b2tts_dia b2tts_dia;
int   b2tts_dia_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_dia ** out);
void  b2tts_dia_free(b2tts_dia * m);
int   b2tts_dia_info(const b2tts_dia * m, int * n_heads, int * out_vocab, int * encoder_context, int * max_generation);
int   b2tts_dia_generate_greedy(b2tts_dia * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps,
                                int32_t * out_tokens, float * out_logits, int32_t * n_generated);
int   b2tts_dia_generate(b2tts_dia * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const b2tts_sampling * sampling,
                         int32_t * out_tokens, float * out_logits, int32_t * n_generated);
/* parity helper, as b2tts_parler_generate_teacher_forced */
int   b2tts_dia_generate_teacher_forced(b2tts_dia * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps,
                                        const int32_t * teacher, int32_t * out_tokens, float * out_logits);
/* generation_configuration::max_tokens (dia_runner::generate, model.cpp:873-879): replaces the model's max_generation_size in check_stopping when > max_delay */
int   b2tts_dia_set_max_generation(b2tts_dia * m, int max_tokens);
/* F16 GGUFs, greedy / teacher-forced, <= 8 utterances, decoder width <= 2 048: the decoder loop runs inside the persistent decode kernel (csrc/pdk.cuh;
 * B2TTS_AR_PDK=0: launch per op); -> cooperative launches so far and the decode steps they covered */
void  b2tts_dia_pdk_stats(const b2tts_dia * m, uint64_t * launches, uint64_t * steps);
size_t b2tts_dia_weight_bytes(const b2tts_dia * m);
float b2tts_dia_last_ms(const b2tts_dia * m);

/* ------------------------------------------------------------------------------------------------------------------
 * Piecewise weight hand-off for the codec and autoregressive models, the same three steps as b2tts_kokoro_{create,assign_weight,prepare}: what the reference's
 * runner_from_file drives through tts_model_loader::from_file (uint32 metadata), tts_generation_runner::assign_weight (every tensor, full GGUF name, ggml type
 * 0 = F32 / 1 = F16, ggml dims) and prepare_post_load (reference src/models/loaders.cpp:79-89).  A Parler or Dia runner hands "audio_encoder.*" tensors to its
 * b2tts_dac and the rest to the decoder; an Orpheus runner "snac.*" to its b2tts_snac (parler/model.cpp:271-318, dia/model.cpp:3-132, orpheus/model.cpp:436-447). */
int   b2tts_dac_create(b2tts_ctx * ctx, int n_kv, const char * const * kv_keys, const uint32_t * kv_vals, b2tts_dac ** out);
int   b2tts_dac_assign_weight(b2tts_dac * m, const char * name, int ggml_type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
int   b2tts_dac_prepare(b2tts_dac * m);
int   b2tts_snac_create(b2tts_ctx * ctx, int n_kv, const char * const * kv_keys, const uint32_t * kv_vals, b2tts_snac ** out);
int   b2tts_snac_assign_weight(b2tts_snac * m, const char * name, int ggml_type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
int   b2tts_snac_prepare(b2tts_snac * m);
int   b2tts_orpheus_create(b2tts_ctx * ctx, int n_kv, const char * const * kv_keys, const uint32_t * kv_vals, b2tts_orpheus ** out);
int   b2tts_orpheus_assign_weight(b2tts_orpheus * m, const char * name, int ggml_type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
int   b2tts_orpheus_prepare(b2tts_orpheus * m);
int   b2tts_parler_create(b2tts_ctx * ctx, int n_kv, const char * const * kv_keys, const uint32_t * kv_vals, b2tts_parler ** out);
int   b2tts_parler_assign_weight(b2tts_parler * m, const char * name, int ggml_type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
int   b2tts_parler_prepare(b2tts_parler * m);
int   b2tts_dia_create(b2tts_ctx * ctx, int n_kv, const char * const * kv_keys, const uint32_t * kv_vals, b2tts_dia ** out);
int   b2tts_dia_assign_weight(b2tts_dia * m, const char * name, int ggml_type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
int   b2tts_dia_prepare(b2tts_dia * m);

#ifdef __cplusplus
}
#endif
#endif /* B2TTS_H */
