// capi.cu -- extern "C" boundary of libb2tts.so (declared in include/b2tts.h).
#include "../../include/b2tts.h"
#include "kokoro.h"
#include "dac.h"
#include "orpheus.h"
#include "parler.h"
#include "t5.h"
#include "dia.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace b2 {
static thread_local char g_err[1024] = "";
void set_error(const char * fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace b2

using namespace b2;

struct b2tts_ctx { Ctx c; };
struct b2tts_kokoro { Kokoro k; };
struct b2tts_dac { Dac d; };
struct b2tts_snac { Snac s; };
struct b2tts_orpheus { Orpheus o; };
struct b2tts_parler { Parler p; };
struct b2tts_t5 { T5 t; };
struct b2tts_dia { Dia d; };

namespace {
// RAII device scratch for the op-level entry points
struct Dev {
    std::vector<void *> ptrs;
    ~Dev() { for (void * p : ptrs) cudaFree(p); }
    template <class T> T * get(size_t n) { void * p = nullptr; if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc failed"); return nullptr; } ptrs.push_back(p); return (T *) p; }
    // blocking copy on the legacy stream, then a device-wide sync: the kernels that read it run on ctx->stream (non-blocking), which does not order itself after the legacy stream
    template <class T> T * put(const T * h, size_t n) { T * d = get<T>(n); if (d && n) { cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice); cudaDeviceSynchronize(); } return d; }
};
int finish(Ctx * c, void * dst, const void * src, size_t bytes) {
    B2_CUDA(cudaStreamSynchronize(c->stream));
    B2_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
}  // namespace

extern "C" {

int b2tts_ctx_create(int device, b2tts_ctx ** out) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); set_error("no CUDA device visible: libb2tts has no CPU fallback"); return 1; }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return 1; }
    B2_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor); return 1; }
    b2tts_ctx * c = new b2tts_ctx();
    c->c.device = device;
    B2_CUDA(cudaStreamCreateWithFlags(&c->c.stream, cudaStreamNonBlocking));
    *out = c;
    return 0;
}
void b2tts_ctx_destroy(b2tts_ctx * ctx) {
    if (!ctx) return;
    if (ctx->c.stream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
}
const char * b2tts_last_error(void) { return g_err; }
uint64_t b2tts_launch_count(const b2tts_ctx * ctx) { return ctx ? ctx->c.launches : 0; }
void * b2tts_stream(const b2tts_ctx * ctx) { return ctx ? (void *) ctx->c.stream : nullptr; }
uint64_t b2tts_gemm_launches(const b2tts_ctx * ctx, int which) { return which == 0 ? ctx->c.umma_launches : ctx->c.mma_sync_launches; }
int b2tts_prof_enable(b2tts_ctx * ctx, int on) {
    Ctx & c = ctx->c;
    for (auto & r : c.recs) { c.pool.push_back(r.a); c.pool.push_back(r.b); }
    c.recs.clear();
    c.prof = on != 0;
    return 0;
}
int b2tts_prof_read(b2tts_ctx * ctx, int kind, double * total_ms, double * flops, double * bytes, uint64_t * launches) {
    Ctx & c = ctx->c;
    B2_CUDA(cudaStreamSynchronize(c.stream));
    double ms = 0, fl = 0, by = 0; uint64_t n = 0;
    const char * dump = getenv("B2TTS_PROF_DUMP");   // diagnostics: append one line per profiled launch of this kind
    FILE * df = dump ? fopen(dump, "a") : nullptr;
    for (auto & r : c.recs) {
        if (r.kind != kind) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
        ms += t; fl += r.flops; by += r.bytes; n++;
        if (df) fprintf(df, "%d %.4f %.4g %.4g %s\n", r.kind, t, r.flops, r.bytes, r.tag);
    }
    if (df) fclose(df);
    if (total_ms) *total_ms = ms; if (flops) *flops = fl; if (bytes) *bytes = by; if (launches) *launches = n;
    return 0;
}

int b2tts_kokoro_create(b2tts_ctx * ctx, int n_kv, const char * const * kv_keys, const uint32_t * kv_vals, b2tts_kokoro ** out) {
    if (!ctx) { set_error("null context"); return 1; }
    b2tts_kokoro * m = new b2tts_kokoro();
    m->k.ctx = &ctx->c;
    for (int i = 0; i < n_kv; i++) m->k.kv[kv_keys[i]] = kv_vals[i];
    *out = m;
    return 0;
}
int b2tts_kokoro_assign_weight(b2tts_kokoro * m, const char * name, int ggml_type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {
    return m->k.assign(name, ggml_type, n_dims, ne, data, nbytes);
}
int b2tts_kokoro_prepare(b2tts_kokoro * m) { B2_CUDA(cudaSetDevice(m->k.ctx->device)); return m->k.prepare(); }
int b2tts_kokoro_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_kokoro ** out) {
    if (!ctx) { set_error("null context"); return 1; }
    B2_CUDA(cudaSetDevice(ctx->c.device));
    b2tts_kokoro * m = new b2tts_kokoro();
    m->k.ctx = &ctx->c;
    if (load_gguf_into(&m->k, path)) { m->k.free_all(); delete m; return 1; }
    *out = m;
    return 0;
}
void b2tts_kokoro_free(b2tts_kokoro * m) { if (m) { m->k.free_all(); delete m; } }

// ---- piecewise weight hand-off for the other models: what runner_from_file drives through tts_model_loader::from_file (metadata), assign_weight (every
// tensor of the GGUF; names outside the model's prefix are ignored by the caller) and prepare_post_load (reference src/models/loaders.cpp:79-89)
#define B2TTS_HANDOFF(NAME, FIELD)                                                                                                                                   \
    int b2tts_##NAME##_create(b2tts_ctx * ctx, int n_kv, const char * const * kv_keys, const uint32_t * kv_vals, b2tts_##NAME ** out) {                               \
        if (!ctx) { set_error("null context"); return 1; }                                                                                                            \
        b2tts_##NAME * m = new b2tts_##NAME();                                                                                                                        \
        m->FIELD.ctx = &ctx->c;                                                                                                                                       \
        for (int i = 0; i < n_kv; i++) m->FIELD.kv[kv_keys[i]] = kv_vals[i];                                                                                          \
        *out = m;                                                                                                                                                     \
        return 0;                                                                                                                                                     \
    }                                                                                                                                                                 \
    int b2tts_##NAME##_assign_weight(b2tts_##NAME * m, const char * name, int ggml_type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {          \
        if (!m) { set_error("null model"); return 1; }                                                                                                                \
        return m->FIELD.assign(name, ggml_type, n_dims, ne, data, nbytes);                                                                                            \
    }                                                                                                                                                                 \
    int b2tts_##NAME##_prepare(b2tts_##NAME * m) { if (!m) { set_error("null model"); return 1; } B2_CUDA(cudaSetDevice(m->FIELD.ctx->device)); return m->FIELD.prepare(); }
B2TTS_HANDOFF(dac, d)
B2TTS_HANDOFF(snac, s)
B2TTS_HANDOFF(orpheus, o)
B2TTS_HANDOFF(parler, p)
B2TTS_HANDOFF(dia, d)
#undef B2TTS_HANDOFF

// ---- DAC codec decoder
int b2tts_dac_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_dac ** out) {
    if (!ctx) { set_error("null context"); return 1; }
    B2_CUDA(cudaSetDevice(ctx->c.device));
    b2tts_dac * m = new b2tts_dac();
    m->d.ctx = &ctx->c;
    if (load_gguf_into(&m->d, path)) { m->d.free_all(); delete m; return 1; }
    *out = m;
    return 0;
}
void b2tts_dac_free(b2tts_dac * m) { if (m) { m->d.free_all(); delete m; } }
int b2tts_dac_info(const b2tts_dac * m, int * n_heads, int * up_sampling_factor, int * codebook_size) {
    if (!m) { set_error("null model"); return 1; }
    if (n_heads) *n_heads = m->d.n_heads;
    if (up_sampling_factor) *up_sampling_factor = m->d.up_factor;
    if (codebook_size) *codebook_size = m->d.n_codes;
    return 0;
}
int b2tts_dac_decode_batch(b2tts_dac * m, int n_utterances, const uint32_t * const * codes, const int32_t * frames, const float ** pcm, int64_t * n_samples) {
    if (!m) { set_error("null model"); return 1; }
    return m->d.decode_batch(n_utterances, codes, frames, pcm, n_samples);
}
float b2tts_dac_last_ms(const b2tts_dac * m) { return m ? m->d.timing_ms : 0.f; }

// ---- SNAC codec decoder
int b2tts_snac_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_snac ** out) {
    if (!ctx) { set_error("null context"); return 1; }
    B2_CUDA(cudaSetDevice(ctx->c.device));
    b2tts_snac * m = new b2tts_snac();
    m->s.ctx = &ctx->c;
    if (load_gguf_into(&m->s, path)) { m->s.free_all(); delete m; return 1; }
    *out = m;
    return 0;
}
void b2tts_snac_free(b2tts_snac * m) { if (m) { m->s.free_all(); delete m; } }
int b2tts_snac_info(const b2tts_snac * m, int * up_sampling_factor, int * codebook_size) {
    if (!m) { set_error("null model"); return 1; }
    if (up_sampling_factor) *up_sampling_factor = m->s.up_factor;
    if (codebook_size) *codebook_size = m->s.n_codes;
    return 0;
}
int b2tts_snac_decode_batch(b2tts_snac * m, int n_utterances, const uint32_t * const * codes, const int32_t * fine_frames, const float ** pcm, int64_t * n_samples) {
    if (!m) { set_error("null model"); return 1; }
    return m->s.decode_batch(n_utterances, codes, fine_frames, pcm, n_samples);
}
float b2tts_snac_last_ms(const b2tts_snac * m) { return m ? m->s.timing_ms : 0.f; }
// ---- Orpheus AR decode (first correct path)
int b2tts_orpheus_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_orpheus ** out) {
    if (!ctx) { set_error("null context"); return 1; }
    B2_CUDA(cudaSetDevice(ctx->c.device));
    b2tts_orpheus * m = new b2tts_orpheus();
    m->o.ctx = &ctx->c;
    if (load_gguf_into(&m->o, path)) { m->o.free_all(); delete m; return 1; }
    *out = m;
    return 0;
}
void b2tts_orpheus_free(b2tts_orpheus * m) { if (m) { m->o.free_all(); delete m; } }
int b2tts_orpheus_info(const b2tts_orpheus * m, int * vocab_size, int * n_layers, int * hidden_size) {
    if (!m) { set_error("null model"); return 1; }
    if (vocab_size) *vocab_size = m->o.vocab;
    if (n_layers) *n_layers = m->o.n_layers;
    if (hidden_size) *hidden_size = m->o.hidden;
    return 0;
}
int b2tts_orpheus_generate_greedy(b2tts_orpheus * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, int32_t * out_tokens,
                                  float * out_logits) {
    if (!m) { set_error("null model"); return 1; }
    return m->o.generate_greedy(n_sequences, prompts, n_prompt, n_steps, out_tokens, out_logits);
}
static ArSampling to_sampling(const b2tts_sampling * s) {
    ArSampling a;
    if (s) { a.do_sample = s->do_sample; a.top_k = s->top_k; a.top_p = s->top_p; a.temperature = s->temperature; a.repetition_penalty = s->repetition_penalty; a.seed = s->seed; }
    return a;
}
int b2tts_orpheus_generate(b2tts_orpheus * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const b2tts_sampling * sampling,
                           int32_t * out_tokens, float * out_logits) {
    if (!m) { set_error("null model"); return 1; }
    const ArSampling a = to_sampling(sampling);
    return m->o.generate(n_sequences, prompts, n_prompt, n_steps, &a, out_tokens, out_logits);
}
int b2tts_orpheus_generate_until_stop(b2tts_orpheus * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int max_steps, const b2tts_sampling * sampling,
                                      int32_t * out_tokens, int32_t * n_generated) {
    if (!m) { set_error("null model"); return 1; }
    if (!n_generated) { set_error("n_generated must not be null"); return 1; }
    const ArSampling a = to_sampling(sampling);
    return m->o.generate(n_sequences, prompts, n_prompt, max_steps, &a, out_tokens, nullptr, n_generated);
}
int b2tts_orpheus_set_stopping_token(b2tts_orpheus * m, int token_id) { if (!m) { set_error("null model"); return 1; } m->o.stopping_token = token_id; return 0; }
void b2tts_orpheus_pdk_stats(const b2tts_orpheus * m, uint64_t * launches, uint64_t * steps) { if (launches) *launches = m ? m->o.pdk_launches : 0; if (steps) *steps = m ? m->o.pdk_steps : 0; }
size_t b2tts_orpheus_step_weight_bytes(const b2tts_orpheus * m) {
    if (!m) return 0;
    const Orpheus & o = m->o;
    auto wb = [](const ArW & w, size_t n) { return w.qtype ? n * (w.qtype == 2 ? 18 : w.qtype == 6 ? 22 : 34) / 32 : n * (w.f16 ? 2 : 4); };
    const size_t H = (size_t) o.hidden, KV = (size_t) o.kv_hidden, F = (size_t) o.ffn;
    size_t b = 0;
    for (const OrpheusLayer & L : o.layers) b += wb(L.wq, H * H) + wb(L.wk, KV * H) + wb(L.wv, KV * H) + wb(L.wo, H * H) + wb(L.wgate, F * H) + wb(L.wup, F * H) + wb(L.wdown, H * F) + 2 * H * 4;
    return b + wb(o.head, (size_t) o.vocab * H) + H * 4;
}
size_t b2tts_orpheus_weight_bytes(const b2tts_orpheus * m) { return m ? m->o.weight_bytes : 0; }
float b2tts_orpheus_last_ms(const b2tts_orpheus * m) { return m ? m->o.timing_ms : 0.f; }
// ---- T5 conditional-prompt encoder (the pass before Parler's decode loop)
int b2tts_t5_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_t5 ** out) {
    if (!ctx) { set_error("null context"); return 1; }
    B2_CUDA(cudaSetDevice(ctx->c.device));
    b2tts_t5 * m = new b2tts_t5();
    m->t.ctx = &ctx->c;
    if (load_gguf_into(&m->t, path)) { m->t.free_all(); delete m; return 1; }
    *out = m;
    return 0;
}
void b2tts_t5_free(b2tts_t5 * m) { if (m) { m->t.free_all(); delete m; } }
int b2tts_t5_info(const b2tts_t5 * m, int * n_layers, int * hidden_size, int * output_size, int * vocab_size, int * context_length, int * eos_token_id) {
    if (!m) { set_error("null model"); return 1; }
    if (n_layers) *n_layers = m->t.n_layers;
    if (hidden_size) *hidden_size = m->t.hidden;
    if (output_size) *output_size = m->t.output_size();
    if (vocab_size) *vocab_size = m->t.vocab;
    if (context_length) *context_length = m->t.max_ctx;
    if (eos_token_id) *eos_token_id = m->t.eos;
    return 0;
}
int b2tts_t5_encode(b2tts_t5 * m, int n_prompts, const uint32_t * const * tokens, const int32_t * n_tokens, float * encodings) {
    if (!m) { set_error("null model"); return 1; }
    if (n_prompts > 0 && (!tokens || !n_tokens || !encodings)) { set_error("t5: null argument"); return 1; }
    for (int b = 0; b < n_prompts; b++) if (!tokens[b]) { set_error("t5: prompt %d is a null pointer", b); return 1; }
    return m->t.encode(n_prompts, tokens, n_tokens, encodings);
}
float b2tts_t5_last_ms(const b2tts_t5 * m) { return m ? m->t.timing_ms : 0.f; }
int b2tts_t5_last_used_gemm(const b2tts_t5 * m) { return m && m->t.last_used_gemm ? 1 : 0; }
size_t b2tts_t5_weight_bytes(const b2tts_t5 * m) { return m ? m->t.weight_bytes : 0; }
// ---- Parler AR decode (first correct path)
int b2tts_parler_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_parler ** out) {
    if (!ctx) { set_error("null context"); return 1; }
    B2_CUDA(cudaSetDevice(ctx->c.device));
    b2tts_parler * m = new b2tts_parler();
    m->p.ctx = &ctx->c;
    if (load_gguf_into(&m->p, path)) { m->p.free_all(); delete m; return 1; }
    *out = m;
    return 0;
}
void b2tts_parler_free(b2tts_parler * m) { if (m) { m->p.free_all(); delete m; } }
int b2tts_parler_info(const b2tts_parler * m, int * n_heads, int * out_vocab, int * n_layers, int * hidden_size) {
    if (!m) { set_error("null model"); return 1; }
    if (n_heads) *n_heads = m->p.n_out;
    if (out_vocab) *out_vocab = m->p.vocab;
    if (n_layers) *n_layers = m->p.n_layers;
    if (hidden_size) *hidden_size = m->p.hidden;
    return 0;
}
int b2tts_parler_generate_greedy(b2tts_parler * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, int32_t * out_tokens,
                                 float * out_logits) {
    if (!m) { set_error("null model"); return 1; }
    return m->p.generate_greedy(n_sequences, prompts, n_prompt, n_steps, out_tokens, out_logits);
}
int b2tts_parler_generate(b2tts_parler * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const b2tts_sampling * sampling,
                          int32_t * out_tokens, float * out_logits, int32_t * n_generated) {
    if (!m) { set_error("null model"); return 1; }
    const ArSampling a = to_sampling(sampling);
    return m->p.generate(n_sequences, prompts, n_prompt, n_steps, &a, out_tokens, out_logits, n_generated);
}
int b2tts_parler_generate_teacher_forced(b2tts_parler * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const int32_t * teacher,
                                         int32_t * out_tokens, float * out_logits) {
    if (!m) { set_error("null model"); return 1; }
    if (!teacher) { set_error("null teacher tokens"); return 1; }
    return m->p.generate(n_sequences, prompts, n_prompt, n_steps, nullptr, out_tokens, out_logits, nullptr, teacher);
}
int b2tts_parler_set_text_encoding(b2tts_parler * m, const float * encoding, int n_rows) {
    if (!m) { set_error("null model"); return 1; }
    return m->p.set_text_encoding(encoding, n_rows);
}
float b2tts_parler_last_ms(const b2tts_parler * m) { return m ? m->p.timing_ms : 0.f; }
size_t b2tts_parler_weight_bytes(const b2tts_parler * m) { return m ? m->p.weight_bytes : 0; }
size_t b2tts_parler_step_weight_bytes(const b2tts_parler * m) {
    if (!m) return 0;
    const Parler & p = m->p;
    auto wb = [](const ArW & w, size_t n) { return w.qtype ? n * (w.qtype == 2 ? 18 : w.qtype == 6 ? 22 : 34) / 32 : n * (w.f16 ? 2 : 4); };
    const size_t H = (size_t) p.hidden, F = (size_t) p.ffn;
    size_t b = 0;
    for (const ParlerLayer & L : p.layers)
        b += wb(L.wq, H * H) + wb(L.wk, H * H) + wb(L.wv, H * H) + wb(L.wo, H * H) + wb(L.cq, H * H) + wb(L.co, H * H) + wb(L.fc1, F * H) + wb(L.fc2, H * F) + 6 * H * 4 +
             2 * (size_t) p.n_enc * H * 4;                                   // + the layer's cross K / V store (fp32, shared by the batch)
    b += wb(p.heads_w, (size_t) p.n_out * p.vocab * H) + 2 * H * 4;
    return b;
}
void b2tts_parler_pdk_stats(const b2tts_parler * m, uint64_t * launches, uint64_t * steps) { if (launches) *launches = m ? m->p.pdk_launches : 0; if (steps) *steps = m ? m->p.pdk_steps : 0; }
// ---- Dia AR decode (first correct path)
int b2tts_dia_load_gguf(b2tts_ctx * ctx, const char * path, b2tts_dia ** out) {
    if (!ctx) { set_error("null context"); return 1; }
    B2_CUDA(cudaSetDevice(ctx->c.device));
    b2tts_dia * m = new b2tts_dia();
    m->d.ctx = &ctx->c;
    if (load_gguf_into(&m->d, path)) { m->d.free_all(); delete m; return 1; }
    *out = m;
    return 0;
}
void b2tts_dia_free(b2tts_dia * m) { if (m) { m->d.free_all(); delete m; } }
int b2tts_dia_info(const b2tts_dia * m, int * n_heads, int * out_vocab, int * encoder_context, int * max_generation) {
    if (!m) { set_error("null model"); return 1; }
    if (n_heads) *n_heads = m->d.n_out;
    if (out_vocab) *out_vocab = m->d.vocab;
    if (encoder_context) *encoder_context = m->d.enc_ctx;
    if (max_generation) *max_generation = m->d.max_gen;
    return 0;
}
int b2tts_dia_generate_greedy(b2tts_dia * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, int32_t * out_tokens,
                              float * out_logits, int32_t * n_generated) {
    if (!m) { set_error("null model"); return 1; }
    return m->d.generate_greedy(n_sequences, prompts, n_prompt, n_steps, out_tokens, out_logits, n_generated);
}
int b2tts_dia_generate(b2tts_dia * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const b2tts_sampling * sampling,
                       int32_t * out_tokens, float * out_logits, int32_t * n_generated) {
    if (!m) { set_error("null model"); return 1; }
    const ArSampling a = to_sampling(sampling);
    return m->d.generate(n_sequences, prompts, n_prompt, n_steps, &a, out_tokens, out_logits, n_generated);
}
int b2tts_dia_generate_teacher_forced(b2tts_dia * m, int n_sequences, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const int32_t * teacher,
                                      int32_t * out_tokens, float * out_logits) {
    if (!m) { set_error("null model"); return 1; }
    if (!teacher) { set_error("null teacher tokens"); return 1; }
    return m->d.generate(n_sequences, prompts, n_prompt, n_steps, nullptr, out_tokens, out_logits, nullptr, teacher);
}
int b2tts_dia_set_max_generation(b2tts_dia * m, int max_tokens) {
    if (!m) { set_error("null model"); return 1; }
    if (max_tokens > m->d.max_delay) m->d.max_gen = max_tokens;
    return 0;
}
void b2tts_dia_pdk_stats(const b2tts_dia * m, uint64_t * launches, uint64_t * steps) { if (launches) *launches = m ? m->d.pdk_launches : 0; if (steps) *steps = m ? m->d.pdk_steps : 0; }
size_t b2tts_dia_weight_bytes(const b2tts_dia * m) { return m ? m->d.weight_bytes : 0; }
float b2tts_dia_last_ms(const b2tts_dia * m) { return m ? m->d.timing_ms : 0.f; }

int b2tts_snac_reset_noise(b2tts_snac * m) { if (!m) { set_error("null model"); return 1; } m->s.reset_noise(); return 0; }
int b2tts_kokoro_n_voices(const b2tts_kokoro * m) { return (int) m->k.voice_names.size(); }
const char * b2tts_kokoro_voice_name(const b2tts_kokoro * m, int i) { return (i >= 0 && i < (int) m->k.voice_names.size()) ? m->k.voice_names[i].c_str() : nullptr; }
size_t b2tts_kokoro_weight_bytes(const b2tts_kokoro * m) { return m->k.weight_bytes; }

int b2tts_kokoro_run_batch(b2tts_kokoro * m, int batch, const uint32_t * tokens, const int32_t * n_tokens, const char * voice, const uint64_t * noise_skip,
                           const float ** pcm, int64_t * n_samples, const float ** durations) {
    B2_CUDA(cudaSetDevice(m->k.ctx->device));
    return m->k.run_batch(batch, tokens, n_tokens, voice, noise_skip, pcm, n_samples, durations);
}
int b2tts_kokoro_run_batch_device(b2tts_kokoro * m, int batch, const uint32_t * tokens, const int32_t * n_tokens, const char * voice, const uint64_t * noise_skip,
                                  const float ** pcm_device, int64_t * row_stride, int64_t * n_samples) {
    B2_CUDA(cudaSetDevice(m->k.ctx->device));
    m->k.keep_on_device = true;
    const int rc = m->k.run_batch(batch, tokens, n_tokens, voice, noise_skip, nullptr, n_samples, nullptr);
    m->k.keep_on_device = false;
    if (!rc) { if (pcm_device) *pcm_device = m->k.last_pcm_dev; if (row_stride) *row_stride = m->k.last_pcm_stride; }
    return rc;
}
int b2tts_kokoro_run_chunks(b2tts_kokoro * m, int batch, const uint32_t * tokens, const int32_t * n_tokens, const char * voice, uint64_t noise_skip_first,
                            const float ** pcm, int64_t * n_samples, const float ** durations) {
    B2_CUDA(cudaSetDevice(m->k.ctx->device));
    m->k.chain_noise = true; m->k.chain_noise_start = noise_skip_first;
    const int rc = m->k.run_batch(batch, tokens, n_tokens, voice, nullptr, pcm, n_samples, durations);
    m->k.chain_noise = false;
    return rc;
}
int b2tts_kokoro_last_timings(const b2tts_kokoro * m, float ms[3]) { for (int i = 0; i < 3; i++) ms[i] = m->k.timings[i]; return 0; }

int b2tts_kokoro_set_taps(b2tts_kokoro * m, int enable) { m->k.taps_on = enable != 0; return 0; }
int b2tts_kokoro_tap_info(b2tts_kokoro * m, const char * name, int64_t * rows, int64_t * cols, int64_t * padded_len) {
    auto it = m->k.taps.find(name);
    if (it == m->k.taps.end()) { set_error("no tap named '%s' (enable taps and run first)", name); return 1; }
    if (rows) *rows = it->second.rows; if (cols) *cols = it->second.cols; if (padded_len) *padded_len = it->second.padded;
    return 0;
}
int b2tts_kokoro_tap_read(b2tts_kokoro * m, const char * name, float * dst, size_t count) {
    auto it = m->k.taps.find(name);
    if (it == m->k.taps.end()) { set_error("no tap named '%s'", name); return 1; }
    const Tap & t = it->second;
    if ((int64_t) count != t.rows * t.cols) { set_error("tap '%s' holds %lld floats, asked for %zu", name, (long long) (t.rows * t.cols), count); return 1; }
    B2_CUDA(cudaSetDevice(m->k.ctx->device));
    B2_CUDA(cudaMemcpy2D(dst, t.cols * 4, t.ptr, t.ld * 4, t.cols * 4, t.rows, cudaMemcpyDeviceToHost));
    return 0;
}
int b2tts_kokoro_override(b2tts_kokoro * m, const char * name, const float * src, size_t count) {
    if (count == 0) { m->k.overrides.erase(name); return 0; }
    m->k.overrides[name] = std::vector<float>(src, src + count);
    return 0;
}

// ------------------------------------------------------------------------------------------ op-level entry points
int b2tts_op_conv_transpose_1d(b2tts_ctx * ctx, const float * kernel, int K, int coutg, int cin, const float * x, int L, int stride, int pad, int out_pad,
                               int groups, float * y) {
    Ctx * c = &ctx->c; Dev d;
    const int Lout = (L - 1) * stride - 2 * pad + (K - 1) + out_pad + 1;
    float * dk = d.put(kernel, (size_t) K * coutg * cin); float * dx = d.put(x, (size_t) L * cin); float * dy = d.get<float>((size_t) Lout * coutg * groups);
    if (!dk || !dx || !dy) return 1;
    if (op_conv_transpose_1d(c, dk, K, coutg, cin, dx, L, stride, pad, out_pad, groups, dy, Lout)) return 1;
    return finish(c, y, dy, (size_t) Lout * coutg * groups * 4);
}

int b2tts_op_conv_1d(b2tts_ctx * ctx, const float * kernel, int K, int cin, int cout, const float * x, int L, int stride, int pad, int dil, int f16_kernel,
                     float * y) {
    if (!f16_kernel) { set_error("b2tts_op_conv_1d: only the F16-kernel path (fp16 operands, fp32 accumulate) is implemented"); return 1; }
    Ctx * c = &ctx->c; Dev d;
    const int Lout = (L + 2 * pad - dil * (K - 1) - 1) / stride + 1;
    const int cp = round_up(cin, 64), np = cout > 64 ? round_up(cout, 128) : 64;
    std::vector<__half> w((size_t) np * K * cp, __float2half(0.f)), a((size_t) L * cp, __float2half(0.f));
    for (int co = 0; co < cout; co++) for (int ci = 0; ci < cin; ci++) for (int k = 0; k < K; k++) w[((size_t) co * K + k) * cp + ci] = __float2half(kernel[((size_t) co * cin + ci) * K + k]);
    for (int ci = 0; ci < cin; ci++) for (int t = 0; t < L; t++) a[(size_t) t * cp + ci] = __float2half(x[(size_t) ci * L + t]);
    __half * dw = d.put(w.data(), w.size()); __half * da = d.put(a.data(), a.size()); float * dy = d.get<float>((size_t) Lout * cout);
    if (!dw || !da || !dy) return 1;
    ConvGemmParams p;
    p.A = da; p.lda = cp; p.W = dw; p.outF = dy; p.ldo = cout; p.B = 1; p.LmaxIn = L; p.LmaxOut = Lout; p.N = cout; p.Npad = np; p.KW = K; p.CinPad = cp;
    p.stride = stride; p.dil = dil; p.pad = pad;
    if (conv_gemm(c, p)) return 1;
    std::vector<float> t((size_t) Lout * cout);
    if (finish(c, t.data(), dy, t.size() * 4)) return 1;
    for (int co = 0; co < cout; co++) for (int o = 0; o < Lout; o++) y[(size_t) co * Lout + o] = t[(size_t) o * cout + co];
    return 0;
}

float b2tts_sample_uniform(uint64_t seed, uint64_t row, uint64_t step) { return sample_uniform_host(seed, row, step); }
int b2tts_op_sample(b2tts_ctx * ctx, const float * logits, int rows, int vocab, int do_sample, int top_k, float top_p, float temperature, float repetition_penalty,
                    int32_t * last_ids, int32_t * rep_counts, uint64_t seed, int step, int32_t * tokens) {
    Ctx * c = &ctx->c; Dev d;
    SampleParams p;
    p.rows = rows; p.V = vocab; p.do_sample = do_sample; p.top_k = top_k; p.top_p = top_p; p.temperature = temperature; p.repetition_penalty = repetition_penalty; p.seed = seed;
    p.logits = d.put(logits, (size_t) rows * vocab);
    p.scratch = d.get<float>((size_t) rows * vocab);
    std::vector<int> hstep(1, step);
    int * dstep = d.put(hstep.data(), 1);
    p.d_step = dstep;
    p.out = d.get<int>((size_t) rows * (step + 1));
    if (last_ids && rep_counts) { p.last_ids = d.put((const int *) last_ids, (size_t) rows); p.rep_counts = d.put((const int *) rep_counts, (size_t) rows); }
    if (!p.logits || !p.scratch || !dstep || !p.out || sample_rows(c, p)) return 1;
    if (finish(c, tokens, p.out + (size_t) step * rows, (size_t) rows * 4)) return 1;
    if (last_ids && rep_counts) { if (finish(c, last_ids, p.last_ids, (size_t) rows * 4) || finish(c, rep_counts, p.rep_counts, (size_t) rows * 4)) return 1; }
    return 0;
}

int b2tts_op_cumsum(b2tts_ctx * ctx, const float * x, int L, int rows, float * y) {
    Ctx * c = &ctx->c; Dev d; float * dx = d.put(x, (size_t) L * rows); float * dy = d.get<float>((size_t) L * rows);
    if (!dx || !dy || op_cumsum(c, dx, L, rows, dy)) return 1;
    return finish(c, y, dy, (size_t) L * rows * 4);
}
int b2tts_op_vad_trim(b2tts_ctx * ctx, const float * pcm, const int64_t * n_samples, int B, float sample_rate, int ms_per_frame, int frame_threshold,
                      float normalized_energy_threshold, int trailing_silent_frames, int early_cutoff_seconds_threshold, float early_cutoff_energy_threshold,
                      int64_t * n_out, float * energies_out) {
    if (!ctx) { set_error("null context"); return 1; }
    Ctx * c = &ctx->c;
    if (B <= 0) return 0;
    if (!pcm || !n_samples || !n_out) { set_error("vad: null argument"); return 1; }
    if (ms_per_frame <= 0) { set_error("vad: ms_per_frame must be positive (the reference divides by it)"); return 1; }
    const int spf = (int) (ms_per_frame * sample_rate / 1000.0f);                                  // vad.cpp:20
    if (spf <= 0) { set_error("vad: ms_per_frame * sample_rate / 1000 < 1 (the reference divides by zero here)"); return 1; }
    const int early_frames = (int) ((early_cutoff_seconds_threshold * 1000) / ms_per_frame);       // vad.cpp:22
    std::vector<long long> off((size_t) B + 1, 0), eoff((size_t) B + 1, 0);
    int max_frames = 0;
    for (int b = 0; b < B; b++) {
        if (n_samples[b] < 0) { set_error("vad: utterance %d has a negative length", b); return 1; }
        const long long nf = n_samples[b] / spf;
        if (nf > 0x7fffffff) { set_error("vad: utterance %d has more frames than the reference's int can count", b); return 1; }
        off[(size_t) b + 1] = off[(size_t) b] + n_samples[b]; eoff[(size_t) b + 1] = eoff[(size_t) b] + nf;
        max_frames = std::max(max_frames, (int) nf);
    }
    Dev d;
    float * dp = d.put(pcm, (size_t) off[(size_t) B]); long long * doff = d.put(off.data(), off.size()), * deoff = d.put(eoff.data(), eoff.size());
    float * de = d.get<float>((size_t) eoff[(size_t) B]); long long * dn = d.get<long long>((size_t) B);
    if (!dp || !doff || !deoff || !de || !dn) return 1;
    if (vad_trim_rows(c, dp, doff, deoff, B, max_frames, spf, frame_threshold, normalized_energy_threshold, trailing_silent_frames, early_frames,
                      early_cutoff_energy_threshold, de, dn)) return 1;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64_t");
    if (finish(c, n_out, dn, (size_t) B * 8)) return 1;
    if (energies_out && eoff[(size_t) B]) B2_CUDA(cudaMemcpy(energies_out, de, (size_t) eoff[(size_t) B] * 4, cudaMemcpyDeviceToHost));
    return 0;
}
static int unary(b2tts_ctx * ctx, int which, const float * x, int64_t n, float arg, float * y) {
    Ctx * c = &ctx->c; Dev d; float * dx = d.put(x, (size_t) n); float * dy = d.get<float>((size_t) n);
    if (!dx || !dy || op_unary(c, which, dx, n, arg, dy)) return 1;
    return finish(c, y, dy, (size_t) n * 4);
}
int b2tts_op_mod(b2tts_ctx * ctx, const float * x, int64_t n, float mod_val, float * y) { return unary(ctx, 0, x, n, mod_val, y); }
int b2tts_op_round(b2tts_ctx * ctx, const float * x, int64_t n, float * y) { return unary(ctx, 1, x, n, 0.f, y); }
int b2tts_op_reciprocal(b2tts_ctx * ctx, const float * x, int64_t n, float * y) { return unary(ctx, 2, x, n, 0.f, y); }
int b2tts_op_upscale_linear(b2tts_ctx * ctx, const float * x, int L, int rows, int factor, float * y) {
    Ctx * c = &ctx->c; Dev d; float * dx = d.put(x, (size_t) L * rows); float * dy = d.get<float>((size_t) L * rows * factor);
    if (!dx || !dy || op_upscale_linear(c, dx, L, rows, factor, dy)) return 1;
    return finish(c, y, dy, (size_t) L * rows * factor * 4);
}
int b2tts_op_snake(b2tts_ctx * ctx, const float * alpha, int C, const float * x, int L, float * y) {
    Ctx * c = &ctx->c; Dev d; float * da = d.put(alpha, (size_t) C); float * dx = d.put(x, (size_t) C * L); float * dy = d.get<float>((size_t) C * L);
    if (!da || !dx || !dy || op_snake(c, da, C, dx, L, dy)) return 1;
    return finish(c, y, dy, (size_t) C * L * 4);
}
int b2tts_op_stft(b2tts_ctx * ctx, const float * x, int L, int n_fft, int hop, float * mag, float * phase) {
    if (n_fft != 20 || hop != 5) { set_error("b2tts_op_stft: only n_fft=20 hop=5 (Kokoro's iSTFTNet) is implemented"); return 1; }
    Ctx * c = &ctx->c; Dev d;
    const int frames = L / hop + 1, one = L;
    float * dx = d.put(x, (size_t) L); float * df = d.get<float>((size_t) frames * 22); int * dl = d.put(&one, 1);
    if (!dx || !df || !dl || stft20(c, dx, L, 1, dl, frames, nullptr, 0, 0, df, 22)) return 1;
    std::vector<float> t((size_t) frames * 22);
    if (finish(c, t.data(), df, t.size() * 4)) return 1;
    for (int f = 0; f < frames; f++) for (int k = 0; k < 11; k++) { mag[(size_t) f * 11 + k] = t[(size_t) f * 22 + k]; phase[(size_t) f * 11 + k] = t[(size_t) f * 22 + 11 + k]; }
    return 0;
}
int b2tts_op_istft(b2tts_ctx * ctx, const float * mag, const float * phase, int frames, int n_fft, int hop, float * y) {
    if (n_fft != 20 || hop != 5) { set_error("b2tts_op_istft: only n_fft=20 hop=5 is implemented"); return 1; }
    Ctx * c = &ctx->c; Dev d;
    std::vector<float> t((size_t) frames * 22);
    for (int f = 0; f < frames; f++) for (int k = 0; k < 11; k++) { t[(size_t) f * 22 + k] = mag[(size_t) f * 11 + k]; t[(size_t) f * 22 + 11 + k] = phase[(size_t) f * 11 + k]; }
    const int S = (frames - 1) * hop;
    float * ds = d.put(t.data(), t.size()); float * dy = d.get<float>((size_t) S); int * dl = d.put(&frames, 1);
    if (!ds || !dy || !dl || istft20(c, ds, 22, 1, dl, frames, dy, S)) return 1;
    return finish(c, y, dy, (size_t) S * 4);
}
int b2tts_op_uniform(b2tts_ctx * ctx, uint64_t skip, int64_t count, float * y) {
    Ctx * c = &ctx->c; Dev d; float * dy = d.get<float>((size_t) count);
    if (!dy || op_uniform(c, skip, count, dy)) return 1;
    return finish(c, y, dy, (size_t) count * 4);
}

int b2tts_op_bilstm(b2tts_ctx * ctx, const float * w_ih, const float * w_hh, const float * b_ih, const float * b_hh, int In, int H, const float * x, int B,
                    int Lmax, const int32_t * len, float * y) {
    if (H != 256) { set_error("b2tts_op_bilstm: hidden size must be 256"); return 1; }
    Ctx * c = &ctx->c; Dev d;
    const int ip = round_up(In, 64);
    std::vector<__half> wih((size_t) 2048 * ip, __float2half(0.f)), whh((size_t) 2 * 1024 * 256), xa((size_t) B * Lmax * ip, __float2half(0.f));
    std::vector<float> bih(2048), bhh(2048);
    for (int dd = 0; dd < 2; dd++) for (int g = 0; g < 4; g++) for (int u = 0; u < H; u++) {
        const size_t src = (size_t) dd * 4 * H + g * H + u, row = (size_t) dd * 4 * H + (size_t) u * 4 + g;
        for (int k = 0; k < In; k++) wih[row * ip + k] = __float2half(w_ih[src * In + k]);
        for (int k = 0; k < H; k++) whh[src * H + k] = __float2half(w_hh[src * H + k]);
        bih[row] = b_ih[src]; bhh[src] = b_hh[src];
    }
    int maxLen = 0;
    for (int b = 0; b < B; b++) { maxLen = std::max(maxLen, (int) len[b]); for (int t = 0; t < Lmax; t++) for (int k = 0; k < In; k++) xa[((size_t) b * Lmax + t) * ip + k] = __float2half(x[((size_t) b * Lmax + t) * In + k]); }
    __half * dwih = d.put(wih.data(), wih.size()); __half * dwhh = d.put(whh.data(), whh.size()); __half * dx = d.put(xa.data(), xa.size());
    float * dbih = d.put(bih.data(), bih.size()); float * dbhh = d.put(bhh.data(), bhh.size());
    std::vector<int> l32(len, len + B); int * dl = d.put(l32.data(), (size_t) B);
    float * xp = d.get<float>((size_t) B * Lmax * 2048); float * dy = d.get<float>((size_t) B * Lmax * 512);
    if (!dwih || !dwhh || !dx || !dbih || !dbhh || !dl || !xp || !dy) return 1;
    B2_CUDA(cudaMemset(dy, 0, (size_t) B * Lmax * 512 * 4));
    ConvGemmParams p;
    p.A = dx; p.lda = ip; p.W = dwih; p.bias = dbih; p.outF = xp; p.ldo = 2048; p.B = B; p.LmaxIn = Lmax; p.LmaxOut = Lmax; p.lenIn = dl; p.lenOut = dl;
    p.N = 2048; p.Npad = 2048; p.KW = 1; p.CinPad = ip;
    if (conv_gemm(c, p)) return 1;
    LstmParams lp;
    lp.xp = xp; lp.whh = dwhh; lp.bhh = dbhh; lp.out = dy; lp.ldo = 512; lp.len = dl; lp.B = B; lp.Lmax = Lmax; lp.maxLen = maxLen;
    if (bilstm(c, lp)) return 1;
    return finish(c, y, dy, (size_t) B * Lmax * 512 * 4);
}

}  // extern "C"
