// dia.h -- Dia autoregressive decode on the B200, first correct path.
//
// Replaces dia_runner::build_dia_graph / decode / set_inputs (encoder pass, cross K/V store, CFG-paired decoder step, cfg_scale) and the token
// loop of generate_from_batch with check_stopping (reference src/models/dia/model.cpp:324-637,705-737,806-864; src/util.cpp:175-200) under
// sampler::max, for a batch of independent prompts.  Every utterance is TWO sequences throughout, like the reference: the conditional prompt and
// an all-zero unconditional one.  Same plain design as orpheus.h / parler.h (CUDA-core kernels from ar_kernels.cuh; F32 and F16 matrices with the
// reference's numerics for each); logic checked under tests/emu; on a B200 the F32 greedy path reproduces the reference's tokens (profiles/r1i_rowb_first_contact.log).
#pragma once
#include "kokoro.h"   // HostTensor, Arena

namespace b2 {

struct DiaEncLayer { float * pre_sa = nullptr, * post_sa = nullptr; ArW wq, wk, wv, wo, gate, up, down; };
struct DiaDecLayer {
    float * pre_sa = nullptr, * pre_ca = nullptr, * pre_mlp = nullptr;
    ArW sq, sk, sv, so, cq, ck, cv, co, gate, up, down;
};

struct Dia {
    Ctx * ctx = nullptr;
    std::map<std::string, uint32_t>   kv;
    std::map<std::string, HostTensor> host;
    bool prepared = false;
    size_t weight_bytes = 0;
    std::vector<void *> dev_allocs;

    int enc_layers = 0, dec_layers = 0, heads = 0, rep = 1, enc_heads = 0, head_dim = 0, enc_ctx = 0, n_out = 0, vocab = 0, max_gen = 0, max_delay = 15;
    int enc_hidden = 0, hidden = 0, kv_hidden = 0, enc_inner = 0, enc_ffn = 0, ffn = 0, enc_vocab = 0;
    int bos = 1026, eos = 1024, pad = 1025;
    int sm_count = 0;
    uint64_t pdk_launches = 0, pdk_steps = 0;   // cooperative launches of the persistent decode kernel (pdk.cuh) and the decode steps they covered
    float cfg = 3.0f;
    float * enc_embed = nullptr, * enc_norm = nullptr, * tables = nullptr /* [n_out][vocab][hidden] */, * dec_norm = nullptr;
    ArW heads_w;   // [n_out * vocab][hidden]
    __half * heads_hi = nullptr, * heads_lo = nullptr;   // F32 heads (what the quantize tool leaves): fp16 (hi, 2^11-scaled lo) planes for the persistent decode kernel
    std::vector<DiaEncLayer> enc;
    std::vector<DiaDecLayer> dec;

    Arena arena;
    float timing_ms = 0.f;
    cudaEvent_t ev[2] = {nullptr, nullptr};

    int assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
    int prepare();
    // greedy generation for B prompts of byte tokens (each at most enc_ctx long), at most n_steps frames each; an utterance stops early exactly where
    // check_stopping would end the reference's loop: n_generated[b] (may be NULL) is the number of frames it produced, rows past that are zero.
    // out_tokens [B][n_steps][n_out]; out_logits (optional) [B][n_steps][n_out][vocab] (the CFG-combined logits)
    int generate_greedy(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, int32_t * out_tokens, float * out_logits, int32_t * n_generated = nullptr) {
        return generate(B, prompts, n_prompt, n_steps, nullptr, out_tokens, out_logits, n_generated);
    }
    // the same loop under the reference sampler's settings (sampler.cu): sampling == nullptr or do_sample == 0 is the greedy sampler::max
    // teacher (optional, [B][n_steps][n_out]): tokens fed back instead of the produced ones (teacher-forced parity checks, see parler.h)
    int generate(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const ArSampling * sampling, int32_t * out_tokens, float * out_logits, int32_t * n_generated,
                 const int32_t * teacher = nullptr);
    void free_all();
};

int load_gguf_into(Dia * m, const char * path);   // gguf_reader.cpp

}  // namespace b2
