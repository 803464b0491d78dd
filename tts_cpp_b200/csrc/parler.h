// parler.h -- Parler-TTS autoregressive decode on the B200, first correct path.
//
// Replaces parler_tts_runner::build_parler_graph / decode / set_inputs / parler_build_kv_store / prep_cross_key_values and the token loop of
// generate_from_batch with its delay pattern (reference src/models/parler/model.cpp:110-173,387-470,520-614,762-786) under sampler::max
// (src/sampler.cpp), for a batch of independent prompts that share the model's stored conditional-prompt encoding.
// Same plain design as orpheus.h (CUDA-core kernels from ar_kernels.cuh, one launch per op); F32 and F16 matrices (the GGUFs `quantize
// --quantized-type F16` writes) with the reference's numerics for each.  Logic checked under tests/emu; on a B200 the F32 greedy path reproduces the reference's tokens (profiles/r1i_rowb_first_contact.log).
#pragma once
#include "kokoro.h"   // HostTensor, Arena

namespace b2 {

struct ParlerLayer {
    float * ln1_w = nullptr, * ln1_b = nullptr; ArW wq, wk, wv, wo;
    float * ln2_w = nullptr, * ln2_b = nullptr; ArW cq, co, wck, wcv; float * cross_k = nullptr, * cross_v = nullptr;   // cross_k / cross_v [n_enc][hidden] = encoding . wck^T / wcv^T
    float * ln3_w = nullptr, * ln3_b = nullptr; ArW fc1, fc2;
};

struct Parler {
    Ctx * ctx = nullptr;
    std::map<std::string, uint32_t>   kv;
    std::map<std::string, HostTensor> host;
    bool prepared = false;
    size_t weight_bytes = 0;
    std::vector<void *> dev_allocs;

    int n_layers = 0, heads = 0, head_dim = 0, hidden = 0, ffn = 0, n_out = 0, vocab = 0, n_enc = 0, max_ctx = 0, tab_rows = 0, prompt_vocab = 0;
    int bos = 1025, eos = 1024;
    float * embed_prompts = nullptr, * pos_embed = nullptr, * tables = nullptr /* [n_out][tab_rows][hidden] */;
    ArW heads_w;   // [n_out * vocab][hidden]
    __half * heads_hi = nullptr, * heads_lo = nullptr;   // F32 output heads as fp16 (hi, 2^11-scaled lo) planes: the persistent decode kernel's fp32-faithful tensor-core product (pdk.cuh)
    int sm_count = 0;
    uint64_t pdk_launches = 0, pdk_steps = 0;             // persistent-kernel launches / decode steps they covered (b2tts_parler_pdk_stats)
    float * ln_w = nullptr, * ln_b = nullptr;
    std::vector<ParlerLayer> layers;

    Arena arena;
    float timing_ms = 0.f;
    cudaEvent_t ev[2] = {nullptr, nullptr};

    int assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
    int prepare();
    // replace the stored conditional-prompt encoding ([n_rows][hidden], e.g. the T5 encoder's output for a new description) and recompute the cross K / V of every
    // layer: parler_tts_model::prep_cross_key_values(n_threads, response) as update_conditional_prompt calls it (reference model.cpp:110-173,510-518)
    int set_text_encoding(const float * enc, int n_rows);
    // greedy generation of n_steps audio frames for B prompts (generate_from_batch's loop with a step cap instead of check_stopping):
    // out_tokens [B][n_steps][n_out]; out_logits (optional) [B][n_steps][n_out][vocab]
    int generate_greedy(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, int32_t * out_tokens, float * out_logits) {
        return generate(B, prompts, n_prompt, n_steps, nullptr, out_tokens, out_logits, nullptr);
    }
    // the same loop under the reference sampler's settings (sampler.cu): sampling == nullptr or do_sample == 0 is the greedy sampler::max
    // n_generated != nullptr turns on the reference's stop rule (parler_context::eos_seen feeding + check_stopping, model.cpp:715-732,795-832): n_generated[b] is
    // the number of frames sequence b produced before the reference's loop would have ended, rows past it are zero; nullptr: fixed-length generation
    // teacher (optional, [B][n_steps][n_out]): the tokens fed back through the delay pattern instead of the produced ones (outputs are still the produced tokens
    // and their logits) -- teacher-forced parity checks: block-quantised models flip near-tied tokens on summation-order noise and then diverge
    int generate(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const ArSampling * sampling, int32_t * out_tokens, float * out_logits,
                 int32_t * n_generated, const int32_t * teacher = nullptr);
    int max_generation = 0;
    void free_all();
};

int load_gguf_into(Parler * m, const char * path);   // gguf_reader.cpp

}  // namespace b2
