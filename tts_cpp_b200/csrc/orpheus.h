// orpheus.h -- Orpheus (llama-3 style) autoregressive decode on the B200, first correct path.
//
// Replaces orpheus_runner::build_orpheus_graph / decode / set_inputs / orpheus_build_kv_store / generate_from_batch's token loop
// (reference src/models/orpheus/model.cpp:196-353,389-398) and sampler::max (src/sampler.cpp) for a batch of independent sequences.
// v1 is deliberately plain: fp32 weights (the only dtype the reference supports for Orpheus), fp32 CUDA-core GEMVs that stream each
// weight matrix once per step for the whole batch, a compact (un-expanded) GQA KV cache, softmax with the reference's double-accumulated
// sum, argmax on the device.  Written after round 1's GPU budget was spent: parity tests are xfail(strict=False) until validated.
#pragma once
#include "kokoro.h"   // HostTensor, Arena

namespace b2 {

struct OrpheusLayer {
    float * in_norm = nullptr, * post_norm = nullptr;
    ArW wq, wk, wv, wo, wgate, wup, wdown;      // F32 (the reference's only Orpheus dtype), F16, or Q4_0 / Q5_0 / Q8_0 blocks (BASELINE config 5: q8_0 -- our own writer, the reference's quantize tool refuses Orpheus)
};

struct Orpheus {
    Ctx * ctx = nullptr;
    std::map<std::string, uint32_t>   kv;
    std::map<std::string, HostTensor> host;
    bool prepared = false;
    size_t weight_bytes = 0;
    std::vector<void *> dev_allocs;

    int vocab = 0, heads = 0, kv_heads = 0, head_dim = 0, hidden = 0, kv_hidden = 0, ffn = 0, n_layers = 0;
    int stopping_token = -1;
    float * embed = nullptr, * out_norm = nullptr, * rope_ff = nullptr;
    ArW head;
    int sm_count = 0;
    uint64_t pdk_launches = 0, pdk_steps = 0;   // cooperative launches of the persistent decode kernel (pdk.cuh) and the decode steps they covered
    int max_context = 0;                        // prompt + generated positions the model supports (orpheus.context_length when present, else unbounded by metadata)
    std::vector<OrpheusLayer> layers;
    // B2TTS_AR_MMA=1: fp16 (hi, 2^11-scaled lo) splits of the F32 matrices for the tensor-core batched GEMV (ar_kernels.cuh gemv_mma_kernel<true>), keyed by the fp32 pointer
    std::map<const float *, std::pair<const void *, const void *>> split;

    Arena arena;
    float timing_ms = 0.f;
    cudaEvent_t ev[2] = {nullptr, nullptr};

    int assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
    int prepare();
    // greedy continuation of B prompts for n_steps tokens each (generate_from_batch's loop without the stop condition):
    // out_tokens [B][n_steps]; out_logits (optional) [B][n_steps][vocab]
    int generate_greedy(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, int32_t * out_tokens, float * out_logits) {
        return generate(B, prompts, n_prompt, n_steps, nullptr, out_tokens, out_logits);
    }
    // the same loop under the reference sampler's settings (sampler.cu): sampling == nullptr or do_sample == 0 is the greedy sampler::max
    // n_generated (optional): turns on the reference's stop rule (generate_from_batch, model.cpp:389-398: the loop ends when the stopping token was produced): per
    // sequence the number of tokens up to and including its stopping token (n_steps when it never came); the batch stops stepping once every sequence has ended
    int generate(int B, const uint32_t * const * prompts, const int32_t * n_prompt, int n_steps, const ArSampling * sampling, int32_t * out_tokens, float * out_logits,
                 int32_t * n_generated = nullptr);
    void free_all();
};

int load_gguf_into(Orpheus * m, const char * path);   // gguf_reader.cpp

}  // namespace b2
