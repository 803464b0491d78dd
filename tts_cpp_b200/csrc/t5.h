// t5.h -- the T5 conditional-prompt encoder of Parler-TTS on the B200 (the step BEFORE Parler's decode loop).
//
// Replaces t5_runner::build_t5_graph / set_inputs / run (reference src/models/parler/t5/model.cpp:183-363) below the unigram tokenizer, for a ragged BATCH of
// independent prompts: token ids in, the [n_tokens][output_size] encoding out -- what parler_tts_runner::update_conditional_prompt hands to
// prep_cross_key_values (src/models/parler/model.cpp:510-518), i.e. the input of b2tts_parler_set_text_encoding.  Launch-per-op over the kernels of
// ar_kernels.cuh (the storage-aware GEMV family: F32, F16 with fp16-rounded activations, Q8_0 / Q5_0 / Q4_0) plus its own (t5.cu): RMS norm with T5's
// eps, bidirectional attention with the relative-position bias, gated GELU; above 32 rows, F16 matrices go through the tcgen05 GEMM of gemm_umma.cu instead.  An encoder pass runs once per voice description, not per audio frame: it is latency-,
// not bandwidth-critical, so the first correct path is the product here (tests/emu runs it on the CPU; GPU parity: tests/test_t5_gpu.py).
#pragma once
#include "kokoro.h"   // HostTensor, Arena, ArW

namespace b2 {

struct T5Layer {
    float * attn_norm = nullptr, * ffn_norm = nullptr;
    ArW q, k, v, o, wi0 /* ffn_up: the GELU'd branch */, wi1 /* ffn_gate */, wo /* ffn_down */;
};

struct T5 {
    Ctx * ctx = nullptr;
    std::map<std::string, uint32_t>   kv;
    std::map<std::string, HostTensor> host;
    bool prepared = false;
    size_t weight_bytes = 0;
    std::vector<void *> dev_allocs;

    int n_layers = 24, heads = 32, head_dim = 64 /* fixed upstream, model.h:46 */, hidden = 2048, ffn = 0, vocab = 0, out_size = 1536, max_ctx = 512, buckets = 32, eos = 1;
    float * embd = nullptr, * out_norm = nullptr, * down_bias = nullptr;
    float * bias_lut = nullptr;    // [heads][2 * max_ctx - 1]: relative_attn_bias[bucket(key - query)][head] for key - query = -(max_ctx - 1) .. max_ctx - 1
    ArW down; bool has_down = false;
    std::vector<T5Layer> layers;

    Arena arena;
    float timing_ms = 0.f;
    bool  last_used_gemm = false;   // the last encode sent its F16 projections through the tensor-core GEMM (> 32 rows; t5.cu)
    cudaEvent_t ev[2] = {nullptr, nullptr};

    int assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
    int prepare();
    // t5_runner::run for B prompts: tokens[b][0 .. n_tokens[b]) (the caller appends EOS like t5_runner::generate) -> out = the encodings back to back,
    // prompt b's rows at sum_{a < b} n_tokens[a], each row output_size() floats
    int encode(int B, const uint32_t * const * tokens, const int32_t * n_tokens, float * out);
    int output_size() const { return has_down ? out_size : hidden; }
    void free_all();
};

int load_gguf_into(T5 * m, const char * path);   // gguf_reader.cpp

}  // namespace b2
