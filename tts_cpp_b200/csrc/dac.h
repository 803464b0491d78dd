// dac.h -- DAC neural audio codec decoder on the B200: codebook indices -> PCM.
//
// Replaces dac_runner::run / build_dac_graph / dac_build_audio_inputs (reference src/decoder/dac_model.cpp:100-123,146-212)
// and general_neural_audio_codec::build_layer / build_residual_unit / build_quantize_layer
// (src/decoder/general_neural_audio_codec.cpp:133-172) for a batch of independent utterances.
#pragma once
#include "kokoro.h"   // HostTensor, W16, Arena

namespace b2 {

// one Conv1d of the codec.  An F32 kernel means the reference computes the convolution in fp32 (ggml_conv_1d chooses an F32 im2col when
// kernel and input are F32, ggml/src/ggml.c:3877-3881): the tensor cores then run it as three fp16 products of split operands
// (x = hi + lo, W = Whi + Wlo: hi*Whi + lo*Whi + hi*Wlo), operand channels = 3 * Cin.  An F16 kernel takes the plain fp16 path.
struct DacConv {
    W16     w;
    float * b = nullptr;
    int     Cin = 0, Cout = 0, K = 1, dil = 1, pad = 0;
    bool    split = true;
};

struct DacUnit { float * a1 = nullptr, * a2 = nullptr; DacConv c1, c2; };

struct DacLayer {
    float * alpha = nullptr;
    int     Cin = 0, Cout = 0, stride = 1, pad = 0;
    W16     w3;                 // polyphase ConvTranspose (K == 2*stride): N = stride*Cout phases, 2 taps, 3*Cin split channels
    float * b_rep = nullptr;    // bias replicated per phase
    DacUnit res[3];
};

struct Dac {
    Ctx * ctx = nullptr;
    std::map<std::string, uint32_t>   kv;
    std::map<std::string, HostTensor> host;   // until prepare()
    bool prepared = false;
    size_t weight_bytes = 0;
    std::vector<void *> dev_allocs;

    int n_heads = 9, n_codes = 1024, latent = 1024, up_factor = 512;
    float * tables = nullptr;   // [n_heads][n_codes][latent]: out_proj(codebook row) + bias per head (the quantizer is a table lookup)
    DacConv initial, final_conv;
    DacLayer layers[4];
    float * final_alpha = nullptr;

    Arena arena;
    float * pcm_pinned = nullptr; size_t pcm_pinned_cap = 0;
    float timing_ms = 0.f;
    cudaEvent_t ev[2] = {nullptr, nullptr};

    int assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
    int prepare();
    // codes[b]: frames[b] * n_heads indices, frame-major (the layout dac_runner::run takes); pcm[b]: frames[b] * up_factor samples in a
    // runner-owned pinned host buffer valid until the next call (like tts_response.data, dac_model.cpp:191)
    int decode_batch(int B, const uint32_t * const * codes, const int32_t * frames, const float ** pcm, int64_t * n_samples);
    void free_all();
};

int load_gguf_into(Dac * m, const char * path);   // gguf_reader.cpp

// SNAC codec decoder (reference src/decoder/snac_model.cpp:86-208): three code streams at L/4, L/2, L frames; depthwise convolutions;
// a noise block after every ConvTranspose (x += conv1x1(x) * n[t], n ~ N(0,1) from the reference's process-wide
// std::normal_distribution over std::default_random_engine, src/util.cpp:74-80).
struct SnacLayer {
    float * alpha = nullptr;
    int     Cin = 0, Cout = 0, stride = 1, pad = 0;
    W16     w3; float * b_rep = nullptr;                 // polyphase ConvTranspose, as DacLayer
    DacConv noise;                                       // 1x1, no bias
    struct Unit { float * a1 = nullptr, * dw_w = nullptr, * dw_b = nullptr; int dil = 1; float * a2 = nullptr; DacConv c2; } res[3];
};

struct Snac {
    Ctx * ctx = nullptr;
    std::map<std::string, uint32_t>   kv;
    std::map<std::string, HostTensor> host;
    bool prepared = false;
    size_t weight_bytes = 0;
    std::vector<void *> dev_allocs;

    int n_codes = 4096, latent = 768, up_factor = 512;
    int repeats[3] = {4, 2, 1};
    int noise_steps[4] = {8, 64, 256, 512};
    float * tables = nullptr;        // [3][n_codes][latent]
    float * in_w = nullptr, * in_b = nullptr;            // depthwise k7
    DacConv up, final_conv;
    SnacLayer layers[4];
    float * final_alpha = nullptr;
    void * noise_engine = nullptr;   // std::minstd_rand0 + std::normal_distribution<float> state (host): the reference's static generator

    Arena arena;
    float * pcm_pinned = nullptr; size_t pcm_pinned_cap = 0;
    float timing_ms = 0.f;
    cudaEvent_t ev[2] = {nullptr, nullptr};

    int assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
    int prepare();
    // codes[b]: L/4 coarse, then L/2 medium, then L fine indices (the three vectors snac_runner::run takes, concatenated); L = fine_frames[b]
    int decode_batch(int B, const uint32_t * const * codes, const int32_t * fine_frames, const float ** pcm, int64_t * n_samples);
    void reset_noise();              // back to the state of a fresh process
    void free_all();
};

int load_gguf_into(Snac * m, const char * path);  // gguf_reader.cpp

}  // namespace b2
