// kokoro.h -- B200-native Kokoro-82M model: HBM-resident weights + batched duration/generation forward.
#pragma once
#include "kernels.cuh"
#include <map>
#include <string>
#include <vector>

namespace b2 {

struct HostTensor {             // fp32 host copy of one GGUF tensor (numpy shape = reversed ggml ne)
    std::vector<float>   v;
    std::vector<int64_t> shape; // outermost first
    bool                 f16 = false;
    int                  qtype = 0;   // ggml type of a block-quantised tensor (2 Q4_0, 6 Q5_0, 8 Q8_0): `raw` holds its blocks, `v` the dequantised values
    std::vector<uint8_t> raw;
};

struct W16 {                    // fp16 GEMM weight [Npad][KW][CinPad]
    __half * w = nullptr;
    int N = 0, Npad = 0, KW = 1, Cin = 0, CinPad = 0;
};

struct ArW {                    // a weight matrix [N][K] of the autoregressive decode paths: fp32, fp16 when the GGUF stores it as F16, or block-quantised (qtype 2 / 6 / 8)
    const void * p = nullptr;   // fp32 / fp16 values; quantised: the 4- or 8-bit plane, rows of K/2 (Q4_0, Q5_0: byte j of a block = elements j | j+16) or K (Q8_0) bytes
    bool f16 = false;
    int  qtype = 0;
    const void * scales = nullptr;   // quantised: [N][K/32] fp16 block scales
    const void * qh = nullptr;       // Q5_0: [N][K/32] uint32, the fifth bits
};

struct Lstm {
    W16      wih;               // N = 2048 rows ordered (dir, unit, gate)
    float *  bih = nullptr;     // [2048] same order
    __half * whh = nullptr;     // [2][1024][256]
    float *  bhh = nullptr;     // [2][1024]
};

struct StyleSlot { int goff = 0, boff = 0, C = 0; };

struct AdaBlock {
    int cin = 0, cout = 0; bool pool = false, has1x1 = false;
    W16 conv1, conv2, conv1x1;
    float * b1 = nullptr, * b2 = nullptr, * poolw = nullptr, * poolb = nullptr;
    StyleSlot n1, n2;
};

struct GenResBlock {
    int C = 0, K = 0; int pad[3] = {0, 0, 0}, dil[3] = {1, 1, 1};
    W16 c1[3], c2[3];
    float * b1[3], * b2[3], * a1[3], * a2[3];
    StyleSlot s1[3], s2[3];
};

struct Arena {
    char * base = nullptr; size_t cap = 0, off = 0;
    int reserve(size_t bytes);
    void * alloc(size_t bytes);
    void release();
};

struct Tap { const void * ptr = nullptr; int64_t rows = 0, cols = 0, ld = 0, padded = 0; };

struct Kokoro {
    Ctx * ctx = nullptr;
    std::map<std::string, uint32_t>   kv;
    std::map<std::string, HostTensor> host;     // until prepare()
    bool prepared = false;
    size_t weight_bytes = 0;
    std::vector<void *> dev_allocs;

    // ALBERT
    int n_vocab = 0, n_positions = 512;      // rows of the token / position tables (ids are validated against them before they reach the device)
    float * tok_embd = nullptr, * pos_embd = nullptr, * type_embd = nullptr, * in_nw = nullptr, * in_nb = nullptr;
    float * embd_w = nullptr, * embd_b = nullptr;
    W16 qkv, o, ffn, ffn_out;
    float * qkv_b = nullptr, * o_b = nullptr, * ffn_b = nullptr, * ffn_out_b = nullptr;
    float * attn_norm_w = nullptr, * attn_norm_b = nullptr, * ffn_norm_w = nullptr, * ffn_norm_b = nullptr;
    int recurrence = 12, heads = 12;
    // prosody predictor
    W16 encode; float * encode_b = nullptr;
    Lstm dp_lstm[3], dur_lstm, shared_lstm, text_lstm;
    StyleSlot dp_ada[3];
    W16 dur_proj; float * dur_proj_b = nullptr;
    AdaBlock f0_blocks[3], n_blocks[3];
    W16 f0_proj, n_proj; float * f0_proj_b = nullptr, * n_proj_b = nullptr;
    // text encoder
    __half * text_embd = nullptr;
    W16 te_conv[3]; float * te_b[3], * te_gamma[3], * te_beta[3];
    // decoder
    float f0_conv_w[3], n_conv_w[3], f0_conv_b[1], n_conv_b[1];
    W16 asr_conv; float * asr_conv_b = nullptr;
    AdaBlock enc_block, dec_blocks[4];
    // generator
    float m_src_w[9]; float m_src_b = 0.f;
    struct Up {
        float * w = nullptr, * b = nullptr; int K = 0, Cin = 0, Cout = 0, stride = 0, pad = 0;
        // polyphase tensor-core form (K == 2*stride): N = stride*Cout phases, 2 taps, 3*Cin split-fp16 channels
        bool poly = false; W16 w3; float * b_rep = nullptr;
    } ups[2];
    // strided noise conv: `wp` is its polyphase form (stride 1 over rows of `stride` input frames, see Kokoro::prepare) for the tcgen05 kernel
    struct NoiseConv { W16 w, wp; float * b = nullptr; int stride = 1, pad = 0, padp = 0; bool poly = false; } nconv[2];
    GenResBlock nres[2], res[6];
    W16 conv_post; float * conv_post_b = nullptr; int post_pad = 3;
    // style projections (F32): kind 0 = prosody style (voice[:,128:256]), 1 = decoder style (voice[:,0:128])
    std::vector<float> sty_w_host[2], sty_b_host[2];
    float * sty_w[2] = {nullptr, nullptr}, * sty_b[2] = {nullptr, nullptr};
    int sty_n[2] = {0, 0};
    // voices
    std::vector<std::string> voice_names;
    std::map<std::string, std::vector<float>> voices_host;

    // runtime
    Arena a1, a2;
    float * pcm_pinned = nullptr; size_t pcm_pinned_cap = 0;
    float * lens_pinned = nullptr; size_t lens_pinned_cap = 0;
    std::vector<const float *> pcm_ptrs;
    bool taps_on = false;
    std::map<std::string, Tap> taps;
    std::map<std::string, std::vector<float>> overrides;
    float timings[3] = {0, 0, 0};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

    int assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes);
    int prepare();
    int run_batch(int B, const uint32_t * tokens, const int32_t * n_tokens, const char * voice, const uint64_t * noise_skip, const float ** pcm,
                  int64_t * n_samples, const float ** durations);
    // chain_noise: the utterances are consecutive generate() calls of ONE reference process (the chunks of a long prompt, kokoro/model.cpp:1430-1447, or a drained
    // queue): utterance b's noise starts where utterance b-1's ended (9 * 600 * T draws each, util.cpp:66-72,140-172), the first one at chain_noise_start.  The
    // offsets need the durations, so they are set between the two passes.
    // keep_on_device: skip the device -> host copy of the PCM (the multi-GPU gather sends it over NCCL from device memory); last_pcm_dev [B][last_pcm_stride] stays
    // valid until the next call on this model
    bool keep_on_device = false; const float * last_pcm_dev = nullptr; int64_t last_pcm_stride = 0;
    bool chain_noise = false; uint64_t chain_noise_start = 0;
    std::vector<unsigned long long> chain_skips;
    void free_all();
};

int load_gguf_into(Kokoro * m, const char * path);   // gguf_reader.cpp

}  // namespace b2
