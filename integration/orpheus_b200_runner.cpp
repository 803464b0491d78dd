// integration/orpheus_b200_runner.cpp -- the binding a TTS.cpp maintainer adds to put libb2tts.so under the reference's own API for Orpheus.
//
// A translation unit of the REFERENCE's library (compiled in its tree when -DTTS_B200=ON; nothing here is compiled into libb2tts.so), type-checked by
// `make -C oracle binding_check`.  Host side kept from the reference: loader registry, bpe_tokenizer, the prompt frame (voice prefix, prepended / appended
// control tokens: orpheus_runner::batch_from_sentence, src/models/orpheus/model.cpp:355-369).  Replaced beneath orpheus_runner::generate (model.cpp:406-427):
// the decode loop + sampler (b2tts_orpheus_generate), the loop's exit test and prepare_output_tokens restated here (model.cpp:371-398), and the SNAC decode
// (b2tts_snac_decode_batch).
#include "models/loaders.h"
#include "tokenizer.h"
#include "util.h"

#include "b2tts.h"

#include <algorithm>
#include <array>
#include <cstring>
#include <random>

namespace {

constexpr std::array<const char *, 7> b200_orpheus_voices{ "zoe", "zac", "jess", "leo", "mia", "julia", "leah" };   // model.cpp:7
constexpr std::array<uint32_t, 2> b200_orpheus_prepended{ 128259, 128000 };                                          // model.cpp:8
constexpr std::array<uint32_t, 4> b200_orpheus_appended{ 128009, 128260, 128261, 128257 };                           // model.cpp:9

struct orpheus_b200_runner : tts_generation_runner {
    b2tts_ctx *     ctx       = nullptr;
    b2tts_orpheus * decoder   = nullptr;
    b2tts_snac *    snac      = nullptr;
    bpe_tokenizer * tokenizer = nullptr;
    uint32_t max_generation = 2100, max_context = 1048, stopping_token = 128258;   // orpheus_model defaults (model.h:36-38)
    const uint32_t heads[7] = { 0, 1, 2, 2, 1, 2, 2 };                              // orpheus_model::heads (model.h:44)

    orpheus_b200_runner(const tts_model_loader & loader, bpe_tokenizer * t) : tts_generation_runner{ loader }, tokenizer{ t } {
        sampling_rate   = 24000.0f;
        supports_voices = true;
        if (b2tts_ctx_create(/*device*/ 0, &ctx)) TTS_ABORT("%s\n", b2tts_last_error());
    }
    ~orpheus_b200_runner() override {
        b2tts_orpheus_free(decoder);
        b2tts_snac_free(snac);
        b2tts_ctx_destroy(ctx);
        delete tokenizer;
    }

    void assign_weight(const char * name, ggml_tensor & t) override {               // orpheus_runner::assign_weight routes on the prefix (model.cpp:436-447)
        const std::string_view n{ name };
        int rc = 0;
        if (n.starts_with("snac."))         rc = b2tts_snac_assign_weight(snac, name, (int) t.type, ggml_n_dims(&t), t.ne, t.data, ggml_nbytes(&t));
        else if (n.starts_with("orpheus.")) rc = b2tts_orpheus_assign_weight(decoder, name, (int) t.type, ggml_n_dims(&t), t.ne, t.data, ggml_nbytes(&t));
        if (rc) TTS_ABORT("%s\n", b2tts_last_error());
    }
    void prepare_post_load() override {
        if (b2tts_orpheus_prepare(decoder) || b2tts_snac_prepare(snac)) TTS_ABORT("%s\n", b2tts_last_error());
    }
    std::vector<std::string_view> list_voices() override { return { b200_orpheus_voices.begin(), b200_orpheus_voices.end() }; }

    void generate(const char * sentence, tts_response & output, const generation_configuration & config) override {
        if (!config.voice.empty() && std::find_if(b200_orpheus_voices.begin(), b200_orpheus_voices.end(), [&](const char * v) { return config.voice == v; }) == b200_orpheus_voices.end())
            TTS_ABORT("Voice '%s' is not a valid voice for Orpheus.", config.voice.c_str());
        std::vector<uint32_t> prompt(b200_orpheus_prepended.begin(), b200_orpheus_prepended.end());
        std::string text = sentence;
        if (!config.voice.empty()) text = config.voice + ": " + text;
        tokenizer->tokenize(text, prompt);
        prompt.insert(prompt.end(), b200_orpheus_appended.begin(), b200_orpheus_appended.end());
        if (prompt.size() > max_context) TTS_ABORT("The prompt was too large for the default context window. Try splitting up or shortenning the prompt.");
        const uint32_t * prompts[1]  = { prompt.data() };
        const int32_t    n_prompt[1] = { (int32_t) prompt.size() };
        b2tts_sampling s;
        s.do_sample = config.sample; s.top_k = config.top_k; s.top_p = config.top_p; s.temperature = config.temperature; s.repetition_penalty = config.repetition_penalty;
        s.seed = ((uint64_t) std::random_device{}() << 32) | std::random_device{}();
        // generate_from_batch's exit test runs on the device: the stream ends right after the stopping token (kept) or at max_generation, and the decode loop stops
        // stepping there (a short utterance no longer pays for max_generation steps)
        std::vector<int32_t> stream(max_generation);
        int32_t n_gen = 0;
        if (b2tts_orpheus_generate_until_stop(decoder, 1, prompts, n_prompt, (int) max_generation, &s, stream.data(), &n_gen)) TTS_ABORT("%s\n", b2tts_last_error());
        const size_t n = (size_t) n_gen;
        // prepare_output_tokens: whole 7-token frames; token ii of a frame minus 128266 + ii * 4096 goes to SNAC level heads[ii]
        std::vector<uint32_t> level[3];
        for (size_t i = 0; i < n / 7; i++)
            for (int ii = 0; ii < 7; ii++) level[heads[ii]].push_back((uint32_t) stream[i * 7 + ii] - 128266u - (uint32_t) ii * 4096u);
        std::vector<uint32_t> codes(level[0]);                                     // coarse L/4, medium L/2, fine L -- the layout b2tts_snac_decode_batch takes
        codes.insert(codes.end(), level[1].begin(), level[1].end());
        codes.insert(codes.end(), level[2].begin(), level[2].end());
        const uint32_t * cptr[1]      = { codes.data() };
        const int32_t    fine[1]      = { (int32_t) level[2].size() };
        const float *    pcm[1]       = { nullptr };
        int64_t          n_samples[1] = { 0 };
        if (fine[0] > 0 && b2tts_snac_decode_batch(snac, 1, cptr, fine, pcm, n_samples)) TTS_ABORT("%s\n", b2tts_last_error());
        output.data      = const_cast<float *>(pcm[0]);
        output.n_outputs = (size_t) n_samples[0];
    }
};

struct orpheus_b200_loader final : tts_model_loader {
    orpheus_b200_loader() : tts_model_loader{ "orpheus" } {}
    unique_ptr<tts_generation_runner> from_file(gguf_context * meta, ggml_context *, int, bool, const generation_configuration &) const override {
        auto r = make_unique<orpheus_b200_runner>(*this, bpe_tokenizer_from_gguf(meta));
        std::vector<const char *> keys;
        std::vector<uint32_t>     vals;
        for (int i = 0; i < gguf_get_n_kv(meta); i++) {
            if (gguf_get_kv_type(meta, i) != GGUF_TYPE_UINT32) continue;
            keys.push_back(gguf_get_key(meta, i));
            vals.push_back(gguf_get_val_u32(meta, i));
            const char * k = keys.back();
            if (!strcmp(k, "orpheus.stopping_token_id")) r->stopping_token = vals.back();
            if (!strcmp(k, "orpheus.max_generation_size")) r->max_generation = vals.back();
            if (!strcmp(k, "orpheus.max_context_length")) r->max_context = vals.back();
        }
        if (b2tts_orpheus_create(r->ctx, (int) keys.size(), keys.data(), vals.data(), &r->decoder) ||
            b2tts_snac_create(r->ctx, (int) keys.size(), keys.data(), vals.data(), &r->snac)) TTS_ABORT("%s\n", b2tts_last_error());
        return r;
    }
};

const orpheus_b200_loader orpheus_b200_loader_instance{};

}  // namespace
