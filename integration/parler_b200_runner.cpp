// integration/parler_b200_runner.cpp -- the binding a TTS.cpp maintainer adds to put libb2tts.so under the reference's own API for Parler-TTS.
//
// A translation unit of the REFERENCE's library (it includes the reference's headers and is compiled in its tree when -DTTS_B200=ON); nothing here is
// compiled into libb2tts.so.  It keeps the reference's host side -- loader registry, unigram tokenizer, generation_configuration -- and replaces what runs
// beneath parler_tts_runner::generate: the decode loop + sampler (b2tts_parler_generate, including the stop rule) and the DAC decode
// (b2tts_dac_decode_batch).  `make -C oracle binding_check` type-checks it against the reference headers where /root/reference exists.
//
// Reference interfaces used: tts_generation_runner / tts_model_loader / generation_configuration / tts_response (include/common.h:13-17,45-66,68-94), the
// loader registry (src/models/loaders.cpp:11-31,79-89), unigram_tokenizer (src/tokenizer.h:35-53).  Behaviour mirrored: parler_model_loader::from_file
// (src/models/parler/loader.cpp:8-26), batch_from_sentence (model.cpp:473-498: tokens + EOS), parler_tts_runner::generate / generate_from_batch /
// adjust_output_tokens (model.cpp:734-792,838-861), assign_weight's routing of "audio_encoder.*" to the DAC (model.cpp:499-512).
// update_conditional_prompt runs the T5 encoder on the GPU (b2tts_t5_*) and hands its output to b2tts_parler_set_text_encoding (cross K / V recomputed on the
// device).  Not supported: use_cross_attn = false.
#include "models/loaders.h"
#include "models/parler/t5/model.h"
#include "tokenizer.h"
#include "util.h"

#include "b2tts.h"

#include <cstring>
#include <random>

namespace {

struct parler_b200_runner : tts_generation_runner {
    b2tts_ctx *         ctx       = nullptr;
    b2tts_parler *      decoder   = nullptr;
    b2tts_dac *         dac       = nullptr;
    b2tts_t5 *          t5        = nullptr;      // the conditional-prompt encoder, loaded on the first update_conditional_prompt
    std::string         t5_path;
    unigram_tokenizer * tokenizer = nullptr;
    uint32_t n_output_heads = 9, audio_vocab_size = 1024, max_generation = 2580;
    int n_threads = 4;

    parler_b200_runner(const tts_model_loader & loader, unigram_tokenizer * t) : tts_generation_runner{ loader }, tokenizer{ t } {
        sampling_rate = 44100.0f;
        if (b2tts_ctx_create(/*device*/ 0, &ctx)) TTS_ABORT("%s\n", b2tts_last_error());
    }
    ~parler_b200_runner() override {
        b2tts_parler_free(decoder);
        b2tts_t5_free(t5);
        b2tts_dac_free(dac);
        b2tts_ctx_destroy(ctx);
        delete tokenizer;
    }

    // runner_from_file streams every tensor of the GGUF (loaders.cpp:79-88); the reference routes on the name prefix (parler/model.cpp:499-512)
    void assign_weight(const char * name, ggml_tensor & t) override {
        const std::string_view n{ name };
        int rc = 0;
        if (n.starts_with("audio_encoder.")) rc = b2tts_dac_assign_weight(dac, name, (int) t.type, ggml_n_dims(&t), t.ne, t.data, ggml_nbytes(&t));
        else if (n.starts_with("decoder."))  rc = b2tts_parler_assign_weight(decoder, name, (int) t.type, ggml_n_dims(&t), t.ne, t.data, ggml_nbytes(&t));
        if (rc) TTS_ABORT("%s\n", b2tts_last_error());
    }
    void prepare_post_load() override {
        if (b2tts_parler_prepare(decoder) || b2tts_dac_prepare(dac)) TTS_ABORT("%s\n", b2tts_last_error());
    }
    // parler_tts_runner::update_conditional_prompt (model.cpp:510-518) entirely on the device: the reference's own unigram tokenizer (host string work, kept), EOS
    // appended as t5_runner::generate does (t5/model.cpp:365-371), the T5 encoder pass on the GPU (b2tts_t5_encode; the GGUF stays loaded between calls), then the
    // cross-attention K / V of every layer recomputed from its output (b2tts_parler_set_text_encoding = prep_cross_key_values)
    void update_conditional_prompt(const char * file_path, const char * prompt) override {
        if (!t5 || t5_path != file_path) {
            b2tts_t5_free(t5); t5 = nullptr;
            if (b2tts_t5_load_gguf(ctx, file_path, &t5)) TTS_ABORT("%s\n", b2tts_last_error());
            t5_path = file_path;
        }
        int out_size = 0, eos = 1;
        if (b2tts_t5_info(t5, nullptr, nullptr, &out_size, nullptr, nullptr, &eos)) TTS_ABORT("%s\n", b2tts_last_error());
        if (!tokenizer->init) tokenizer->initialize_tokenizer();              // text_encoder_from_file does this for the shared tokenizer (t5/model.cpp:385-387)
        std::vector<uint32_t> ids;
        tokenizer->tokenize(prompt, ids);
        ids.push_back((uint32_t) eos);
        std::vector<float> enc(ids.size() * (size_t) out_size);
        const uint32_t * pp = ids.data();
        const int32_t    n  = (int32_t) ids.size();
        if (b2tts_t5_encode(t5, 1, &pp, &n, enc.data()) || b2tts_parler_set_text_encoding(decoder, enc.data(), n)) TTS_ABORT("%s\n", b2tts_last_error());
    }

    // parler_tts_runner::adjust_output_tokens (model.cpp:734-760): undo the delay pattern (head h lags h steps) and drop frames holding a special id
    void frames_from_steps(const int32_t * steps, int n_steps, std::vector<uint32_t> & frames) const {
        const int H = (int) n_output_heads;
        for (int i = 0; i + H - 1 < n_steps; i++) {
            bool keep = true;
            for (int h = 0; h < H && keep; h++) keep = (uint32_t) steps[(size_t) (i + h) * H + h] < audio_vocab_size;
            if (!keep) continue;
            for (int h = 0; h < H; h++) frames.push_back((uint32_t) steps[(size_t) (i + h) * H + h]);
        }
    }

    void generate(const char * sentence, tts_response & output, const generation_configuration & config) override {
        if (!config.use_cross_attn) TTS_ABORT("parler_b200_runner: use_cross_attn = false is not supported.\n");
        std::vector<uint32_t> prompt;
        tokenizer->tokenize(sentence, prompt);
        prompt.push_back(tokenizer->eos_token);                                  // batch_from_sentence (model.cpp:473-479)
        const uint32_t * prompts[1]  = { prompt.data() };
        const int32_t    n_prompt[1] = { (int32_t) prompt.size() };
        b2tts_sampling s;
        s.do_sample = config.sample; s.top_k = config.top_k; s.top_p = config.top_p; s.temperature = config.temperature; s.repetition_penalty = config.repetition_penalty;
        s.seed = ((uint64_t) std::random_device{}() << 32) | std::random_device{}();   // the reference seeds its generator from std::random_device per call
        const int n_steps = (int) max_generation - (int) prompt.size();           // check_stopping ends at position max_generation at the latest
        if (n_steps <= 0) TTS_ABORT("The prompt was too large for the default context window.\n");
        std::vector<int32_t> steps((size_t) n_steps * n_output_heads);
        int32_t n_generated = 0;
        if (b2tts_parler_generate(decoder, 1, prompts, n_prompt, n_steps, &s, steps.data(), nullptr, &n_generated)) TTS_ABORT("%s\n", b2tts_last_error());
        std::vector<uint32_t> frames;
        frames_from_steps(steps.data(), n_generated, frames);
        const uint32_t * codes[1]    = { frames.data() };
        const int32_t    n_frames[1] = { (int32_t) (frames.size() / n_output_heads) };
        const float *    pcm[1]      = { nullptr };
        int64_t          n_samples[1] = { 0 };
        if (n_frames[0] > 0 && b2tts_dac_decode_batch(dac, 1, codes, n_frames, pcm, n_samples)) TTS_ABORT("%s\n", b2tts_last_error());
        output.data      = const_cast<float *>(pcm[0]);                           // runner-owned pinned buffer, valid until the next call (like dac_model.cpp:191)
        output.n_outputs = (size_t) n_samples[0];
    }
};

// registers under the architecture string of the GGML loader: the registry is an emplace (first registration wins, loaders.cpp:11-31),
// so this object has to be linked ahead of the stock parler loader or replace it when TTS_B200 is on
struct parler_b200_loader final : tts_model_loader {
    parler_b200_loader() : tts_model_loader{ "parler-tts" } {}
    unique_ptr<tts_generation_runner> from_file(gguf_context * meta, ggml_context *, int n_threads, bool, const generation_configuration &) const override {
        unigram_tokenizer * ut = unigram_tokenizer_from_gguf(meta);
        ut->initialize_tokenizer();
        auto r = make_unique<parler_b200_runner>(*this, ut);
        r->n_threads = n_threads;
        std::vector<const char *> keys;
        std::vector<uint32_t>     vals;
        for (int i = 0; i < gguf_get_n_kv(meta); i++) {
            if (gguf_get_kv_type(meta, i) != GGUF_TYPE_UINT32) continue;
            keys.push_back(gguf_get_key(meta, i));
            vals.push_back(gguf_get_val_u32(meta, i));
            if (!strcmp(keys.back(), "parler-tts.decoder.output_heads")) r->n_output_heads = vals.back();
            if (!strcmp(keys.back(), "parler-tts.decoder.audio_vocab_size")) r->audio_vocab_size = vals.back();
            if (!strcmp(keys.back(), "parler-tts.decoder.max_generation")) r->max_generation = vals.back();
        }
        if (b2tts_parler_create(r->ctx, (int) keys.size(), keys.data(), vals.data(), &r->decoder) ||
            b2tts_dac_create(r->ctx, (int) keys.size(), keys.data(), vals.data(), &r->dac)) TTS_ABORT("%s\n", b2tts_last_error());
        return r;
    }
};

const parler_b200_loader parler_b200_loader_instance{};

}  // namespace
