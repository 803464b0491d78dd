// integration/dia_b200_runner.cpp -- the binding a TTS.cpp maintainer adds to put libb2tts.so under the reference's own API for Dia.
//
// A translation unit of the REFERENCE's library (compiled in its tree when -DTTS_B200=ON; nothing here is compiled into libb2tts.so), type-checked by
// `make -C oracle binding_check`.  Host side kept from the reference: loader registry, the byte "tokenizer" with its [S1] / [S2] markers
// (dia_runner::tokenize_sentence, src/models/dia/model.cpp:639-683).  Replaced beneath dia_runner::generate (model.cpp:872-897): the encoder pass, the
// CFG-paired decode loop with check_stopping and the sampler (b2tts_dia_generate), adjust_output_tokens' index shuffle restated here (model.cpp:825-847),
// and the DAC decode (b2tts_dac_decode_batch).
#include "models/loaders.h"
#include "util.h"

#include "b2tts.h"

#include <cstring>
#include <random>

namespace {

struct dia_b200_runner : tts_generation_runner {
    b2tts_ctx * ctx     = nullptr;
    b2tts_dia * decoder = nullptr;
    b2tts_dac * dac     = nullptr;
    uint32_t n_output_heads = 9, audio_vocab_size = 1024, max_generation = 3072, max_delay = 15, encoder_context = 1024;
    const uint32_t delay_pattern[9] = { 0, 8, 9, 10, 11, 12, 13, 14, 15 };        // dia_model::delay_pattern (model.h:85)

    explicit dia_b200_runner(const tts_model_loader & loader) : tts_generation_runner{ loader } {
        sampling_rate = 44100.0f;
        if (b2tts_ctx_create(/*device*/ 0, &ctx)) TTS_ABORT("%s\n", b2tts_last_error());
    }
    ~dia_b200_runner() override {
        b2tts_dia_free(decoder);
        b2tts_dac_free(dac);
        b2tts_ctx_destroy(ctx);
    }

    void assign_weight(const char * name, ggml_tensor & t) override {
        const std::string_view n{ name };
        int rc = 0;
        if (n.starts_with("audio_encoder.")) rc = b2tts_dac_assign_weight(dac, name, (int) t.type, ggml_n_dims(&t), t.ne, t.data, ggml_nbytes(&t));
        else if (n.starts_with("dia."))      rc = b2tts_dia_assign_weight(decoder, name, (int) t.type, ggml_n_dims(&t), t.ne, t.data, ggml_nbytes(&t));
        if (rc) TTS_ABORT("%s\n", b2tts_last_error());
    }
    void prepare_post_load() override {
        if (b2tts_dia_prepare(decoder) || b2tts_dac_prepare(dac)) TTS_ABORT("%s\n", b2tts_last_error());
    }

    // tokenize_sentence: a speaker marker in front, a full stop at the end, [S1] / [S2] -> bytes 1 / 2, then the byte of every character
    std::vector<uint32_t> tokenize(std::string sentence) const {
        sentence = strip(sentence);
        const std::string start = sentence.substr(0, 4);
        if (start != "[S1]" && start != "[S2]") sentence = "[S1] " + sentence;
        if (sentence.empty() || sentence.back() != '.') sentence += ".";
        for (const char * tag : { "[S1]", "[S2]" })
            for (size_t pos; (pos = sentence.find(tag)) != std::string::npos;) sentence.replace(pos, 4, std::string(1, tag[2] == '1' ? 1 : 2));
        if (sentence.size() > encoder_context) TTS_ABORT("Dia currently only supports a max of %d characters and received an input of %d characters.", (int) encoder_context, (int) sentence.size());
        std::vector<uint32_t> toks;
        for (auto c : sentence) toks.push_back((uint32_t) c);
        return toks;
    }

    void generate(const char * sentence, tts_response & output, const generation_configuration & config) override {
        if (!(config.max_tokens == 0 || config.max_tokens > (int) max_delay)) TTS_ABORT("max_tokens must be 0 or greater than the maximum delay.\n");
        const uint32_t max_gen = config.max_tokens > (int) max_delay ? (uint32_t) config.max_tokens : max_generation;
        if (b2tts_dia_set_max_generation(decoder, (int) max_gen)) TTS_ABORT("%s\n", b2tts_last_error());
        const std::vector<uint32_t> prompt = tokenize(sentence);
        const uint32_t * prompts[1]  = { prompt.data() };
        const int32_t    n_prompt[1] = { (int32_t) prompt.size() };
        b2tts_sampling s;
        s.do_sample = config.sample; s.top_k = config.top_k; s.top_p = config.top_p; s.temperature = config.temperature; s.repetition_penalty = config.repetition_penalty;
        s.seed = ((uint64_t) std::random_device{}() << 32) | std::random_device{}();
        const int n_steps = (int) max_gen;                                         // check_stopping ends the loop max_delay steps after position max_gen - max_delay
        std::vector<int32_t> steps((size_t) n_steps * n_output_heads);
        int32_t n_generated = 0;
        if (b2tts_dia_generate(decoder, 1, prompts, n_prompt, n_steps, &s, steps.data(), nullptr, &n_generated)) TTS_ABORT("%s\n", b2tts_last_error());
        // adjust_output_tokens: frame i takes head h's token from step i + delay_pattern[h]; frames holding a special id are dropped
        std::vector<uint32_t> frames;
        const int H = (int) n_output_heads;
        for (int i = 0; i < n_generated - (int) max_delay; i++) {
            bool keep = true;
            for (int h = 0; h < H && keep; h++) keep = (uint32_t) steps[(size_t) (i + delay_pattern[h]) * H + h] < audio_vocab_size;
            if (!keep) continue;
            for (int h = 0; h < H; h++) frames.push_back((uint32_t) steps[(size_t) (i + delay_pattern[h]) * H + h]);
        }
        const uint32_t * codes[1]     = { frames.data() };
        const int32_t    n_frames[1]  = { (int32_t) (frames.size() / n_output_heads) };
        const float *    pcm[1]       = { nullptr };
        int64_t          n_samples[1] = { 0 };
        if (n_frames[0] > 0 && b2tts_dac_decode_batch(dac, 1, codes, n_frames, pcm, n_samples)) TTS_ABORT("%s\n", b2tts_last_error());
        output.data      = const_cast<float *>(pcm[0]);
        output.n_outputs = (size_t) n_samples[0];
    }
};

struct dia_b200_loader final : tts_model_loader {
    dia_b200_loader() : tts_model_loader{ "dia" } {}
    unique_ptr<tts_generation_runner> from_file(gguf_context * meta, ggml_context *, int, bool, const generation_configuration &) const override {
        auto r = make_unique<dia_b200_runner>(*this);
        std::vector<const char *> keys;
        std::vector<uint32_t>     vals;
        for (int i = 0; i < gguf_get_n_kv(meta); i++) {
            if (gguf_get_kv_type(meta, i) != GGUF_TYPE_UINT32) continue;
            keys.push_back(gguf_get_key(meta, i));
            vals.push_back(gguf_get_val_u32(meta, i));
            const char * k = keys.back();
            if (!strcmp(k, "dia.decoder.output_heads")) r->n_output_heads = vals.back();
            if (!strcmp(k, "dia.decoder.audio_vocab_size")) r->audio_vocab_size = vals.back();
            if (!strcmp(k, "dia.decoder.max_generation_size")) r->max_generation = vals.back();
            if (!strcmp(k, "dia.max_delay")) r->max_delay = vals.back();
            if (!strcmp(k, "dia.encoder.max_context_length")) r->encoder_context = vals.back();
        }
        if (r->n_output_heads > 9) TTS_ABORT("dia_b200_runner: %u output heads\n", r->n_output_heads);
        if (b2tts_dia_create(r->ctx, (int) keys.size(), keys.data(), vals.data(), &r->decoder) ||
            b2tts_dac_create(r->ctx, (int) keys.size(), keys.data(), vals.data(), &r->dac)) TTS_ABORT("%s\n", b2tts_last_error());
        return r;
    }
};

const dia_b200_loader dia_b200_loader_instance{};

}  // namespace
