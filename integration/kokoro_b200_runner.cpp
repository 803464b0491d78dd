// integration/kokoro_b200_runner.cpp -- the binding a TTS.cpp maintainer adds to put libb2tts.so under the reference's own API.
//
// A translation unit of the REFERENCE's library (it includes the reference's headers and is compiled in its tree, e.g. as
// src/models/kokoro_b200/model.cpp with -DTTS_B200=ON); nothing here is compiled into libb2tts.so.  It keeps the reference's host side --
// registry, loader, phonemizer, tokenizer, chunking -- and replaces what runs beneath kokoro_runner::run with one C-ABI call.
// `make -C oracle binding_check` type-checks it against the reference headers where /root/reference exists (tests/test_host_cpu.py).
//
// Reference interfaces used: tts_generation_runner / tts_model_loader (include/common.h:26-31,68-94), the loader registry
// (src/models/loaders.cpp:11-31,79-89), single_pass_tokenizer (src/tokenizer.h:58-76), phonemizer (src/models/kokoro/phonemizer.h:529),
// strip / split / replace_any (src/util.h:64-67).  Behaviour mirrored: kokoro_runner::generate and tokenize_chunks
// (src/models/kokoro/model.cpp:1340-1450) -- with the chunks of a long prompt synthesised as ONE batch instead of one after another.
#include "models/loaders.h"
#include "models/kokoro/phonemizer.h"
#include "tokenizer.h"
#include "util.h"

#include "b2tts.h"
#include "tts_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

struct kokoro_b200_runner : tts_generation_runner {
    b2tts_ctx *             ctx       = nullptr;
    b2tts_kokoro *          model     = nullptr;
    single_pass_tokenizer * tokenizer = nullptr;
    phonemizer *            phmzr     = nullptr;
    uint32_t max_context_length = 512, bos_token_id = 0, eos_token_id = 0;
    uint64_t noise_draws = 0;            // position in the reference's process-wide uniform stream (src/util.cpp:66-72)
    std::vector<float> joined;           // multi-chunk responses are concatenated here (the reference mallocs and leaks, tts_model.cpp:8-20)

    kokoro_b200_runner(const tts_model_loader & loader, single_pass_tokenizer * t, phonemizer * p) : tts_generation_runner{ loader }, tokenizer{ t }, phmzr{ p } {
        sampling_rate   = 24000.0f;
        supports_voices = true;
        if (b2tts_ctx_create(/*device*/ 0, &ctx)) TTS_ABORT("%s\n", b2tts_last_error());
    }
    ~kokoro_b200_runner() override {
        b2tts_kokoro_free(model);
        b2tts_ctx_destroy(ctx);
        delete phmzr;
    }

    // runner_from_file streams (name, tensor) pairs exactly as it does for the GGML runner (loaders.cpp:79-88)
    void assign_weight(const char * name, ggml_tensor & t) override {
        if (b2tts_kokoro_assign_weight(model, name, (int) t.type, ggml_n_dims(&t), t.ne, t.data, ggml_nbytes(&t))) TTS_ABORT("%s\n", b2tts_last_error());
    }
    void prepare_post_load() override {
        if (b2tts_kokoro_prepare(model)) TTS_ABORT("%s\n", b2tts_last_error());
    }
    std::vector<std::string_view> list_voices() override {
        std::vector<std::string_view> v;
        for (int i = 0; i < b2tts_kokoro_n_voices(model); i++) v.emplace_back(b2tts_kokoro_voice_name(model, i));
        return v;
    }

    // one clause -> token chunks of at most max_context_length (BOS/EOS included), split at the last space that fits
    void chunk_clause(const std::string & clause, uint32_t space_id, std::vector<std::vector<uint32_t>> & chunks) {
        std::vector<uint32_t> toks;
        tokenizer->tokenize(clause, toks);
        const size_t room = max_context_length - 2;
        size_t start = 0;
        while (start < toks.size()) {
            size_t end = std::min(toks.size(), start + room);
            if (end < toks.size()) {
                size_t sp = end;
                while (sp > start && toks[sp - 1] != space_id) sp--;
                if (sp > start) end = sp;
            }
            std::vector<uint32_t> c{ bos_token_id };
            c.insert(c.end(), toks.begin() + (long) start, toks.begin() + (long) end);
            c.push_back(eos_token_id);
            chunks.push_back(std::move(c));
            start = end;
        }
    }

    // kokoro_runner::run for all chunks at once.  The chunks are consecutive run() calls in the reference (model.cpp:1430-1447): chunk b's noise continues the
    // process-wide uniform stream where chunk b-1 left it; the library chains the offsets itself once it knows the durations (b2tts_kokoro_run_chunks).
    void run_chunks(const std::vector<std::vector<uint32_t>> & chunks, const std::string & voice, std::vector<const float *> & pcm, std::vector<int64_t> & ns) {
        std::vector<uint32_t> toks;
        std::vector<int32_t>  n;
        for (auto & c : chunks) { toks.insert(toks.end(), c.begin(), c.end()); n.push_back((int32_t) c.size()); }
        pcm.assign(chunks.size(), nullptr);
        ns.assign(chunks.size(), 0);
        if (const char * dump = getenv("B2TTS_DUMP_TOKENS")) {      // parity tests: the token ids the reference's front end produced for this call, one chunk per line
            if (FILE * f = fopen(dump, "a")) { for (auto & c : chunks) { for (uint32_t t : c) fprintf(f, "%u ", t); fprintf(f, "\n"); } fclose(f); }
        }
        if (b2tts_kokoro_run_chunks(model, (int) chunks.size(), toks.data(), n.data(), voice.c_str(), noise_draws, pcm.data(), ns.data(), nullptr))
            TTS_ABORT("%s\n", b2tts_last_error());
        for (auto v : ns) noise_draws += 9ull * (uint64_t) v;
    }

    // the reference's flow from text to chunks (kokoro_runner::generate, model.cpp:1409-1450)
    void chunks_of(const char * prompt, std::vector<std::vector<uint32_t>> & chunks) {
        std::string normalized = replace_any(prompt, "\n", " ");
        std::string phonemes   = phmzr->text_to_phonemes(normalized);
        std::vector<uint32_t> space;
        tokenizer->tokenize(" ", space);
        const uint32_t space_id = space.empty() ? 0xffffffffu : space[0];
        if (phonemes.size() < (size_t) max_context_length - 2) {
            phonemes = strip(replace_any(phonemes, ".!?", ""));
            if (!phonemes.empty()) chunk_clause(phonemes, space_id, chunks);
        } else {
            for (auto clause : split(phonemes, ".!?")) {
                clause = strip(clause);
                if (!clause.empty()) chunk_clause(clause, space_id, chunks);
            }
        }
    }

    void generate(const char * prompt, tts_response & response, const generation_configuration & config) override {
        const std::string voice = config.voice.empty() ? "af_heart" : config.voice;
        std::vector<std::vector<uint32_t>> chunks;
        chunks_of(prompt, chunks);
        if (chunks.empty()) return;
        std::vector<const float *> pcm; std::vector<int64_t> ns;
        run_chunks(chunks, voice, pcm, ns);
        size_t total = 0;
        for (auto v : ns) total += (size_t) v;
        if (chunks.size() == 1) {
            response.data = const_cast<float *>(pcm[0]);      // borrowed, valid until the next generate (like kokoro_runner, model.cpp:1299)
        } else {
            joined.clear();
            for (size_t b = 0; b < chunks.size(); b++) joined.insert(joined.end(), pcm[b], pcm[b] + ns[b]);
            response.data = joined.data();
        }
        response.n_outputs = total;
    }

    // Non-breaking addition (SURVEY 8b "Batch"): several prompts in ONE forward, with the outputs the reference would give for generate(prompts[0]), generate(prompts[1]), ...
    // in this order on this runner (the noise stream runs through all chunks of all prompts).  responses[i].data points into a runner-owned buffer valid until the
    // next call; a server worker drains its queue into one call of this.
    void generate_batch(const std::vector<const char *> & prompts, std::vector<tts_response> & responses, const generation_configuration & config) {
        const std::string voice = config.voice.empty() ? "af_heart" : config.voice;
        std::vector<std::vector<uint32_t>> chunks;
        std::vector<size_t> first(prompts.size() + 1, 0);
        for (size_t i = 0; i < prompts.size(); i++) { chunks_of(prompts[i], chunks); first[i + 1] = chunks.size(); }
        responses.assign(prompts.size(), tts_response{});
        if (chunks.empty()) return;
        std::vector<const float *> pcm; std::vector<int64_t> ns;
        run_chunks(chunks, voice, pcm, ns);
        size_t total = 0;
        for (auto v : ns) total += (size_t) v;
        joined.clear(); joined.reserve(total);
        std::vector<size_t> at(prompts.size() + 1, 0);
        for (size_t i = 0; i < prompts.size(); i++) {
            for (size_t b = first[i]; b < first[i + 1]; b++) joined.insert(joined.end(), pcm[b], pcm[b] + ns[b]);
            at[i + 1] = joined.size();
        }
        for (size_t i = 0; i < prompts.size(); i++) { responses[i].data = joined.data() + at[i]; responses[i].n_outputs = at[i + 1] - at[i]; }
    }
};

// registers under the architecture string of the GGML loader: the registry is an emplace (first registration wins, loaders.cpp:11-31),
// so this object has to be linked ahead of the stock kokoro loader or replace it when TTS_B200 is on
struct kokoro_b200_loader final : tts_model_loader {
    kokoro_b200_loader() : tts_model_loader{ "kokoro" } {}
    unique_ptr<tts_generation_runner> from_file(gguf_context * meta, ggml_context *, int, bool, const generation_configuration & config) const override {
        auto r = make_unique<kokoro_b200_runner>(*this, single_pass_tokenizer_from_gguf(meta, "tokenizer.ggml.tokens"), phonemizer_from_gguf(meta, config.espeak_voice_id));
        std::vector<const char *> keys;
        std::vector<uint32_t>     vals;
        for (int i = 0; i < gguf_get_n_kv(meta); i++) {
            if (gguf_get_kv_type(meta, i) != GGUF_TYPE_UINT32) continue;
            keys.push_back(gguf_get_key(meta, i));
            vals.push_back(gguf_get_val_u32(meta, i));
            if (!strcmp(keys.back(), "kokoro.duration_predictor.albert.context_length")) r->max_context_length = vals.back();
        }
        if (b2tts_kokoro_create(r->ctx, (int) keys.size(), keys.data(), vals.data(), &r->model)) TTS_ABORT("%s\n", b2tts_last_error());
        return r;
    }
};

const kokoro_b200_loader kokoro_b200_loader_instance{};

}  // namespace

bool tts_b200_generate_batch(tts_generation_runner & runner, const std::vector<const char *> & prompts, std::vector<tts_response> & outputs, const generation_configuration & config) {
    auto * k = dynamic_cast<kokoro_b200_runner *>(&runner);
    if (!k) return false;
    k->generate_batch(prompts, outputs, config);
    return true;
}
