// oracle/ref_parler_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED reference Parler-TTS decode loop (parler_tts_runner::decode + sampler with the delay pattern of
// generate_from_batch, reference src/models/parler/model.cpp:387-470,520-693,762-792) below the tokenizer: prompt token ids in, N greedy
// audio steps (9 codebook tokens each) and the logits of every step out.  Loading follows parler_model_loader::from_file and
// runner_from_file's weight loop (src/models/parler/loader.cpp:8-23, src/models/loaders.cpp:79-89) without a tokenizer.
//
// usage: parler_ref <model.gguf> <prompts.txt> <out_prefix> [--steps N] [--threads T] [--quiet] [--stop] [--encoding f32file rows]
//   writes <out_prefix>.u<k>.tokens.i32 ([N][heads] generated ids) and <out_prefix>.u<k>.logits.f32 ([N][heads][output_vocab])
#include "models/parler/model.h"
#include "ggml.h"
#include "ggml-backend.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

using clk = std::chrono::steady_clock;

int main(int argc, char ** argv) {
    if (argc < 4) { fprintf(stderr, "usage: parler_ref <model.gguf> <prompts.txt> <out_prefix> [--steps N] [--threads T] [--quiet] [--stop] [--encoding f32file rows]\n"); return 2; }
    int threads = 4, steps = 6; bool quiet = false, use_stop = false;
    const char * enc_file = nullptr; int enc_rows = 0;      // --encoding <f32 file> <rows>: a replacement conditional-prompt encoding (what update_conditional_prompt's T5 pass yields)
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
        else if (!strcmp(argv[i], "--stop")) use_stop = true;
        else if (!strcmp(argv[i], "--encoding") && i + 2 < argc) { enc_file = argv[++i]; enc_rows = atoi(argv[++i]); }
    }
    ggml_context * weight_ctx = nullptr;
    gguf_init_params gp; gp.no_alloc = false; gp.ctx = &weight_ctx;
    gguf_context * meta = gguf_init_from_file(argv[1], gp);
    if (!meta) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }

    parler_tts_model * model = new parler_tts_model;
    dac_model * audio_model = new dac_model;
    model->use_cross_attn = true;
    model->setup_from_file(meta, weight_ctx, true);
    audio_model->setup_from_file(meta, weight_ctx, true);
    sampler * samp = new sampler;
    dac_context * dctx = build_new_dac_context(audio_model, threads, true);
    dac_runner * audio_decoder = new dac_runner(audio_model, dctx);
    parler_context * pctx = build_new_parler_context(model, threads, true);
    parler_kv_cache * cache = new parler_kv_cache;
    parler_tts_runner * runner = new parler_tts_runner(model, audio_decoder, pctx, nullptr, samp, cache);
    for (ggml_tensor * cur = ggml_get_first_tensor(weight_ctx); cur; cur = ggml_get_next_tensor(weight_ctx, cur)) {
        if (!cur->data || !*cur->name) continue;
        runner->assign_weight(cur->name, *cur);
    }
    runner->prepare_post_load();
    std::vector<float> enc_data;
    if (enc_file) {                                         // parler_tts_runner::update_conditional_prompt's second half (model.cpp:516) with a given encoding
        enc_data.resize((size_t) enc_rows * model->hidden_size);
        FILE * ef = fopen(enc_file, "rb");
        if (!ef || fread(enc_data.data(), 4, enc_data.size(), ef) != enc_data.size()) { fprintf(stderr, "cannot read %s\n", enc_file); return 2; }
        fclose(ef);
        tts_response resp; resp.data = enc_data.data(); resp.n_outputs = (size_t) enc_rows; resp.hidden_size = model->hidden_size;
        model->prep_cross_key_values(threads, &resp);
    }
    samp->do_sample = false;           // greedy: sampler::max per head, first maximum wins
    samp->repetition_penalty = 1.0f;
    const uint32_t H = model->n_output_heads, V = model->output_vocab_size;

    std::ifstream in(argv[2]);
    std::string line; int u = 0; double wall_s = 0;
    while (std::getline(in, line)) {
        std::stringstream ss(line); std::vector<uint32_t> toks; uint32_t v;
        while (ss >> v) toks.push_back(v);
        if (toks.empty()) continue;
        pctx->reset(H);
        samp->reset();
        pctx->current_position = 0;
        std::vector<uint32_t> positions(toks.size());
        for (size_t i = 0; i < toks.size(); i++) positions[i] = (uint32_t) i;
        parler_ubatch batch;                       // batch_from_sentence (model.cpp:473-498) with explicit ids
        batch.audio_generation = false; batch.current_step = 0; batch.n_tokens = toks.size(); batch.n_audio_tokens = 0;
        batch.sequence_length = toks.size(); batch.tokens = toks.data(); batch.positions = positions.data(); batch.audio_tokens = nullptr; batch.true_order = nullptr;
        std::vector<uint32_t> next_ids; next_ids.reserve(H);
        std::vector<float> all_logits;
        auto t0 = clk::now();
        int audio_steps = 0;
        while (audio_steps < steps && !(use_stop && runner->check_stopping())) {   // generate_from_batch's loop (model.cpp:762-786); --stop: with check_stopping, else a step cap only
            if (runner->decode(batch)) return 3;
            if (!batch.audio_generation) pctx->prompt_end_position += batch.sequence_length;
            if (batch.audio_generation) {
                const float * lg = pctx->logits + (size_t) pctx->current_position * H * V;
                all_logits.insert(all_logits.end(), lg, lg + (size_t) H * V);
                samp->sample(pctx->logits + (size_t) pctx->current_position * H * V, pctx->output_tokens);
                audio_steps++;
            }
            pctx->current_position += batch.sequence_length;
            next_ids.clear();
            uint32_t * last = pctx->output_tokens.data() + (int) pctx->output_tokens.size() - (int) H;
            for (uint32_t i = 0; i < H; i++) next_ids.push_back(batch.current_step > (int) i ? (pctx->eos_seen[i] ? model->eos_token_id : last[i]) : model->bos_token_id);
            batch = parler_ubatch{true, 0, H, 1, nullptr, next_ids.data(), &pctx->current_position, nullptr, batch.current_step + 1};
        }
        wall_s += std::chrono::duration<double>(clk::now() - t0).count();
        std::vector<int32_t> out(pctx->output_tokens.begin(), pctx->output_tokens.end());
        FILE * f = fopen((std::string(argv[3]) + ".u" + std::to_string(u) + ".tokens.i32").c_str(), "wb");
        fwrite(out.data(), 4, out.size(), f); fclose(f);
        f = fopen((std::string(argv[3]) + ".u" + std::to_string(u) + ".logits.f32").c_str(), "wb");
        fwrite(all_logits.data(), 4, all_logits.size(), f); fclose(f);
        if (!quiet) { printf("UTT %d prompt %zu ->", u, toks.size()); for (auto t : out) printf(" %d", t); printf("\n"); }
        u++;
    }
    printf("SUMMARY {\"utterances\": %d, \"steps\": %d, \"wall_s\": %.6f, \"threads\": %d, \"heads\": %u, \"vocab\": %u}\n", u, steps, wall_s, threads, H, V);
    return 0;
}
