// oracle/ref_dia_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED reference Dia decode loop (dia_runner::decode: encoder pass + CFG-paired decoder step, and the token loop of
// generate_from_batch with check_stopping, reference src/models/dia/model.cpp:324-637,705-870) below the tokenizer: byte tokens in, N greedy
// steps (9 codebook tokens each) and the CFG-combined logits of every step out.  Loading follows dia_model_loader::from_file and
// runner_from_file's weight loop (src/models/dia/loader.cpp:8-22, src/models/loaders.cpp:79-89).
//
// usage: dia_ref <model.gguf> <prompts.txt> <out_prefix> [--steps N] [--threads T] [--quiet]
//   prompts.txt : one prompt per line, space separated byte tokens (the reference maps characters to their byte value, [S1]/[S2] to 1/2)
#include "models/dia/model.h"
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
    if (argc < 4) { fprintf(stderr, "usage: dia_ref <model.gguf> <prompts.txt> <out_prefix> [--steps N] [--threads T] [--quiet]\n"); return 2; }
    int threads = 4, steps = 5; bool quiet = false;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
    }
    ggml_context * weight_ctx = nullptr;
    gguf_init_params gp; gp.no_alloc = false; gp.ctx = &weight_ctx;
    gguf_context * meta = gguf_init_from_file(argv[1], gp);
    if (!meta) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }

    dia_model * model = new dia_model;
    dac_model * audio_model = new dac_model;
    model->setup_from_file(meta, weight_ctx, true);
    audio_model->setup_from_file(meta, weight_ctx, true);
    sampler * samp = new sampler;
    dac_context * dacctx = build_new_dac_context(audio_model, threads, true);
    dac_runner * audio_decoder = new dac_runner(audio_model, dacctx);
    dia_context * dctx = build_new_dia_context(model, threads, true);
    dia_kv_cache * cache = new dia_kv_cache;
    dia_runner * runner = new dia_runner(model, audio_decoder, dctx, samp, cache);
    for (ggml_tensor * cur = ggml_get_first_tensor(weight_ctx); cur; cur = ggml_get_next_tensor(weight_ctx, cur)) {
        if (!cur->data || !*cur->name) continue;
        runner->assign_weight(cur->name, *cur);
    }
    runner->prepare_post_load();
    samp->do_sample = false;
    samp->repetition_penalty = 1.0f;
    const uint32_t H = model->n_output_heads, V = model->output_vocab_size, C = model->max_encoder_context_length;

    std::ifstream in(argv[2]);
    std::string line; int u = 0; double wall_s = 0;
    while (std::getline(in, line)) {
        std::stringstream ss(line); std::vector<uint32_t> toks; uint32_t v;
        while (ss >> v) toks.push_back(v);
        if (toks.empty() || toks.size() > C) continue;
        dctx->reset();
        samp->reset();
        dctx->current_position = 0;
        dctx->max_generation_size = model->max_generation_size;
        dia_ubatch batch{1, true};                    // batch_from_sentence (model.cpp:684-694) with explicit byte tokens
        batch.tokens = toks;
        batch.sentence_length = toks.size();
        batch.tokens.resize((size_t) C * 2, 0u);      // conditional prompt padded to C, then the all-pad unconditional sequence
        for (uint32_t i = 0; i < H; i++) batch.audio_tokens.push_back(model->bos_token_id);
        std::vector<float> all_logits;
        auto t0 = clk::now();
        int done = 0;
        while (done < steps && !runner->check_stopping(batch)) {      // generate_from_batch's loop (model.cpp:849-864) with a step cap
            if (runner->decode(batch)) return 3;
            const float * lg = dctx->logits + (size_t) dctx->current_position * H * V;
            all_logits.insert(all_logits.end(), lg, lg + (size_t) H * V);
            samp->sample(dctx->logits + (size_t) dctx->current_position * H * V, dctx->output_tokens);
            dctx->current_position += batch.sequence_length;
            batch = dia_ubatch{1};
            uint32_t * last = dctx->output_tokens.data() + (int) dctx->output_tokens.size() - (int) H;
            batch.audio_tokens.reserve(H);
            for (uint32_t i = 0; i < H; i++) batch.audio_tokens.push_back(dctx->current_position > i ? last[i] : model->bos_token_id);
            done++;
        }
        wall_s += std::chrono::duration<double>(clk::now() - t0).count();
        std::vector<int32_t> out(dctx->output_tokens.begin(), dctx->output_tokens.end());
        FILE * f = fopen((std::string(argv[3]) + ".u" + std::to_string(u) + ".tokens.i32").c_str(), "wb");
        fwrite(out.data(), 4, out.size(), f); fclose(f);
        f = fopen((std::string(argv[3]) + ".u" + std::to_string(u) + ".logits.f32").c_str(), "wb");
        fwrite(all_logits.data(), 4, all_logits.size(), f); fclose(f);
        if (!quiet) { printf("UTT %d prompt %zu steps %d ->", u, toks.size(), done); for (auto t : out) printf(" %d", t); printf("\n"); }
        u++;
    }
    printf("SUMMARY {\"utterances\": %d, \"steps\": %d, \"wall_s\": %.6f, \"threads\": %d, \"heads\": %u, \"vocab\": %u}\n", u, steps, wall_s, threads, H, V);
    return 0;
}
