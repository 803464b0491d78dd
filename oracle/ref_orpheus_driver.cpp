// oracle/ref_orpheus_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED reference Orpheus decode loop (orpheus_runner::decode + sampler, reference
// src/models/orpheus/model.cpp:230-353,389-405, src/sampler.cpp) below the tokenizer: prompt token ids in, greedy (argmax) continuation and
// the logits of every step out.  Loading follows orpheus_model_loader::from_file + runner_from_file's weight loop
// (src/models/orpheus/loader.cpp:8-23, src/models/loaders.cpp:79-89) with no tokenizer (never used below batch_from_sentence).
//
// usage: orpheus_ref <model.gguf> <prompts.txt> <out_prefix> [--steps N] [--threads T] [--quiet]
//   prompts.txt : one prompt per line, space separated token ids
//   writes <out_prefix>.u<k>.tokens.i32 (the N generated ids) and <out_prefix>.u<k>.logits.f32 ([N][vocab])
#include "models/orpheus/model.h"
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
    if (argc < 4) { fprintf(stderr, "usage: orpheus_ref <model.gguf> <prompts.txt> <out_prefix> [--steps N] [--threads T] [--quiet]\n"); return 2; }
    int threads = 4, steps = 8; bool quiet = false;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
    }
    ggml_context * weight_ctx = nullptr;
    gguf_init_params gp; gp.no_alloc = false; gp.ctx = &weight_ctx;
    gguf_context * meta = gguf_init_from_file(argv[1], gp);
    if (!meta) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }

    orpheus_model * model = new orpheus_model;
    snac_model * audio_model = new snac_model;
    model->setup_from_file(meta, weight_ctx, true);
    audio_model->setup_from_file(meta, weight_ctx, true);
    sampler * samp = new sampler;
    snac_context * sctx = build_new_snac_context(audio_model, threads, true);
    snac_runner * audio_decoder = new snac_runner(audio_model, sctx);
    orpheus_context * octx = build_new_orpheus_context(model, threads, true);
    orpheus_kv_cache * cache = new orpheus_kv_cache;
    orpheus_runner * runner = new orpheus_runner(model, audio_decoder, octx, nullptr, samp, cache);
    for (ggml_tensor * cur = ggml_get_first_tensor(weight_ctx); cur; cur = ggml_get_next_tensor(weight_ctx, cur)) {
        if (!cur->data || !*cur->name) continue;
        runner->assign_weight(cur->name, *cur);
    }
    runner->prepare_post_load();
    samp->do_sample = false;           // greedy: sampler::max, first maximum wins (src/sampler.cpp)
    samp->repetition_penalty = 1.0f;

    std::ifstream in(argv[2]);
    std::string line; int u = 0; double wall_s = 0; long n_steps = 0;
    while (std::getline(in, line)) {
        std::stringstream ss(line); std::vector<uint32_t> toks; uint32_t v;
        while (ss >> v) toks.push_back(v);
        if (toks.empty()) continue;
        octx->reset();
        samp->reset();
        orpheus_ubatch batch(toks.size(), toks);
        std::vector<float> all_logits;
        auto t0 = clk::now();
        for (int s = 0; s < steps; s++) {          // generate_from_batch's loop (model.cpp:389-398) without the stop condition
            runner->decode(batch);
            const float * lg = octx->logits + (size_t) octx->n_outputs * model->vocab_size;
            all_logits.insert(all_logits.end(), lg, lg + model->vocab_size);
            samp->sample(octx->logits + (size_t) octx->n_outputs * model->vocab_size, octx->output_tokens);
            octx->n_outputs++;
            batch = orpheus_ubatch{1, {octx->output_tokens.back()}};
        }
        wall_s += std::chrono::duration<double>(clk::now() - t0).count();
        n_steps += steps;
        std::vector<int32_t> out(octx->output_tokens.begin(), octx->output_tokens.end());
        FILE * f = fopen((std::string(argv[3]) + ".u" + std::to_string(u) + ".tokens.i32").c_str(), "wb");
        fwrite(out.data(), 4, out.size(), f); fclose(f);
        f = fopen((std::string(argv[3]) + ".u" + std::to_string(u) + ".logits.f32").c_str(), "wb");
        fwrite(all_logits.data(), 4, all_logits.size(), f); fclose(f);
        if (!quiet) { printf("UTT %d prompt %zu ->", u, toks.size()); for (auto t : out) printf(" %d", t); printf("\n"); }
        u++;
    }
    printf("SUMMARY {\"utterances\": %d, \"steps\": %ld, \"wall_s\": %.6f, \"threads\": %d, \"vocab\": %u}\n", u, n_steps, wall_s, threads, model->vocab_size);
    return 0;
}
