// oracle/ref_snac_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED reference SNAC codec decoder (snac_runner, reference src/decoder/snac_model.cpp:86-208 and
// src/decoder/general_neural_audio_codec.cpp:133-172), compiled by oracle/Makefile from /root/reference, on explicit
// codebook indices and writes the PCM as raw float32.  Loading follows the reference's Orpheus loader for its audio decoder
// (src/models/orpheus/loader.cpp:12-18, src/models/orpheus/model.cpp:440-441).  The noise the decoder injects comes from the
// reference's process-wide std::normal_distribution (src/util.cpp:74-80); utterances are decoded in order in ONE process, so
// utterance k sees the draws after those of utterances 0..k-1.
//
// usage: snac_ref <model.gguf> <codes.txt> <out_prefix> [--threads N] [--quiet]
//   codes.txt : one utterance per line: L/4 coarse, then L/2 medium, then L fine indices (the three streams snac_runner::run takes)
#include "decoder/snac_model.h"
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
    if (argc < 4) { fprintf(stderr, "usage: snac_ref <model.gguf> <codes.txt> <out_prefix> [--threads N] [--quiet]\n"); return 2; }
    int threads = 4; bool quiet = false;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
    }
    ggml_context * weight_ctx = nullptr;
    gguf_init_params gp; gp.no_alloc = false; gp.ctx = &weight_ctx;
    gguf_context * meta = gguf_init_from_file(argv[1], gp);
    if (!meta) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }

    snac_model * model = new snac_model;
    model->setup_from_file(meta, weight_ctx, true);
    snac_context * sctx = build_new_snac_context(model, threads, true);
    snac_runner * runner = new snac_runner(model, sctx);
    const std::string prefix = "snac.";
    for (ggml_tensor * cur = ggml_get_first_tensor(weight_ctx); cur; cur = ggml_get_next_tensor(weight_ctx, cur)) {
        if (!cur->data || !*cur->name) continue;
        const std::string name = cur->name;
        if (name.compare(0, prefix.size(), prefix) == 0) model->assign_weight(name.substr(prefix.size()), cur);
    }
    runner->prepare_post_load();

    std::ifstream in(argv[2]);
    std::string line; int u = 0; double audio_s = 0, wall_s = 0;
    while (std::getline(in, line)) {
        std::stringstream ss(line); std::vector<uint32_t> all; uint32_t v;
        while (ss >> v) all.push_back(v);
        if (all.empty()) continue;
        // all = L/4 + L/2 + L entries  ->  L = 4 * n / 7
        const size_t L = all.size() * 4 / 7;
        std::vector<std::vector<uint32_t>> toks(3);
        toks[0].assign(all.begin(), all.begin() + L / 4);
        toks[1].assign(all.begin() + L / 4, all.begin() + L / 4 + L / 2);
        toks[2].assign(all.begin() + L / 4 + L / 2, all.end());
        tts_response resp; resp.data = nullptr; resp.n_outputs = 0;
        auto t0 = clk::now();
        runner->run(toks, &resp);
        wall_s += std::chrono::duration<double>(clk::now() - t0).count();
        audio_s += (double) resp.n_outputs / 24000.0;
        FILE * f = fopen((std::string(argv[3]) + ".u" + std::to_string(u) + ".pcm.f32").c_str(), "wb");
        if (!f) { fprintf(stderr, "cannot write output\n"); return 2; }
        fwrite(resp.data, sizeof(float), resp.n_outputs, f);
        fclose(f);
        if (!quiet) printf("UTT %d fine_frames %zu samples %zu\n", u, L, (size_t) resp.n_outputs);
        u++;
    }
    printf("SUMMARY {\"utterances\": %d, \"audio_s\": %.6f, \"wall_s\": %.6f, \"threads\": %d}\n", u, audio_s, wall_s, threads);
    return 0;
}
