// oracle/ref_dac_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED reference DAC codec decoder (dac_runner, reference src/decoder/dac_model.cpp:146-212 and
// src/decoder/general_neural_audio_codec.cpp:133-172), compiled by oracle/Makefile from /root/reference, on explicit
// codebook indices and writes the PCM as raw float32.  The loading sequence is the one the reference's Parler / Dia loaders
// perform for their audio decoder (src/models/parler/loader.cpp:12-20, src/models/loaders.cpp:79-89): setup_from_file on the
// "audio_encoder." tensors, one assign_weight per tensor, prepare_post_load.  Only the reference's own public members are called.
//
// usage: dac_ref <model.gguf> <codes.txt> <out_prefix> [--threads N] [--reps R] [--quiet]
//   codes.txt : one utterance per line: frames * n_heads codebook indices, frame-major (the layout dac_runner::run takes)
#include "decoder/dac_model.h"
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
    if (argc < 4) { fprintf(stderr, "usage: dac_ref <model.gguf> <codes.txt> <out_prefix> [--threads N] [--reps R] [--quiet]\n"); return 2; }
    const char * path = argv[1];
    int threads = 4, reps = 1; bool quiet = false;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--reps") && i + 1 < argc) reps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
    }
    ggml_context * weight_ctx = nullptr;
    gguf_init_params gp; gp.no_alloc = false; gp.ctx = &weight_ctx;
    gguf_context * meta = gguf_init_from_file(path, gp);
    if (!meta) { fprintf(stderr, "cannot read %s\n", path); return 2; }

    dac_model * model = new dac_model;
    model->setup_from_file(meta, weight_ctx, true);
    dac_context * dctx = build_new_dac_context(model, threads, true);
    dac_runner * runner = new dac_runner(model, dctx);
    const std::string prefix = "audio_encoder.";
    for (ggml_tensor * cur = ggml_get_first_tensor(weight_ctx); cur; cur = ggml_get_next_tensor(weight_ctx, cur)) {
        if (!cur->data || !*cur->name) continue;
        const std::string name = cur->name;
        if (name.compare(0, prefix.size(), prefix) == 0) model->assign_weight(name.substr(prefix.size()), cur);
    }
    runner->prepare_post_load();

    std::ifstream in(argv[2]);
    std::string line; int u = 0; double audio_s = 0, wall_s = 0;
    while (std::getline(in, line)) {
        std::stringstream ss(line); std::vector<uint32_t> codes; uint32_t v;
        while (ss >> v) codes.push_back(v);
        if (codes.empty()) continue;
        const uint32_t frames = (uint32_t) (codes.size() / model->n_heads);
        tts_response resp; resp.data = nullptr; resp.n_outputs = 0;
        for (int r = 0; r < reps; r++) {
            auto t0 = clk::now();
            runner->run(codes.data(), frames, &resp);
            wall_s += std::chrono::duration<double>(clk::now() - t0).count();
            audio_s += (double) resp.n_outputs / 44100.0;
        }
        FILE * f = fopen((std::string(argv[3]) + ".u" + std::to_string(u) + ".pcm.f32").c_str(), "wb");
        if (!f) { fprintf(stderr, "cannot write output\n"); return 2; }
        fwrite(resp.data, sizeof(float), resp.n_outputs, f);
        fclose(f);
        if (!quiet) printf("UTT %d frames %u samples %zu\n", u, frames, (size_t) resp.n_outputs);
        u++;
    }
    printf("SUMMARY {\"utterances\": %d, \"audio_s\": %.6f, \"wall_s\": %.6f, \"threads\": %d}\n", u, audio_s, wall_s, threads);
    return 0;
}
