// oracle/ref_t5_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED T5 conditional-prompt encoder of the reference (src/models/parler/t5/model.cpp: text_encoder_from_file -> t5_runner::run, the pass
// parler_tts_runner::update_conditional_prompt makes before prep_cross_key_values, src/models/parler/model.cpp:510-518) below its tokenizer, on explicit token ids.
// usage: t5_ref <t5.gguf> <out.bin> <n_threads> <tok0,tok1,...> [<tok0,...> ...]
//   out: per prompt: u32 n_tokens, u32 output_size, f32 encoding[n_tokens][output_size]
#include "models/parler/t5/model.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char ** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s t5.gguf out.bin n_threads tok,tok,... [...]\n", argv[0]); return 2; }
    // a tokenizer object that is never asked to tokenize: the driver enters at t5_runner::run (model.cpp:336)
    unigram_tokenizer * tok = new unigram_tokenizer({}, 0, 0.0f, {});
    tok->init = true;
    t5_runner * runner = text_encoder_from_file(argv[1], atoi(argv[3]), tok, true);
    FILE * f = fopen(argv[2], "wb");
    if (!f) return 2;
    for (int a = 4; a < argc; a++) {
        std::vector<uint32_t> ids;
        for (char * p = strtok(argv[a], ","); p; p = strtok(nullptr, ",")) ids.push_back((uint32_t) strtoul(p, nullptr, 10));
        tts_response r{};
        runner->run(ids.data(), (uint32_t) ids.size(), &r);
        const uint32_t n = (uint32_t) r.n_outputs, hs = (uint32_t) r.hidden_size;
        fwrite(&n, 4, 1, f); fwrite(&hs, 4, 1, f);
        fwrite(r.data, sizeof(float), (size_t) n * hs, f);
    }
    fclose(f);
    return 0;
}
