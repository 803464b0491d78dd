// integration/batch_demo.cpp -- a caller of the drop-in library exactly as examples/cli is (runner_from_file, generation_configuration), plus the one API addition:
// all prompts of a text file (one per line) through tts_b200_generate_batch in ONE batched forward, then the same prompts one by one through generate() on a second
// runner; writes <out>.batch.<i>.f32 / <out>.single.<i>.f32 (raw float PCM) so that a test can compare them.
//   batch_demo <model.gguf> <prompts.txt> <out_prefix>
#include "models/loaders.h"
#include "tts_b200.h"

#include <cstdio>
#include <fstream>
#include <string>

static void dump(const std::string & path, const tts_response & r) {
    FILE * f = fopen(path.c_str(), "wb");
    if (f) { fwrite(r.data, sizeof(float), r.n_outputs, f); fclose(f); }
}

int main(int argc, char ** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s model.gguf prompts.txt out_prefix\n", argv[0]); return 2; }
    std::vector<std::string> lines;
    { std::ifstream in(argv[2]); std::string l; while (std::getline(in, l)) if (!l.empty()) lines.push_back(l); }
    std::vector<const char *> prompts;
    for (auto & l : lines) prompts.push_back(l.c_str());
    const generation_configuration config{ "af_heart", 1, 1.0f, 1.0f, true, "", 0, 1.0f };
    auto a = runner_from_file(argv[1], 4, config, true);
    std::vector<tts_response> outs;
    if (!tts_b200_generate_batch(*a, prompts, outs, config)) { fprintf(stderr, "not a B200 runner\n"); return 1; }
    for (size_t i = 0; i < outs.size(); i++) dump(std::string(argv[3]) + ".batch." + std::to_string(i) + ".f32", outs[i]);
    auto b = runner_from_file(argv[1], 4, config, true);      // a second runner = a fresh noise stream, like a fresh reference process
    for (size_t i = 0; i < prompts.size(); i++) {
        tts_response r{};
        b->generate(prompts[i], r, config);
        dump(std::string(argv[3]) + ".single." + std::to_string(i) + ".f32", r);
    }
    printf("batch_demo: %zu prompts, %s\n", prompts.size(), a->loader.get().arch);
    (void) a.release(); (void) b.release();                   // like the reference's own tools (examples/cli/cli.cpp:97)
    return 0;
}
