// tests/emu/orpheus_main.cpp -- TEST INFRASTRUCTURE ONLY: drive Orpheus::generate_greedy (tts_cpp_b200/csrc/orpheus.cu, compiled against the
// CPU emulation in tests/emu/include) on the prompts of tests/golden/orpheus_vectors.npz and dump token ids + logits for the Python test.
//   orpheus_emu <model.gguf> <prompts.bin> <out.bin> [--single u]
// prompts.bin: int32 B, int32 n_steps, then per prompt int32 n, n x uint32.   out.bin: int32 tokens [B][n_steps], float logits [B][n_steps][vocab]
#include "orpheus.h"
#include <cstdio>
#include <vector>

namespace b2 { const char * emu_last_error(); }

int main(int argc, char ** argv) {
    if (argc < 4) return 2;
    b2::Ctx ctx;
    b2::Orpheus m; m.ctx = &ctx;
    if (b2::load_gguf_into(&m, argv[1])) { fprintf(stderr, "load: %s\n", b2::emu_last_error()); return 1; }
    FILE * f = fopen(argv[2], "rb");
    if (!f) return 2;
    int32_t B = 0, steps = 0;
    if (fread(&B, 4, 1, f) != 1 || fread(&steps, 4, 1, f) != 1) return 2;
    std::vector<std::vector<uint32_t>> pr((size_t) B);
    std::vector<const uint32_t *> pp; std::vector<int32_t> np;
    for (auto & p : pr) { int32_t n = 0; if (fread(&n, 4, 1, f) != 1) return 2; p.resize((size_t) n); if (fread(p.data(), 4, (size_t) n, f) != (size_t) n) return 2; pp.push_back(p.data()); np.push_back(n); }
    fclose(f);
    std::vector<int32_t> tok((size_t) B * steps);
    std::vector<float> logits((size_t) B * steps * m.vocab);
    if (m.generate_greedy(B, pp.data(), np.data(), steps, tok.data(), logits.data())) { fprintf(stderr, "generate: %s\n", b2::emu_last_error()); return 1; }
    f = fopen(argv[3], "wb");
    fwrite(tok.data(), 4, tok.size(), f); fwrite(logits.data(), 4, logits.size(), f);
    fclose(f);
    fprintf(stderr, "emulated %llu launches, %llu blocks\n", (unsigned long long) b2emu::g_launches, (unsigned long long) b2emu::g_blocks);
    return 0;
}
