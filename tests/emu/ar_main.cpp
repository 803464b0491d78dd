// tests/emu/ar_main.cpp -- TEST INFRASTRUCTURE ONLY: drive a model's generate_greedy (tts_cpp_b200/csrc/{orpheus,parler,dia}.cu compiled against the
// CPU emulation in tests/emu/include) on the prompts of the golden vectors and dump token ids + logits for the Python test.
//   ar_emu <orpheus|parler|dia> <model.gguf> <prompts.bin> <out.bin>
// prompts.bin: int32 B, int32 n_steps, then per prompt int32 n, n x uint32.
// out.bin: int32 W (tokens per step), int32 V (logits per step), int32 tokens [B][n_steps][W], float logits [B][n_steps][V]
#include "orpheus.h"
#include "parler.h"
#include "dia.h"
#include "t5.h"
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <vector>

namespace b2 { const char * emu_last_error(); }

static int out_width(const b2::Orpheus &) { return 1; }
static int out_logits(const b2::Orpheus & m) { return m.vocab; }
template <class M> static int out_width(const M & m) { return m.n_out; }
template <class M> static int out_logits(const M & m) { return m.n_out * m.vocab; }

static std::vector<int32_t> g_ngen;      // B2EMU_STOP=1: the reference's stop rule; n_generated is appended to out.bin
template <class M> static int gen(M & m, int B, const uint32_t * const * pp, const int32_t * np, int steps, const b2::ArSampling * s, int32_t * tok, float * lg) { return m.generate(B, pp, np, steps, s, tok, lg); }
static int gen(b2::Parler & m, int B, const uint32_t * const * pp, const int32_t * np, int steps, const b2::ArSampling * s, int32_t * tok, float * lg) {
    if (getenv("B2EMU_STOP")) g_ngen.assign((size_t) B, 0);
    if (const char * ef = getenv("B2EMU_ENCODING")) {             // "<f32 file> <rows>": replace the stored conditional-prompt encoding first
        char path[512]; int rows = 0;
        if (sscanf(ef, "%511s %d", path, &rows) != 2) return 2;
        std::vector<float> enc((size_t) rows * m.hidden);
        FILE * f = fopen(path, "rb");
        if (!f || fread(enc.data(), 4, enc.size(), f) != enc.size()) return 2;
        fclose(f);
        if (m.set_text_encoding(enc.data(), rows)) return 1;
    }
    std::vector<int32_t> teacher;                                  // B2EMU_TEACHER=<file of int32 [B][steps][n_out]>: teacher-forced feedback
    if (const char * tf = getenv("B2EMU_TEACHER")) {
        teacher.resize((size_t) B * steps * m.n_out);
        FILE * f = fopen(tf, "rb");
        if (!f || fread(teacher.data(), 4, teacher.size(), f) != teacher.size()) return 2;
        fclose(f);
    }
    return m.generate(B, pp, np, steps, s, tok, lg, g_ngen.empty() ? nullptr : g_ngen.data(), teacher.empty() ? nullptr : teacher.data());
}
static int gen(b2::Dia & m, int B, const uint32_t * const * pp, const int32_t * np, int steps, const b2::ArSampling * s, int32_t * tok, float * lg) {
    std::vector<int32_t> teacher;                                  // B2EMU_TEACHER=<file of int32 [B][steps][n_out]>: teacher-forced feedback
    if (const char * tf = getenv("B2EMU_TEACHER")) {
        teacher.resize((size_t) B * steps * m.n_out);
        FILE * f = fopen(tf, "rb");
        if (!f || fread(teacher.data(), 4, teacher.size(), f) != teacher.size()) return 2;
        fclose(f);
    }
    if (getenv("B2EMU_STOP")) g_ngen.assign((size_t) B, 0);
    return m.generate(B, pp, np, steps, s, tok, lg, g_ngen.empty() ? nullptr : g_ngen.data(), teacher.empty() ? nullptr : teacher.data());
}

template <class M> static int run(int argc, char ** argv) {
    b2::Ctx ctx;
    M m; m.ctx = &ctx;
    if (b2::load_gguf_into(&m, argv[2])) { fprintf(stderr, "load: %s\n", b2::emu_last_error()); return 1; }
    FILE * f = fopen(argv[3], "rb");
    if (!f) return 2;
    int32_t B = 0, steps = 0;
    if (fread(&B, 4, 1, f) != 1 || fread(&steps, 4, 1, f) != 1) return 2;
    std::vector<std::vector<uint32_t>> pr((size_t) B);
    std::vector<const uint32_t *> pp; std::vector<int32_t> np;
    for (auto & p : pr) { int32_t n = 0; if (fread(&n, 4, 1, f) != 1) return 2; p.resize((size_t) n); if (fread(p.data(), 4, (size_t) n, f) != (size_t) n) return 2; pp.push_back(p.data()); np.push_back(n); }
    fclose(f);
    const int32_t W = out_width(m), V = out_logits(m);
    std::vector<int32_t> tok((size_t) B * steps * W);
    std::vector<float> logits((size_t) B * steps * V);
    b2::ArSampling samp;                                          // B2EMU_SAMPLE="top_k top_p temperature repetition_penalty seed": the stochastic sampler
    if (const char * e = getenv("B2EMU_SAMPLE")) {
        unsigned long long sd = 0;
        if (sscanf(e, "%d %f %f %f %llu", &samp.top_k, &samp.top_p, &samp.temperature, &samp.repetition_penalty, &sd) != 5) return 2;
        samp.seed = sd; samp.do_sample = 1;
    }
    const bool want_logits = !getenv("B2EMU_NO_LOGITS");      // without logits the decode loops may replay a captured CUDA graph (B2TTS_AR_GRAPH=1)
    if (gen(m, B, pp.data(), np.data(), steps, &samp, tok.data(), want_logits ? logits.data() : nullptr)) { fprintf(stderr, "generate: %s\n", b2::emu_last_error()); return 1; }
    f = fopen(argv[4], "wb");
    fwrite(&W, 4, 1, f); fwrite(&V, 4, 1, f);
    fwrite(tok.data(), 4, tok.size(), f); fwrite(logits.data(), 4, logits.size(), f);
    if (!g_ngen.empty()) fwrite(g_ngen.data(), 4, g_ngen.size(), f);
    fclose(f);
    fprintf(stderr, "emulated %llu launches, %llu blocks, %llu graph replays\n", (unsigned long long) b2emu::g_launches, (unsigned long long) b2emu::g_blocks, (unsigned long long) b2emu::g_replays);
    return 0;
}

// ar_emu t5 <t5.gguf> <prompts.bin> <out.bin>: prompts.bin as above (n_steps ignored); out.bin: int32 output_size, then the encodings back to back
static int run_t5(char ** argv) {
    b2::Ctx ctx;
    b2::T5 m; m.ctx = &ctx;
    if (b2::load_gguf_into(&m, argv[2])) { fprintf(stderr, "load: %s\n", b2::emu_last_error()); return 1; }
    FILE * f = fopen(argv[3], "rb");
    if (!f) return 2;
    int32_t B = 0, steps = 0;
    if (fread(&B, 4, 1, f) != 1 || fread(&steps, 4, 1, f) != 1) return 2;
    std::vector<std::vector<uint32_t>> pr((size_t) B);
    std::vector<const uint32_t *> pp; std::vector<int32_t> np;
    size_t rows = 0;
    for (auto & p : pr) { int32_t n = 0; if (fread(&n, 4, 1, f) != 1) return 2; p.resize((size_t) n); if (fread(p.data(), 4, (size_t) n, f) != (size_t) n) return 2; pp.push_back(p.data()); np.push_back(n); rows += (size_t) n; }
    fclose(f);
    const int32_t O = m.output_size();
    std::vector<float> enc(rows * (size_t) O);                // exact size: AddressSanitizer sees a row too many
    if (m.encode(B, pp.data(), np.data(), enc.data())) { fprintf(stderr, "encode: %s\n", b2::emu_last_error()); return 1; }
    f = fopen(argv[4], "wb");
    fwrite(&O, 4, 1, f); fwrite(enc.data(), 4, enc.size(), f);
    fclose(f);
    fprintf(stderr, "emulated %llu launches, %llu blocks\n", (unsigned long long) b2emu::g_launches, (unsigned long long) b2emu::g_blocks);
    return 0;
}

int main(int argc, char ** argv) {
    if (argc < 5) return 2;
    if (!strcmp(argv[1], "t5")) return run_t5(argv);
    if (!strcmp(argv[1], "orpheus")) return run<b2::Orpheus>(argc, argv);
    if (!strcmp(argv[1], "parler")) return run<b2::Parler>(argc, argv);
    if (!strcmp(argv[1], "dia")) return run<b2::Dia>(argc, argv);
    return 2;
}
