// tests/emu/sampler_main.cpp -- TEST INFRASTRUCTURE ONLY: drive sample_rows (tts_cpp_b200/csrc/sampler.cu compiled against the CPU emulation) for a few
// consecutive steps on fixed logits and dump tokens + the uniforms used + the repetition state.
//   sampler_emu <in.bin> <out.bin>
// in : i32 rows, i32 V, i32 do_sample, i32 top_k, f32 top_p, f32 temperature, f32 rp, u64 seed, i32 steps, i32 last[rows], i32 counts[rows], f32 logits[steps][rows][V]
// out: i32 tokens[steps][rows], f32 uniforms[steps][rows], i32 last[rows], i32 counts[rows]
#include "kernels.cuh"
#include <cstdio>
#include <vector>

namespace b2 { const char * emu_last_error(); }

int main(int argc, char ** argv) {
    if (argc < 3) return 2;
    FILE * f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t rows, V, do_sample, top_k, steps; float top_p, temperature, rp; uint64_t seed;
    if (fread(&rows, 4, 1, f) != 1 || fread(&V, 4, 1, f) != 1 || fread(&do_sample, 4, 1, f) != 1 || fread(&top_k, 4, 1, f) != 1 || fread(&top_p, 4, 1, f) != 1 ||
        fread(&temperature, 4, 1, f) != 1 || fread(&rp, 4, 1, f) != 1 || fread(&seed, 8, 1, f) != 1 || fread(&steps, 4, 1, f) != 1) return 2;
    std::vector<int> last((size_t) rows), counts((size_t) rows);
    std::vector<float> logits((size_t) steps * rows * V);
    if (fread(last.data(), 4, (size_t) rows, f) != (size_t) rows || fread(counts.data(), 4, (size_t) rows, f) != (size_t) rows || fread(logits.data(), 4, logits.size(), f) != logits.size()) return 2;
    fclose(f);
    b2::Ctx ctx;
    std::vector<float> scratch((size_t) rows * V);
    std::vector<int> out((size_t) steps * rows);
    std::vector<float> us((size_t) steps * rows);
    int step = 0;
    for (step = 0; step < steps; step++) {
        b2::SampleParams p;
        p.logits = logits.data() + (size_t) step * rows * V; p.rows = rows; p.V = V; p.do_sample = do_sample; p.top_k = top_k; p.top_p = top_p; p.temperature = temperature;
        p.repetition_penalty = rp; p.last_ids = last.data(); p.rep_counts = counts.data(); p.scratch = scratch.data(); p.seed = seed; p.d_step = &step; p.out = out.data();
        if (b2::sample_rows(&ctx, p)) { fprintf(stderr, "sample_rows: %s\n", b2::emu_last_error()); return 1; }
        for (int r = 0; r < rows; r++) us[(size_t) step * rows + r] = b2::sample_uniform_host(seed, (unsigned long long) r, (unsigned long long) step);
    }
    f = fopen(argv[2], "wb");
    fwrite(out.data(), 4, out.size(), f); fwrite(us.data(), 4, us.size(), f); fwrite(last.data(), 4, last.size(), f); fwrite(counts.data(), 4, counts.size(), f);
    fclose(f);
    return 0;
}
