// tests/emu/vad_main.cpp -- TEST INFRASTRUCTURE ONLY: drive vad_trim_rows (tts_cpp_b200/csrc/vad.cu compiled against the CPU emulation) on given PCM.
//   vad_emu <in.bin> <out.bin>      (the file formats of oracle/ref_vad_driver.cpp: the same input file goes to the compiled reference)
// in : u32 B, f32 sample_rate, i32 ms_per_frame, i32 frame_threshold, f32 normalized_energy_threshold, i32 trailing_silent_frames, i32 early_cutoff_seconds_threshold,
//      f32 early_cutoff_energy_threshold, i64 n[B], f32 pcm[sum n]
// out: i64 n_outputs[B], f32 energies[sum n_b / spf]
#include "kernels.cuh"
#include <cstdio>
#include <vector>

namespace b2 { const char * emu_last_error(); }

int main(int argc, char ** argv) {
    if (argc < 3) return 2;
    FILE * f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint32_t B; float sr, nthr, ethr; int32_t ms, fthr, trail, esec;
    if (fread(&B, 4, 1, f) != 1 || fread(&sr, 4, 1, f) != 1 || fread(&ms, 4, 1, f) != 1 || fread(&fthr, 4, 1, f) != 1 || fread(&nthr, 4, 1, f) != 1 || fread(&trail, 4, 1, f) != 1 ||
        fread(&esec, 4, 1, f) != 1 || fread(&ethr, 4, 1, f) != 1) return 2;
    std::vector<long long> n(B);
    if (fread(n.data(), 8, B, f) != B) return 2;
    const int spf = (int) (ms * sr / 1000.0f), early = (int) ((esec * 1000) / ms);
    std::vector<long long> off(B + 1, 0), eoff(B + 1, 0);
    int max_frames = 0;
    for (uint32_t b = 0; b < B; b++) { off[b + 1] = off[b] + n[b]; eoff[b + 1] = eoff[b] + n[b] / spf; if (n[b] / spf > max_frames) max_frames = (int) (n[b] / spf); }
    std::vector<float> pcm((size_t) off[B]);                    // exact-size allocations: AddressSanitizer sees any read past an utterance
    if (fread(pcm.data(), 4, pcm.size(), f) != pcm.size()) return 2;
    fclose(f);
    std::vector<float> en((size_t) eoff[B]);
    std::vector<long long> out(B);
    b2::Ctx ctx;
    if (b2::vad_trim_rows(&ctx, pcm.data(), off.data(), eoff.data(), (int) B, max_frames, spf, fthr, nthr, trail, early, ethr, en.data(), out.data())) {
        fprintf(stderr, "vad_trim_rows: %s\n", b2::emu_last_error());
        return 1;
    }
    f = fopen(argv[2], "wb");
    fwrite(out.data(), 8, B, f); fwrite(en.data(), 4, en.size(), f);
    fclose(f);
    return 0;
}
