// oracle/ref_vad_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED apply_energy_voice_inactivity_detection of the reference (examples/cli/vad.cpp:11-68, compiled from where it lies) on given PCM.
// usage: vad_ref <in.bin> <out.bin>
//   in : u32 B, f32 sample_rate, i32 ms_per_frame, i32 frame_threshold, f32 normalized_energy_threshold, i32 trailing_silent_frames,
//        i32 early_cutoff_seconds_threshold, f32 early_cutoff_energy_threshold, i64 n[B], f32 pcm[sum n]
//   out: i64 n_outputs[B] (the trimmed lengths; two's complement of the reference's size_t), then per utterance the frame energies the reference's own
//        energy() (vad.cpp:3-9) returns for its n / samples_per_frame whole frames: f32 e[n_frames]
#include "vad.h"

#include <cstdint>
#include <cstdio>
#include <vector>

int main(int argc, char ** argv) {
    if (argc < 3) return 2;
    FILE * f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint32_t B; float sr, nthr, ethr; int32_t ms, fthr, trail, esec;
    if (fread(&B, 4, 1, f) != 1 || fread(&sr, 4, 1, f) != 1 || fread(&ms, 4, 1, f) != 1 || fread(&fthr, 4, 1, f) != 1 || fread(&nthr, 4, 1, f) != 1 || fread(&trail, 4, 1, f) != 1 ||
        fread(&esec, 4, 1, f) != 1 || fread(&ethr, 4, 1, f) != 1) return 2;
    std::vector<int64_t> n(B);
    if (fread(n.data(), 8, B, f) != B) return 2;
    size_t total = 0;
    for (auto v : n) total += (size_t) v;
    std::vector<float> pcm(total);
    if (fread(pcm.data(), 4, total, f) != total) return 2;
    fclose(f);
    std::vector<int64_t> out(B);
    std::vector<float> en;
    const int spf = (int) (ms * sr / 1000.0f);                                  // vad.cpp:20
    size_t at = 0;
    for (uint32_t b = 0; b < B; b++) {
        for (int i = 0; i < (int) (n[b] / spf); i++) en.push_back(energy(pcm.data() + at + (size_t) i * spf, spf));
        tts_response r{};
        r.data = pcm.data() + at; r.n_outputs = (size_t) n[b];
        apply_energy_voice_inactivity_detection(r, sr, ms, fthr, nthr, trail, esec, ethr);
        out[b] = (int64_t) r.n_outputs;
        at += (size_t) n[b];
    }
    f = fopen(argv[2], "wb");
    fwrite(out.data(), 8, B, f);
    fwrite(en.data(), 4, en.size(), f);
    fclose(f);
    return 0;
}
