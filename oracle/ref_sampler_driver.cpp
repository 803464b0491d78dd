// oracle/ref_sampler_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED reference sampler (reference src/sampler.cpp, src/sampler.h) on given logits:
//   * the deterministic stages of sampler::sample in the order it runs them (max -> [softmax] -> topk -> [softmax] -> topp, sampler.cpp:3-42) and dumps the
//     nucleus (picks), its probabilities and max_head_probs per head;
//   * sampler::sample itself n_draws times (its generator is seeded from std::random_device, so only the DISTRIBUTION of its draws can be pinned): the
//     histogram of sampled token ids per head.
// usage: sampler_ref <in.bin> <out.bin>
//   in : u32 heads, u32 vocab, f32 temperature, u32 top_k, f32 top_p, f32 repetition_penalty, i32 last_token_ids[heads], u32 repetition_counts[heads],
//        u32 n_draws, f32 logits[heads][vocab]
//   out: per head: u32 n_picks, u32 picks[n_picks], f32 probs[n_picks], f32 max_head_prob; then u32 hist[heads][vocab]
#include "sampler.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

int main(int argc, char ** argv) {
    if (argc < 3) return 2;
    FILE * f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint32_t H, V, top_k, n_draws; float temperature, top_p, rp;
    if (fread(&H, 4, 1, f) != 1 || fread(&V, 4, 1, f) != 1 || fread(&temperature, 4, 1, f) != 1 || fread(&top_k, 4, 1, f) != 1 || fread(&top_p, 4, 1, f) != 1 || fread(&rp, 4, 1, f) != 1) return 2;
    std::vector<int32_t> last(H); std::vector<uint32_t> counts(H);
    if (fread(last.data(), 4, H, f) != H || fread(counts.data(), 4, H, f) != H || fread(&n_draws, 4, 1, f) != 1) return 2;
    std::vector<float> logits((size_t) H * V);
    if (fread(logits.data(), 4, logits.size(), f) != logits.size()) return 2;
    fclose(f);
    auto make = [&]() { sampler s; s.n_output_heads = H; s.vocab_size = V; s.temperature = temperature; s.top_k = top_k; s.top_p = top_p; s.repetition_penalty = rp;
                        s.do_sample = true; s.last_token_ids = last; s.repetition_counts = counts; return s; };
    FILE * o = fopen(argv[2], "wb");
    {   // the deterministic stages, exactly as sampler::sample sequences them
        sampler s = make();
        std::vector<float> lg = logits;
        std::vector<uint32_t> max_vals; std::vector<float> max_head_probs; std::vector<std::vector<size_t>> picks;
        bool performed_softmax = false;
        s.max(lg.data(), max_vals);
        if (top_p < 1.0) { s.softmax(lg.data(), picks, max_vals); performed_softmax = true; }
        if (top_k > 0 && top_k < V) picks = s.topk(lg.data(), performed_softmax);
        if (top_p >= 1.0) s.softmax(lg.data(), picks, max_vals);
        if (top_p < 1.0) s.topp(lg.data(), picks, max_head_probs);
        for (uint32_t i = 0; i < H; i++) {
            std::vector<uint32_t> p;
            if (picks.empty()) { for (uint32_t j = 0; j < V; j++) p.push_back(j); } else { for (size_t j : picks[i]) p.push_back((uint32_t) j); }
            uint32_t n = (uint32_t) p.size();
            fwrite(&n, 4, 1, o); fwrite(p.data(), 4, n, o);
            for (uint32_t j : p) fwrite(&lg[(size_t) i * V + j], 4, 1, o);
            float m = top_p < 1.0 ? max_head_probs[i] : 1.0f;
            fwrite(&m, 4, 1, o);
        }
    }
    std::vector<uint32_t> hist((size_t) H * V, 0);
    for (uint32_t d = 0; d < n_draws; d++) {
        sampler s = make();
        std::vector<float> lg = logits;
        std::vector<uint32_t> out;
        s.sample(lg.data(), out);
        for (uint32_t i = 0; i < H && i < out.size(); i++) if (out[i] < V) hist[(size_t) i * V + out[i]]++;
    }
    fwrite(hist.data(), 4, hist.size(), o);
    fclose(o);
    return 0;
}
