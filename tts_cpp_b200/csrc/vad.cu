// vad.cu -- trailing-silence trim on the device (the step right after the forward / codec decode).
//
// Replaces apply_energy_voice_inactivity_detection + energy (reference examples/cli/vad.cpp:3-68; `tts-cli --vad`, "particularly useful for Parler TTS") for a BATCH
// of utterances whose PCM is resident in HBM, so that only the kept samples have to cross PCIe: per utterance one integer comes back (the new n_outputs).
//
//   vad_energy_kernel  frame energies e[i] = sum of squares of frame i's samples_per_frame samples, IN THE REFERENCE'S ORDER (one running fp32 sum per frame,
//                      vad.cpp:5-7) -- a frame is one thread's job; the samples reach it through a shared-memory tile that a whole block fills with coalesced
//                      512-byte row segments (32 frames x 128 samples per tile, pitch 129: the 32 summing lanes hit 32 different banks).  HBM-bound: every
//                      sample is read exactly once, 4 B per sample.
//   vad_decide_kernel  the two scans of vad.cpp:31-66 (running min / max, the early cut-off on `early_cutoff_frames` consecutive frames at or below the absolute
//                      threshold, the trailing run of frames below the min-max-normalised threshold), one thread per utterance: O(n_frames) integer / compare work.
//
// Bit-exactness: the result is an integer; it is the reference's for every input whose threshold comparisons are not exact ties.  The energies themselves are
// reproduced bit for bit against the reference BUILD the goldens come from (oracle/_ref, gcc -O2 -march=x86-64-v3): gcc keeps the running sum in order, computes the
// squares of the first (count & ~3) samples with a vector multiply followed by scalar adds (two roundings) and contracts the last (count & 3) into fused
// multiply-adds (one rounding) -- mirrored below, like the window-square-sum of the Kokoro path.
#include "kernels.cuh"

namespace b2 {

#define VAD_TF 32    // frames per block
#define VAD_TS 128   // samples per tile row

__global__ void __launch_bounds__(128) vad_energy_kernel(const float * pcm, const long long * off, const long long * eoff, int spf, float * energies) {
    __shared__ float tile[VAD_TF][VAD_TS + 1];
    const int b = blockIdx.y;
    const long long base = off[b], n = off[b + 1] - off[b];
    const int n_frames = (int) (n / spf);
    const int f0 = blockIdx.x * VAD_TF;
    if (f0 >= n_frames) return;                                   // uniform per block
    const int nf = n_frames - f0 < VAD_TF ? n_frames - f0 : VAD_TF;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = spf & ~3;                                   // see the header: [0, split) multiply then add, [split, spf) fused
    float en = 0.0f;
    for (int s0 = 0; s0 < spf; s0 += VAD_TS) {
        const int ns = spf - s0 < VAD_TS ? spf - s0 : VAD_TS;
        for (int f = warp; f < nf; f += 4) {                      // a warp fills one frame's row segment: 128 consecutive floats
            const float * src = pcm + base + (long long) (f0 + f) * spf + s0;
            for (int s = lane; s < ns; s += 32) tile[f][s] = src[s];
        }
        __syncthreads();
        if (threadIdx.x < nf) {
            const float * row = tile[threadIdx.x];
            for (int s = 0; s < ns; s++) {
                const float x = row[s];
                en = (s0 + s < split) ? __fadd_rn(en, __fmul_rn(x, x)) : fmaf(x, x, en);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < nf) energies[eoff[b] + f0 + threadIdx.x] = en;
}

__global__ void vad_decide_kernel(const long long * off, const long long * eoff, const float * energies, int B, int spf, int frame_threshold, float norm_threshold,
                                  int trailing, int early_frames, float early_threshold, long long * n_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long n = off[b + 1] - off[b];
    const int n_frames = (int) (n / spf);
    const float * e = energies + eoff[b];
    float mx = 0.0f, mn = 0.0f;
    int silent = 0;
    for (int i = 0; i < n_frames; i++) {                          // vad.cpp:31-51
        const float v = e[i];
        if (i == 0) { mx = v; mn = v; }
        else if (v > mx) mx = v;
        else if (v < mn) mn = v;
        silent = (v <= early_threshold) ? silent + 1 : 0;
        if (silent >= early_frames) { n_out[b] = (long long) ((i + trailing - silent) * spf); return; }      // int arithmetic, then widened: as the reference
    }
    int run = 0;
    for (int i = n_frames; i > 0; i--) {                          // vad.cpp:55-62; max == min gives NaN or inf, neither is < threshold: the loop ends like the reference's
        const float fe = __fdiv_rn(__fsub_rn(e[i - 1], mn), __fsub_rn(mx, mn));
        if (fe < norm_threshold) run++; else break;
    }
    long long out = n;
    if (run >= frame_threshold) out -= (long long) ((run - trailing) * spf);    // a negative product widens to a huge size_t upstream: the same bits in two's complement
    n_out[b] = out;
}

// d_pcm: the utterances back to back; d_off[B + 1]: their sample offsets; d_eoff[B + 1]: offsets into d_energies (n_b / spf frames each); all device pointers
int vad_trim_rows(Ctx * ctx, const float * d_pcm, const long long * d_off, const long long * d_eoff, int B, int max_frames, int spf, int frame_threshold,
                  float norm_threshold, int trailing, int early_frames, float early_threshold, float * d_energies, long long * d_n_out) {
    if (B <= 0) return 0;
    if (spf <= 0) { set_error("vad: ms_per_frame * sample_rate / 1000 < 1 (the reference divides by zero here)"); return 1; }
    if (max_frames > 0) {
        vad_energy_kernel<<<dim3((unsigned) cdiv(max_frames, VAD_TF), (unsigned) B), 128, 0, ctx->stream>>>(d_pcm, d_off, d_eoff, spf, d_energies);
        B2_LAUNCH_CHECK(ctx);
    }
    vad_decide_kernel<<<cdiv(B, 64), 64, 0, ctx->stream>>>(d_off, d_eoff, d_energies, B, spf, frame_threshold, norm_threshold, trailing, early_frames, early_threshold, d_n_out);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b2
