// sampler.cu -- the reference sampler on the device ("sampler.cpp hot loop").
//
// Restates sampler::sample (reference src/sampler.cpp:3-70) for many heads at once, one block per head: sampler::max (first maximum; also the stabiliser of
// the softmax), the repetition penalty (the last sampled token's logit divided by penalty^count, in double like std::pow), temperature, top-k (sorted by
// value, ties to the lower id), the softmax with its sequential fp32 denominator, top-p trimming WITHOUT renormalisation (the uniform is scaled by
// min(prob_sum, top_p) instead) and the cumulative draw.  The reference draws its uniform from a std::minstd_rand freshly seeded from std::random_device on
// every call, so nothing about its random stream can be reproduced: here the uniform is a counter-based hash of (seed, row, step), which makes runs
// repeatable and lets a captured CUDA graph of a decode step be replayed.  oracle/sampler_port.py is the checker (pinned to the reference stage by stage
// and, for the draw rule, by a histogram of the reference's own draws).
//
// The nucleus (the k best entries in (value descending, id ascending) order -- the reference's std::sort with a stable tie rule) comes from a radix select:
// four 8-bit histogram passes over an order-preserving integer key find the k-th largest key and how many entries equal to it belong to the nucleus, one more
// pass collects them (ties at the threshold: lowest ids first), and the <= 1 024 picks are put in order by rank counting in shared memory -- 5 reads of the row
// instead of k (Orpheus: V = 156 940, k = 50).  The sequential fp32 sums that fix the rounding of the reference are done by one thread.
#include "kernels.cuh"

#include <cmath>

namespace b2 {
namespace {

__device__ __host__ inline float uniform_from_counter(unsigned long long seed, unsigned long long row, unsigned long long step) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (row + 1) + 0xD1B54A32D192ED03ull * (step + 1);      // splitmix64 finaliser over a mixed counter
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (float) (x >> 40) * (1.0f / 16777216.0f);                                                           // 24 bits -> [0, 1)
}

struct ArgMax { float v; int i; };
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// block-wide (value, index) maximum, ties to the lower index; every thread gets the result
__device__ ArgMax block_argmax(float v, int i, float * sv, int * si) {
    const int tid = threadIdx.x;
    sv[tid] = v; si[tid] = i;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o && better(sv[tid + o], si[tid + o], sv[tid], si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; }
        __syncthreads();
    }
    ArgMax r{sv[0], si[0]};
    __syncthreads();
    return r;
}

// order-preserving key of a float (larger value <-> larger key; -0 and +0 share a key like `==` says; NaN, which `better` never picks, sorts below -inf)
__device__ __forceinline__ unsigned order_key(float v) {
    if (v != v) return 0u;
    const unsigned b = __float_as_uint(v + 0.0f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// The k best of val(0 .. V-1) in (value descending, id ascending) order -> pick_idx / pick_val[0 .. k), 1 <= k <= min(V, SAMPLE_MAX_TOP_K); 256 threads.
// hist: [8 warps][256] counters, tmp_*: [SAMPLE_MAX_TOP_K] unordered picks, misc: [4] ints -- all shared memory.  Every thread must call it (barriers inside).
// have_prev: only the entries strictly AFTER (prev_key, prev_id) in that order take part -- the continuation of an earlier call whose last pick that was, so that a
// nucleus of more than SAMPLE_MAX_TOP_K entries is produced chunk by chunk (k <= the number of such entries).
template <class F>
__device__ void block_topk(F val, int V, int k, int * pick_idx, float * pick_val, int * tmp_idx, float * tmp_val, unsigned * hist, int * misc,
                           bool have_prev = false, unsigned prev_key = 0u, int prev_id = -1) {
    const int tid = threadIdx.x, warp = tid >> 5;
    auto live = [&](unsigned key, int id) -> bool { return !have_prev || key < prev_key || (key == prev_key && id > prev_id); };
    unsigned prefix = 0u, mask = 0u;
    int need = k;                                              // entries still to be found among those whose key matches `prefix` under `mask`
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 8 * 256; i += 256) hist[i] = 0u;
        __syncthreads();
        for (int ii = tid; ii < V; ii += 256) {
            const unsigned key = order_key(val(ii));
            if ((key & mask) == prefix && live(key, ii)) atomicAdd(&hist[warp * 256 + ((key >> shift) & 255u)], 1u);      // a histogram per warp: 8x less contention
        }
        __syncthreads();
        unsigned c = 0u;
#pragma unroll
        for (int w = 0; w < 8; w++) c += hist[w * 256 + tid];
        __syncthreads();
        hist[tid] = c;                                         // row 0 becomes the digit histogram, then its suffix sums S[d] = #entries with digit >= d
        hist[256 + tid] = c;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const unsigned v = tid + o < 256 ? hist[tid + o] : 0u;
            __syncthreads();
            hist[tid] += v;
            __syncthreads();
        }
        {   // the digit d with S[d] >= need > S[d+1]: exactly one thread sees it (S[0] >= need by construction)
            const unsigned above = tid < 255 ? hist[tid + 1] : 0u;
            if (hist[tid] >= (unsigned) need && above < (unsigned) need) { misc[0] = tid; misc[1] = need - (int) above; misc[2] = (int) hist[256 + tid]; }
        }
        __syncthreads();
        prefix |= (unsigned) misc[0] << shift; mask |= 255u << shift; need = misc[1];
        __syncthreads();
    }
    // prefix = the k-th largest key T; `need` of the misc[2] entries equal to T belong to the nucleus, together with the k - need entries above T
    const unsigned T = prefix;
    const int n_eq = misc[2], n_gt = k - need;
    if (tid == 0) misc[3] = 0;
    __syncthreads();
    if (n_eq == need) {                                        // no tie across the boundary (the usual case): everything >= T, in any order
        for (int ii = tid; ii < V; ii += 256) {
            const float v = val(ii);
            const unsigned key = order_key(v);
            if (key >= T && live(key, ii)) { const int slot = atomicAdd(&misc[3], 1); tmp_idx[slot] = ii; tmp_val[slot] = v; }
        }
    } else {                                                   // ties at the threshold: the `need` lowest ids among them
        for (int ii = tid; ii < V; ii += 256) {
            const float v = val(ii);
            const unsigned key = order_key(v);
            if (key > T && live(key, ii)) { const int slot = atomicAdd(&misc[3], 1); tmp_idx[slot] = ii; tmp_val[slot] = v; }
        }
        const int chunk = (V + 255) / 256, lo = tid * chunk, hi = lo + chunk < V ? lo + chunk : V;      // thread t owns ids [t * chunk, (t + 1) * chunk)
        unsigned mine = 0u;
        for (int ii = lo; ii < hi; ii++) mine += (order_key(val(ii)) == T && live(T, ii)) ? 1u : 0u;
        __syncthreads();                                       // (hist is free again: everyone left the select loop)
        hist[tid] = mine;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {                    // inclusive prefix sums over the threads
            const unsigned v = tid >= o ? hist[tid - o] : 0u;
            __syncthreads();
            hist[tid] += v;
            __syncthreads();
        }
        int rank = (int) (hist[tid] - mine);                   // equal entries with lower ids
        for (int ii = lo; ii < hi && rank < need; ii++) {
            const float v = val(ii);
            if (order_key(v) == T && live(T, ii)) { tmp_idx[n_gt + rank] = ii; tmp_val[n_gt + rank] = v; rank++; }
        }
    }
    __syncthreads();
    for (int i = tid; i < k; i += 256) {                        // rank counting: (value, id) pairs are distinct, so the ranks are a permutation of 0 .. k-1
        const float v = tmp_val[i]; const int id = tmp_idx[i];
        int rank = 0;
        for (int j = 0; j < k; j++) rank += better(tmp_val[j], tmp_idx[j], v, id) ? 1 : 0;
        pick_idx[rank] = id; pick_val[rank] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) sample_rows_kernel(const SampleParams p) {
    __shared__ float sv[256]; __shared__ int si[256];
    __shared__ int pick_idx[SAMPLE_MAX_TOP_K]; __shared__ float pick_val[SAMPLE_MAX_TOP_K];
    __shared__ int tmp_idx[SAMPLE_MAX_TOP_K]; __shared__ float tmp_val[SAMPLE_MAX_TOP_K];
    __shared__ unsigned hist[8 * 256]; __shared__ int misc[4];
    __shared__ int s_n, s_tok; __shared__ float s_mh;
    const int row = blockIdx.x, tid = threadIdx.x, V = p.V;
    const float * lg = p.logits + (size_t) row * V;
    const int step = p.d_step ? *p.d_step : 0;
    const bool has_rp = p.repetition_penalty != 1.0f;
    const int last = has_rp ? p.last_ids[row] : -1;
    double pen_d = 1.0;                                   // applied as float(v / pow(penalty, count)): the double division of the reference
    if (has_rp && last >= 0) pen_d = pow((double) p.repetition_penalty, (double) p.rep_counts[row]);
    auto eff = [&](int ii) -> float { const float v = lg[ii]; return (has_rp && ii == last) ? (float) ((double) v / pen_d) : v; };

    // sampler::max
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int ii = tid; ii < V; ii += 256) { const float v = eff(ii); if (better(v, ii, bv, bi)) { bv = v; bi = ii; } }
    const ArgMax mx = block_argmax(bv, bi, sv, si);
    int tok = mx.i == 0x7fffffff ? 0 : mx.i;               // (all-NaN logits: token 0, never an out-of-range id)
    if (p.do_sample) {
        const bool has_t = p.temperature != 1.0f;
        const float max_val = has_t ? mx.v / p.temperature : mx.v;
        const bool nucleus_k = p.top_k > 0 && p.top_k < V;
        const float u = uniform_from_counter(p.seed, (unsigned long long) row, (unsigned long long) step);
        float * probs_all = p.scratch ? p.scratch + (size_t) row * V : nullptr;
        if (p.top_p < 1.0f) {
            // softmax over the whole vocabulary first, then the picks in order of probability until top_k or top_p is reached
            for (int ii = tid; ii < V; ii += 256) { float v = eff(ii); if (has_t) v /= p.temperature; probs_all[ii] = expf(v - max_val); }
            __syncthreads();
            if (tid == 0) { float s = 0.f; for (int ii = 0; ii < V; ii++) s += probs_all[ii]; s_mh = s; }
            __syncthreads();
            const float denom = s_mh;
            __syncthreads();
            for (int ii = tid; ii < V; ii += 256) probs_all[ii] = probs_all[ii] / denom;
            __syncthreads();
            // the picks come in chunks of up to SAMPLE_MAX_TOP_K, each the continuation of the one before (one chunk whenever top_k is set, and for any peaked distribution)
            const int ktotal = nucleus_k ? p.top_k : V;
            auto prob = [&](int ii) { return probs_all[ii]; };
            float prob_sum = 0.f; int n = 0, chunks = 0; bool done = false;      // the same in every thread: all of them walk the sorted picks
            unsigned prev_key = 0u; int prev_id = -1;
            while (!done && n < ktotal) {
                const int k = ktotal - n < SAMPLE_MAX_TOP_K ? ktotal - n : SAMPLE_MAX_TOP_K;
                block_topk(prob, V, k, pick_idx, pick_val, tmp_idx, tmp_val, hist, misc, chunks > 0, prev_key, prev_id);
                for (int j = 0; j < k; j++) { prob_sum += pick_val[j]; n++; if (prob_sum >= p.top_p) { done = true; break; } }
                chunks++;
                if (!done && n < ktotal) { prev_key = order_key(pick_val[k - 1]); prev_id = pick_idx[k - 1]; __syncthreads(); }      // (the next call overwrites the picks)
            }
            const float mh = fminf(prob_sum, p.top_p);
            if (chunks == 1) {
                if (tid == 0) { s_n = n; s_mh = mh; }                              // the nucleus is in shared memory: the common draw below
            } else {
                // a nucleus of several chunks: produce them once more for the cumulative draw (same picks, same order)
                __syncthreads();
                const float a = u * mh;
                float c = 0.f; int t = -1, seen = 0, ch = 0;
                while (t < 0) {
                    const int k = n - seen < SAMPLE_MAX_TOP_K ? n - seen : SAMPLE_MAX_TOP_K;
                    block_topk(prob, V, k, pick_idx, pick_val, tmp_idx, tmp_val, hist, misc, ch > 0, prev_key, prev_id);
                    for (int j = 0; j < k; j++) { c += pick_val[j]; seen++; if (a <= c || seen >= n) { t = pick_idx[j]; break; } }
                    ch++;
                    if (t < 0) { prev_key = order_key(pick_val[k - 1]); prev_id = pick_idx[k - 1]; __syncthreads(); }
                }
                if (tid == 0) { s_tok = t; s_n = 0; }
            }
        } else if (nucleus_k) {
            // top-k by value, then the softmax over the picks
            block_topk(eff, V, p.top_k, pick_idx, pick_val, tmp_idx, tmp_val, hist, misc);
            if (tid == 0) {
                float s = 0.f;
                for (int n = 0; n < p.top_k; n++) { float v = pick_val[n]; if (has_t) v /= p.temperature; v = expf(v - max_val); pick_val[n] = v; s += v; }
                for (int n = 0; n < p.top_k; n++) pick_val[n] = pick_val[n] / s;
                s_n = p.top_k; s_mh = 1.0f;
            }
        } else {
            // no nucleus: the softmax over the whole vocabulary, drawn in index order
            for (int ii = tid; ii < V; ii += 256) { float v = eff(ii); if (has_t) v /= p.temperature; probs_all[ii] = expf(v - max_val); }
            __syncthreads();
            if (tid == 0) {
                float s = 0.f;
                for (int ii = 0; ii < V; ii++) s += probs_all[ii];
                float c = 0.f; int t = V - 1;
                for (int ii = 0; ii < V; ii++) { c += probs_all[ii] / s; if (u <= c) { t = ii; break; } }
                s_tok = t; s_n = 0;
            }
        }
        __syncthreads();
        if (tid == 0 && s_n > 0) {
            const float a = p.top_p < 1.0f ? u * s_mh : u;
            float c = 0.f; int t = pick_idx[s_n - 1];
            for (int n = 0; n < s_n; n++) { c += pick_val[n]; if (a <= c || n >= s_n - 1) { t = pick_idx[n]; break; } }
            s_tok = t;
        }
        __syncthreads();
        tok = s_tok;
    }
    if (tid == 0) {
        if (has_rp && p.do_sample) {                      // sampler::sample's repetition bookkeeping (sampler::max does none)
            int cnt = p.rep_counts[row];
            if (last != tok) cnt = 0;
            p.last_ids[row] = tok; p.rep_counts[row] = cnt + 1;
        }
        if (p.out_stride_steps) p.out[(size_t) row * p.out_stride_steps + step] = tok;
        else p.out[(size_t) step * p.rows + row] = tok;
        if (p.cur_tok) p.cur_tok[row] = tok;
    }
}

}  // namespace

float sample_uniform_host(unsigned long long seed, unsigned long long row, unsigned long long step) { return uniform_from_counter(seed, row, step); }

int sample_rows(Ctx * ctx, const SampleParams & p) {
    if (p.rows <= 0) return 0;
    if (p.do_sample) {
        if (p.top_k > SAMPLE_MAX_TOP_K) { set_error("sampler: top_k %d > %d", p.top_k, SAMPLE_MAX_TOP_K); return 1; }
        const bool nucleus_k = p.top_k > 0 && p.top_k < p.V;
        if ((p.top_p < 1.0f || !nucleus_k) && !p.scratch) { set_error("sampler: top_p < 1 or top_k == 0 needs the [rows][V] scratch buffer"); return 1; }
        if (p.repetition_penalty != 1.0f && (!p.last_ids || !p.rep_counts)) { set_error("sampler: repetition penalty needs the last_ids / rep_counts state"); return 1; }
    }
    sample_rows_kernel<<<p.rows, 256, 0, ctx->stream>>>(p);
    B2_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace b2
