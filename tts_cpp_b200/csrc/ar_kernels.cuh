// ar_kernels.cuh -- plain CUDA-core kernels shared by the autoregressive decode paths (orpheus.cu, parler.cu, dia.cu).
//
// First correct versions: fp32 weights and activations as the reference computes them (ggml_mul_mat on F32 operands), one launch per op.
// Each kernel states the ggml op it restates.  Their LOGIC is checked under the CPU emulation in tests/emu (tests/test_emu_cpu.py);
// the -m gpu tests are the parity gate.  Included into each model's translation unit (internal linkage).
#pragma once
#include "common.cuh"

namespace b2 {
namespace {

// rows of a step: (sequence, position, token).  Decode steps have one row per sequence, built on the device from the last argmax.
__global__ void decode_rows_kernel(const int * __restrict__ n_prompt, const int * __restrict__ cur_tok, int B, int step, int Tmax, int * row_seq, int * row_pos, int * row_tok,
                                   int * row_base, int * row_len) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int pos = n_prompt[b] + step - 1;
    row_seq[b] = b; row_pos[b] = pos; row_tok[b] = cur_tok[b];
    row_base[b] = b * Tmax; row_len[b] = pos + 1;      // causal: the new row sees its sequence's cache up to and including itself
}

__global__ void embed_kernel(const int * __restrict__ row_tok, const float * __restrict__ embed, int H, float * __restrict__ x) {   // ggml_get_rows
    const int r = blockIdx.x;
    const float * src = embed + (size_t) row_tok[r] * H;
    for (int c = threadIdx.x; c < H; c += blockDim.x) x[(size_t) r * H + c] = src[c];
}

// ggml_rms_norm (float squares accumulated in a double, scale = 1/sqrtf(mean + eps)) followed by the weight multiply (model.cpp:122-125)
__global__ void rmsnorm_kernel(const float * __restrict__ x, const float * __restrict__ w, int H, int R, float * __restrict__ y) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float * row = x + (size_t) r * H;
    double s = 0.0;
    for (int c = lane; c < H; c += 32) s += (double) (row[c] * row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = (float) (s / (double) H);
    const float scale = 1.0f / sqrtf(mean + 1e-5f);
    for (int c = lane; c < H; c += 32) y[(size_t) r * H + c] = (row[c] * scale) * w[c];
}

// Y[r][n] = sum_k X[r][k] * W[n][k] (+ res[r][n]; res may alias Y: each element is read and written by the same thread): one warp per output n,
// the weight row is read once per chunk of 8 rows
// (ggml_mul_mat with F32 weights and activations; K % 4 == 0)
constexpr int GR = 8;
__global__ void __launch_bounds__(256) gemv_rows_kernel(const float * __restrict__ X, int ldx, const float * __restrict__ W, int K, int N, int R,
                                                        const float * res, float * Y, int ldy) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const float * wrow = W + (size_t) n * K;
    for (int r0 = 0; r0 < R; r0 += GR) {
        float acc[GR];
#pragma unroll
        for (int j = 0; j < GR; j++) acc[j] = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
            const float4 w4 = *reinterpret_cast<const float4 *>(wrow + k);
#pragma unroll
            for (int j = 0; j < GR; j++) {
                if (r0 + j < R) {
                    const float4 x4 = *reinterpret_cast<const float4 *>(X + (size_t) (r0 + j) * ldx + k);
                    acc[j] = fmaf(x4.w, w4.w, fmaf(x4.z, w4.z, fmaf(x4.y, w4.y, fmaf(x4.x, w4.x, acc[j]))));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < GR; j++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
            if (lane == 0 && r0 + j < R) Y[(size_t) (r0 + j) * ldy + n] = res ? acc[j] + res[(size_t) (r0 + j) * ldy + n] : acc[j];
        }
    }
}

// the same product for an F16 weight matrix: ggml_mul_mat converts the activation rows to fp16 first (the vec_dot_type of F16 is F16, ggml-cpu.c
// mul_mat from_float) and accumulates the exact fp16 x fp16 products in fp32 -- also what halves the bytes streamed per step
__global__ void __launch_bounds__(256) gemv_rows_h_kernel(const float * __restrict__ X, int ldx, const __half * __restrict__ W, int K, int N, int R,
                                                          const float * res, float * Y, int ldy) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const __half * wrow = W + (size_t) n * K;
    for (int r0 = 0; r0 < R; r0 += GR) {
        float acc[GR];
#pragma unroll
        for (int j = 0; j < GR; j++) acc[j] = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
            const __half2 * wp = reinterpret_cast<const __half2 *>(wrow + k);
            const float2 w01 = __half22float2(wp[0]), w23 = __half22float2(wp[1]);
#pragma unroll
            for (int j = 0; j < GR; j++) {
                if (r0 + j < R) {
                    const float4 x4 = *reinterpret_cast<const float4 *>(X + (size_t) (r0 + j) * ldx + k);
                    const float x0 = __half2float(__float2half_rn(x4.x)), x1 = __half2float(__float2half_rn(x4.y));
                    const float x2 = __half2float(__float2half_rn(x4.z)), x3 = __half2float(__float2half_rn(x4.w));
                    acc[j] = fmaf(x3, w23.y, fmaf(x2, w23.x, fmaf(x1, w01.y, fmaf(x0, w01.x, acc[j]))));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < GR; j++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
            if (lane == 0 && r0 + j < R) Y[(size_t) (r0 + j) * ldy + n] = res ? acc[j] + res[(size_t) (r0 + j) * ldy + n] : acc[j];
        }
    }
}

// NeoX RoPE over the whole head with per-pair frequency factors (ggml_rope_ext mode 2, theta base 5e5, ggml-cpu rope cache: theta starts at
// the position and is multiplied by theta_scale pair after pair), applied to q in place and to k on its way into the cache; v is copied
// (orpheus_build_kv_store, model.cpp:196-228 -- here the cache is compact: the 3x head expansion is done by indexing in the attention)
__global__ void rope_append_kernel(float * q, const float * __restrict__ k, const float * __restrict__ v, const float * __restrict__ ff, const int * __restrict__ row_seq,
                                   const int * __restrict__ row_pos, int heads, int kv_heads, int hd, float theta_scale, float * Kc, float * Vc, int Tmax) {
    const int r = blockIdx.x, h = blockIdx.y;                  // h < heads: a query head; h >= heads: kv head h - heads
    const int b = row_seq[r], pos = row_pos[r], half = hd >> 1;
    const int KV = kv_heads * hd, H = heads * hd;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float theta = (float) pos;
        for (int j = 0; j < i; j++) theta *= theta_scale;
        const float th = theta / ff[i];
        const float c = cosf(th), s = sinf(th);
        if (h < heads) {
            float * p = q + (size_t) r * H + (size_t) h * hd;
            const float x0 = p[i], x1 = p[i + half];
            p[i] = x0 * c - x1 * s; p[i + half] = x0 * s + x1 * c;
        } else {
            const int kh = h - heads;
            const float * p = k + (size_t) r * KV + (size_t) kh * hd;
            float * d = Kc + ((size_t) b * Tmax + pos) * KV + (size_t) kh * hd;
            const float x0 = p[i], x1 = p[i + half];
            d[i] = x0 * c - x1 * s; d[i + half] = x0 * s + x1 * c;
            const float * pv = v + (size_t) r * KV + (size_t) kh * hd;
            float * dv = Vc + ((size_t) b * Tmax + pos) * KV + (size_t) kh * hd;
            dv[i] = pv[i]; dv[i + half] = pv[i + half];
        }
    }
}

// attention of one query row over cache positions [row_base[r], row_base[r] + row_len[r]) of Kc / Vc (rows of kv_heads * hd floats): softmax(q.K^T * scale) V
// with ggml_soft_max's double-accumulated sum (ggml-cpu.c soft_max: max, expf, ggml_float sum, scale by (float)(1/sum)).  A causal or block mask of
// -inf entries is the same as restricting the range: exp(-inf) adds an exact zero.  Query head h reads kv head h / (heads / kv_heads) -- the
// reference's repeat-interleaved GQA cache (orpheus model.cpp:196-228, dia model.cpp:426-437) without the copies.  Tcap >= max row_len sizes the scores.
__global__ void __launch_bounds__(128) attention_kernel(const float * __restrict__ q, const float * __restrict__ Kc, const float * __restrict__ Vc,
                                                        const int * __restrict__ row_base, const int * __restrict__ row_len, int heads, int kv_heads, int hd,
                                                        int Tcap, float scale, float * __restrict__ out) {
    extern __shared__ float sc[];          // [T] scores, then 128 floats + 128 doubles of reduction scratch behind them
    const int r = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    const size_t base = (size_t) row_base[r];
    const int T = row_len[r];
    const int KV = kv_heads * hd, H = heads * hd, kh = h / (heads / kv_heads);
    float * redf = sc + ((Tcap + 1) & ~1);      // keeps the double scratch behind it 8-byte aligned
    double * redd = reinterpret_cast<double *>(redf + 128);
    const float * qv = q + (size_t) r * H + (size_t) h * hd;
    float mx = -INFINITY;
    for (int t = tid; t < T; t += 128) {
        const float * kr = Kc + (base + t) * KV + (size_t) kh * hd;
        float a = 0.f;
        for (int d = 0; d < hd; d++) a = fmaf(qv[d], kr[d], a);
        a *= scale;
        sc[t] = a;
        mx = fmaxf(mx, a);
    }
    redf[tid] = mx;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (tid < o) redf[tid] = fmaxf(redf[tid], redf[tid + o]); __syncthreads(); }
    mx = redf[0];
    double sum = 0.0;
    for (int t = tid; t < T; t += 128) { const float e = expf(sc[t] - mx); sc[t] = e; sum += (double) e; }
    redd[tid] = sum;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (tid < o) redd[tid] += redd[tid + o]; __syncthreads(); }
    const float inv = (float) (1.0 / redd[0]);
    for (int d = tid; d < hd; d += 128) {
        float a = 0.f;
        for (int t = 0; t < T; t++) a = fmaf(sc[t] * inv, Vc[(base + t) * KV + (size_t) kh * hd + d], a);
        out[(size_t) r * H + (size_t) h * hd + d] = a;
    }
}
static inline size_t attention_smem_bytes(int Tcap) { return (size_t) ((Tcap + 1) & ~1) * 4 + 128 * 4 + 128 * 8; }

__global__ void silu_mul_kernel(float * g, const float * __restrict__ u, size_t n) {      // ggml_silu (x / (1 + expf(-x))) * up
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float x = g[i]; g[i] = (x / (1.0f + expf(-x))) * u[i]; }
}

__global__ void gather_rows_f32_kernel(const float * __restrict__ x, const int * __restrict__ idx, int H, float * __restrict__ y) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < H; c += blockDim.x) y[(size_t) b * H + c] = x[(size_t) idx[b] * H + c];
}

// sampler::max: the first maximum wins
__global__ void __launch_bounds__(256) argmax_kernel(const float * __restrict__ logits, int V, int * cur_tok, int * out_tokens, int n_steps, int step) {
    __shared__ float sv[256]; __shared__ int si[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float * lg = logits + (size_t) b * V;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 256) { const float v = lg[i]; if (v > best) { best = v; bi = i; } }
    sv[tid] = best; si[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { if (sv[tid + o] > sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; } }
        __syncthreads();
    }
    if (tid == 0) { cur_tok[b] = si[0]; out_tokens[(size_t) b * n_steps + step] = si[0]; }
}


// ggml_norm (ggml-cpu.c:7114-7163: mean and variance of the centred row accumulated in double, scale = 1/sqrtf(var + eps)) followed by the
// weight multiply and bias add (parler model.cpp build_norm)
__global__ void layernorm_kernel(const float * __restrict__ x, const float * __restrict__ w, const float * __restrict__ bias, int H, int R, float eps, float * __restrict__ y) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float * row = x + (size_t) r * H;
    double s = 0.0;
    for (int c = lane; c < H; c += 32) s += (double) row[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = (float) (s / (double) H);
    double s2 = 0.0;
    for (int c = lane; c < H; c += 32) { const float v = row[c] - mean; s2 += (double) (v * v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    const float var = (float) (s2 / (double) H);
    const float scale = 1.0f / sqrtf(var + eps);
    for (int c = lane; c < H; c += 32) y[(size_t) r * H + c] = ((row[c] - mean) * scale) * w[c] + bias[c];
}

// x[r] = table[tok[r]] + pos_table[pos[r]] (prompt rows: ggml_get_rows + ggml_add of the positional rows)
__global__ void embed_pos_kernel(const int * __restrict__ row_tok, const int * __restrict__ row_pos, const float * __restrict__ table, const float * __restrict__ pos_table, int H,
                                 float * __restrict__ x) {
    const int r = blockIdx.x;
    const float * src = table + (size_t) row_tok[r] * H;
    const float * ps = pos_table + (size_t) row_pos[r] * H;
    for (int c = threadIdx.x; c < H; c += blockDim.x) x[(size_t) r * H + c] = src[c] + ps[c];
}

// the audio-token input of a decode step: the rows of the n_out codebook tables summed in head order ((e0 + e1) + e2 ...: parler_build_inp_embd,
// model.cpp:320-340; dia model.cpp build_dia_decoder_inp_embd), plus the positional row when there is one.  tables[i] rows are tab_rows apart.
__global__ void codebook_embed_kernel(const int * __restrict__ ids, int n_out, const float * __restrict__ tables, size_t tab_stride, const float * __restrict__ pos_table,
                                      const int * __restrict__ row_pos, int H, float * __restrict__ x) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = tables[(size_t) ids[(size_t) r * n_out] * H + c];
        for (int i = 1; i < n_out; i++) a = tables[(size_t) i * tab_stride + (size_t) ids[(size_t) r * n_out + i] * H + c] + a;
        if (pos_table) a = a + pos_table[(size_t) row_pos[r] * H + c];
        x[(size_t) r * H + c] = a;
    }
}

// NeoX RoPE in place on rows of nh heads (ggml_rope_ext mode 2 without frequency factors; theta advanced by repeated multiplication like the ggml-cpu cache)
__global__ void rope_rows_kernel(float * x, const int * __restrict__ row_pos, int nh, int hd, float theta_scale) {
    const int r = blockIdx.x, h = blockIdx.y, half = hd >> 1;
    const int pos = row_pos[r];
    float * p = x + (size_t) r * nh * hd + (size_t) h * hd;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float theta = (float) pos;
        for (int j = 0; j < i; j++) theta *= theta_scale;
        const float c = cosf(theta), s = sinf(theta);
        const float x0 = p[i], x1 = p[i + half];
        p[i] = x0 * c - x1 * s; p[i + half] = x0 * s + x1 * c;
    }
}

// copy the new k / v rows to their cache slots (row_dst[r] in rows of KV floats): ggml_cpy into the cache views
__global__ void store_kv_kernel(const float * __restrict__ k, const float * __restrict__ v, const int * __restrict__ row_dst, int KV, float * Kc, float * Vc) {
    const int r = blockIdx.x;
    const size_t d = (size_t) row_dst[r] * KV;
    for (int c = threadIdx.x; c < KV; c += blockDim.x) { Kc[d + c] = k[(size_t) r * KV + c]; if (v) Vc[d + c] = v[(size_t) r * KV + c]; }
}

// ggml's GELU for F32 tensors: an fp16 lookup table of the tanh approximation (ggml-cpu.c:1816-1830), in place
__global__ void gelu_f16lut_kernel(float * g, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = g[i];
    float y;
    if (x <= -10.0f) y = 0.0f;
    else if (x >= 10.0f) y = x;
    else {
        const float xh = __half2float(__float2half_rn(x));
        y = __half2float(__float2half_rn(0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh)))));
    }
    g[i] = y;
}

// sampler::max over rows of V logits: the first maximum wins.  row -> out[step * gridDim.x + row], step = *d_step (0 when d_step is null)
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float * __restrict__ logits, int V, int * __restrict__ out, const int * __restrict__ d_step) {
    __shared__ float sv[256]; __shared__ int si[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float * lg = logits + (size_t) b * V;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 256) { const float v = lg[i]; if (v > best) { best = v; bi = i; } }
    sv[tid] = best; si[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { if (sv[tid + o] > sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; } }
        __syncthreads();
    }
    if (tid == 0) out[(size_t) (d_step ? *d_step : 0) * gridDim.x + b] = si[0];
}

// rows of an audio decode step under the delay pattern (parler generate_from_batch, model.cpp:762-786; dia model.cpp:843-858): output head i is fed BOS
// until step i + 1, then the token it produced in the previous step (d_out [steps][B][n_out]).  One row per sequence at position first_pos[b] + step.
// The step number lives in device memory (d_step, advanced by step_advance_kernel) so that one captured CUDA graph of a step can be replayed for every step.
__global__ void delay_rows_kernel(const int * __restrict__ d_out, const int * __restrict__ first_pos, int B, int n_out, const int * __restrict__ d_step, int bos, int Tmax,
                                  int * ids, int * row_pos, int * row_base, int * row_len, int * row_dst) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int step = *d_step;
    const int pos = (first_pos ? first_pos[b] : 0) + step;
    for (int i = 0; i < n_out; i++) ids[b * n_out + i] = step > i ? d_out[((size_t) (step - 1) * B + b) * n_out + i] : bos;
    row_pos[b] = pos; row_base[b] = b * Tmax; row_len[b] = pos + 1; row_dst[b] = b * Tmax + pos;
}

__global__ void step_advance_kernel(int * d_step) { if (threadIdx.x == 0 && blockIdx.x == 0) *d_step += 1; }

}  // namespace
}  // namespace b2
