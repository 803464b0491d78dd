// pdk.cuh -- the persistent decode kernel: a whole run of autoregressive decode steps in ONE cooperative launch (BASELINE north star:
// "persistent ... decode kernels fed by TMA from a paged HBM KV cache, LayerNorm / GELU folded into the epilogues").
//
// What it replaces: the ~11 micro-launches per layer of parler.cu's launch-per-op path (reference build_parler_graph, src/models/parler/model.cpp:520-614, one GGML
// graph per step).  Measured on a B200 (profiles/r2a_*): every one of those launches is a 10-20 us latency chain (one CTA per SM, a dozen dependent L2 round trips),
// 3.97 ms per Parler-Mini step at batch 16 under CUDA-graph replay -- 4 % of the weight-streaming roofline.  Here one grid of (SM count) CTAs walks a device-resident
// PROGRAM of the step -- rows/embedding, then per layer { [LN1] q|k|v GEMV -> self-attention -> o GEMV + residual -> [LN2] cross-q GEMV -> cross-attention -> o GEMV +
// residual -> [LN3] fc1 GEMV + GELU -> fc2 GEMV + residual }, final [LN] heads GEMV, argmax -- separated by grid-wide barriers in global memory, for up to 32 steps
// per launch:
//
//   * warp 8 of every CTA is a TMA PRODUCER: eight lanes stream the CTA's weight tiles (8 output rows x 1 024 k, fp16; `cp.async.bulk` global -> shared completing
//     on an mbarrier) through a ring of ~9 stages IN PROGRAM ORDER.  Weights do not depend on activations, so the producer runs ahead across phases, layers and
//     steps: while the consumers sit in a grid barrier or in an attention phase the ring fills with the next phase's tiles and HBM keeps streaming.
//   * warps 0-7 are CONSUMERS: per GEMV phase they stage the <= 16 activation rows ONCE per CTA as fp16 in shared memory -- with the LayerNorm of the reference
//     (ggml_norm: double-accumulated mean / variance) folded into that staging pass, so no norm kernel and no normalised tensor in HBM -- and run
//     mma.sync.m16n8k16 (batch = M, exact fp16 products, fp32 accumulation: ggml_mul_mat's numerics for F16 matrices) on their k-slice of every tile; the eight
//     partial 16 x 8 tiles are summed in a fixed order and the epilogue applies GELU (fp16 table, ggml-cpu.c:1816-1830) / the residual / the KV-cache append.
//     F32 matrices (the output heads) go through the same pipeline as fp16 (hi, 2^11-scaled lo) plane pairs with the fp32-faithful three-product rule of
//     ar_kernels.cuh (gemv_mma_body<SPLIT>).
//   * the self-attention KV cache is PAGED: pages of 32 positions x all heads, fp16 (or fp32: template parameter), a page table per sequence; the new k / v rows
//     are written by the q|k|v phase's epilogue straight into their page.  Attention runs one (row, head) item per half-CTA: 16-byte loads of K / V rows spread
//     over the threads so that dozens are in flight, scores in shared memory, the reference's softmax (max, expf, double-accumulated sum).
//
// tcgen05 is deliberately not used: M = batch <= 16 rows, the step is HBM- and latency-bound; what matters is bytes in flight and the dependency chain.
// Logic checked in the build container under tests/emu (cooperative launch = all blocks' threads alive at once as fibers; mbarriers, bulk copies and the grid
// barrier have functional models below).
#pragma once
#include "kokoro.h"   // CUDA types, HostTensor, ArW

#include <algorithm>
#include <cstdio>
#include <functional>
#include <vector>

// This header has two faces.  Included plainly (parler.cu, orpheus.cu, dia.cu) it declares the program types and the host entry points; pdk.cu defines
// B2_PDK_IMPLEMENTATION before including it and is the one translation unit that carries the kernel (six instantiations) -- compiled into every model file it was 3 x 7 MB
// of identical device code.
namespace b2 {


constexpr int PK_CONS = 256;                      // consumer threads (8 warps)
constexpr int PK_THREADS = 288;                   // + the producer warp
constexpr int PK_TK = 1024;                       // k extent of a weight tile (halves)
constexpr int PK_PAD = 32;                        // halves of padding per smem row: rows start 16 banks apart (conflict-free 128-bit fragment loads, see GM_PAD)
constexpr int PK_ROWB = (PK_TK + PK_PAD) * 2;     // bytes per tile row in shared memory
constexpr int PK_STAGE = 8 * PK_ROWB;             // bytes per ring stage (one tile: 8 output rows)
constexpr int PK_AK_MAX = 3072;                   // activations are staged in k chunks of P.ak columns (2 048, or 3 072 for models whose hidden size is 3 072: one chunk for their q|k|v, gate|up phases)
constexpr int PK_ROWQ = PK_TK + 32;               // bytes per row of a Q8_0 tile's values in shared memory (rows 8 banks apart: conflict-free 8-byte fragment loads)
constexpr int PK_QSC = 8 * PK_ROWQ;               // byte offset of a Q8_0 tile's block scales within its stage: [8 rows][32 blocks] fp16
constexpr int PK_PAGE = 32;                       // positions per KV page
constexpr int PK_MAXSTAGES = 12;
constexpr int PK_RED_FLOATS = 2 * 8 * 128;        // cross-warp reduction scratch of one unit: eight warps' partial 16 x 8 tiles, twice for a paired unit
constexpr int PK_RED_BYTES = 2 * PK_RED_FLOATS * 4;   // two such buffers used alternately (one block barrier per unit instead of two) / argmax scratch
constexpr int PK_REP = 8;                         // copies of every activation buffer that ALL CTAs read at the start of a phase.  Measured on a B200: 148 SMs asking L2 for the
                                                  // same 64 KB right after a grid barrier wait ~2 us (each line is served to 148 requesters one after the other); with 8 copies
                                                  // (8x the tiny epilogue stores) a line has 18-19 requesters

enum { PK_ROWS = 0, PK_GEMV = 1, PK_ATTN = 2, PK_ARGMAX = 3, PK_ATTNC = 4, PK_QUANT = 5 };      // PK_QUANT: Q8_0-quantise the activation rows once for the whole grid      // PK_ATTNC: combine the position chunks of a split attention phase
enum { PKN_NONE = 0, PKN_LAYER = 1, PKN_RMS = 2 };
enum { PKE_STORE = 0, PKE_RES = 1, PKE_GELU = 2, PKE_KV = 3, PKE_LOGITS = 4, PKE_ROPE_Q = 5, PKE_ROPE_K = 6, PKE_SWIGLU = 7 };
enum { PKP_NONE = 0, PKP_ROPE = 1, PKP_SWIGLU = 2 };      // paired units: two tiles per k-tile (the two NeoX halves of a head's rows; the gate and the up rows of the same columns), one joint epilogue
enum { PKM_PARLER = 0, PKM_ORPHEUS = 1, PKM_DIA = 2 };

struct PkSeg {                                    // one matrix of a GEMV phase; unit = 8 consecutive output rows
    const __half * W;                             // [N][K] fp16 (split: the high plane)
    const __half * Wl;                            // split: the low plane (scaled by 2^11), else null
    float * Y; const float * res;                 // STORE / RES / LOGITS: Y[r * ldy + n] (+ res[r * ldy + n])
    __half * Y16;                                 // GELU: the activated values as fp16 [r * ldy + n] (their only consumer, fc2, rounds its input rows to fp16 anyway)
    size_t yrep;                                  // != 0: Y / Y16 (and res) exist in PK_REP copies this many elements apart; the epilogue writes them all, a CTA reads copy blockIdx % PK_REP
    const __half * Ws; const __half * Wps;        // Q8_0 phases (op.q8): W / Wp point at int8 values [N][K], Ws / Wps at their fp16 block scales [N][K / 32]
    const __half * Wp;                            // paired units: the partner tile's matrix (PKP_ROPE: W itself, rows + hd / 2; PKP_SWIGLU: the up matrix, same rows)
    int N, unit0, epi, ldy, kv, pair, n_units;    // unit0: first unit of this segment within the phase; kv: 0 = K, 1 = V; pair: PKP_*; n_units of this segment
};
struct alignas(16) PkOp {
    int kind, layer;
    // PK_GEMV
    const float * X; const __half * X16;          // input rows: fp32 (residual stream, q) or, when X16 is set, fp16 written by the previous phase (attention output, GELU output)
    size_t xrep;                                  // != 0: X / X16 exist in PK_REP copies this many elements apart
    const float * nw; const float * nb; int ldx, K, norm; float eps; int nseg, n_units; int kv_prefetch; int q8;
    // Q8_0 phases: XQ / XD (GEMV: input rows already quantised by a PK_QUANT op: int8 [16][K] + block scales [16][K / 32], PK_REP copies qrep bytes / qdrep floats apart);
    // QY / QD (PK_QUANT: where to put them).  Quantising inside every CTA's staging pass cost 20-26 us per phase on a B200 -- 148 CTAs repeating the same 16 rows
    const unsigned char * XQ; const float * XD; unsigned char * QY; float * QD; size_t qrep, qdrep;
    PkSeg seg[3];      // q8: the matrices are Q8_0 blocks (ggml_vec_dot_q8_0_q8_0 arithmetic: activations quantised per 32-block, int8 MMA, fp32 scale products)      // kv_prefetch: L2-prefetch this layer's K / V rows first
    // PK_ATTN: cross != 0 -> every row attends to the flat fp32 store ck / cv [cross_len][H]; else to its sequence's pages, positions [0, row_pos[r]]
    const float * q; __half * out16; size_t orep; const float * ck; const float * cv; int cross, cross_len; float scale; size_t cross_row_stride; int tsplit;      // tsplit > 1: every (row, head) item is cut into tsplit position chunks (few items, long contexts), combined by a PK_ATTNC op      // cross_row_stride: elements between the stores of consecutive rows (Dia: one encoding per sequence; 0: all rows share one)      // out16 [R][H] fp16: consumed only by the o-projection, which rounds to fp16
};
struct PkParams {
    const PkOp * ops; int n_ops;
    int R, H, heads, kv_heads, hd, n_out, vocab;
    int model, ak, pos_off;                       // PKM_*; activation chunk columns; position of a row = first_pos + step - pos_off
    // PKM_ORPHEUS rows: x0 = embed[token produced one step earlier]; NeoX RoPE table of the step rope_cs [R][hd / 2] (cos, sin) with ggml's iterated theta and the
    // llama-3 frequency factors; d_out is [R][n_steps_total]; a sequence stops at stop_token
    const float * embed; const float * rope_ff; float2 * rope_cs; float theta_scale; int n_steps_total, stop_token;
    float * amax_v; int * amax_i; int amax_ch;
    // PKM_DIA (reference src/models/dia/model.cpp:806-858): rows 2u (conditional) and 2u + 1 (unconditional) of utterance u share the step's ids; check_stopping's
    // end-of-stream countdown `delay` [R / 2]; logits [R][n_out * vocab] per row, combined by cfg_scale (src/util.cpp:175-200) into logits_cfg [R / 2][n_out * vocab]
    int pad, max_delay; float cfg; int * delay; float * logits_cfg;
    float * att_part;                             // split attention: per (row, head, chunk) hd + 4 floats: max, sum of exps, (pad), unnormalised P.V    // PKM_ORPHEUS argmax over a 150k vocabulary: amax_ch partial (value, index) pairs per row, combined by the row's next rows phase
    int n_stages, a_bytes;                        // shared-memory layout: ring stages, bytes of the activation / attention-scratch region
    unsigned * bar;                               // grid-barrier arrival counter, zeroed before every launch
    int * d_step; int step_begin, n_steps;        // steps [step_begin, step_begin + n_steps) run in this launch
    // rows of a step under the delay pattern + codebook embedding (parler generate_audio_tokens / parler_build_inp_embd; see delay_rows_kernel, codebook_embed_kernel)
    const int * first_pos; int * d_out; const int * d_teacher; int bos, eos, max_gen; int * seen; int * stopped; int * ids; int * row_pos;
    const float * tables; size_t tab_stride; const float * pos_embed; float * x0; size_t x0rep;
    // paged KV cache: layer l's pool at kv_pool + l * kv_layer_bytes; page p = [2 (k, v)][heads][PK_PAGE][hd] elements; page_table [R][max_pages]
    unsigned char * kv_pool; size_t kv_layer_bytes; const int * page_table; int max_pages;
    float * logits; float * logits_all;           // logits [R][n_out * vocab]; logits_all (optional) [n_steps_total][R][n_out * vocab]
    unsigned long long * prof; int prof_step;     // optional timeline of step prof_step: [n_ops][gridDim.x][8] %globaltimer ns (op begin, activations staged, barrier entered, barrier left, ns spent waiting for weight tiles, norm tile ready, -, -)
};

struct PkLaunch {
    const void * kfn = nullptr; size_t smem = 0;
#ifdef B2EMU
    std::function<void(const PkParams &)> kemu;
#endif
};
// shared-memory layout + kernel instantiation for a program; one cooperative launch of Pk.n_steps steps; the %globaltimer timeline; the prompt pass' K / V rows -> pages
int pk_configure(PkParams & Pk, bool kv_f32, int KA, int KAs, int Tscore, PkLaunch & L);
cudaError_t pk_launch(const PkLaunch & L, const PkParams & Pk, int grid, cudaStream_t st);
void pk_prof_begin(PkParams & Pk, size_t n_ops, int grid, cudaStream_t st);
void pk_prof_end(PkParams & Pk, const std::vector<PkOp> & ops, int grid, cudaStream_t st);
void pk_kv_import(bool kv_f32, const float * Kc, const float * Vc, size_t layer_stride, const int * row_src, const int * row_seq, const int * row_pos, const PkParams & P, int n_rows, int n_layers, cudaStream_t st);

}  // namespace b2

#ifdef B2_PDK_IMPLEMENTATION
#include "ar_kernels.cuh"

namespace b2 {
namespace {

// ---------------------------------------------------------------- primitives: mbarrier, bulk copy, named / grid barriers (PTX; functional models under B2EMU)
#ifdef B2EMU
struct PkBar { int pending, count, tx; unsigned phase; };
static inline void pk_mbar_flip(PkBar * b) { if (b->pending == 0 && b->tx == 0) { b->phase ^= 1u; b->pending = b->count; } b2emu::note_progress(); }
static inline void pk_mbar_init(PkBar * b, int count) { b->pending = b->count = count; b->tx = 0; b->phase = 0; }
static inline void pk_mbar_arrive(PkBar * b) { b->pending--; pk_mbar_flip(b); }
static inline void pk_mbar_expect_tx(PkBar * b, unsigned bytes) { b->tx += (int) bytes; b->pending--; pk_mbar_flip(b); }
static inline void pk_mbar_wait(PkBar * b, unsigned parity) { while (b->phase == parity) b2emu::yield_spin(); }
static inline void pk_bulk_g2s(void * dst, const void * src, unsigned bytes, PkBar * b) { memcpy(dst, src, bytes); b->tx -= (int) bytes; pk_mbar_flip(b); }
static inline void pk_bar_sync(int id, int n) { b2emu::named_bar(id, n); }
static inline unsigned pk_ld_acquire(const unsigned * p) { return *p; }
static inline void pk_red_release(unsigned * p, unsigned v) { *p += v; b2emu::note_progress(); }
static inline void pk_spin() { b2emu::yield_spin(); }
static inline void pk_fence_init() {}
#else
typedef uint64_t PkBar;
__device__ __forceinline__ uint32_t pk_smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void pk_mbar_init(PkBar * b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pk_smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void pk_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void pk_mbar_arrive(PkBar * b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pk_smem_u32(b)) : "memory"); }
__device__ __forceinline__ void pk_mbar_expect_tx(PkBar * b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pk_smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void pk_mbar_wait(PkBar * b, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "PK_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra PK_DONE;\n\t"
        "bra PK_WAIT;\n\t"
        "PK_DONE:\n\t"
        "}" ::"r"(pk_smem_u32(b)), "r"(parity) : "memory");
}
// TMA bulk copy (1-D): global -> this CTA's shared memory, completion counted in bytes on an mbarrier
__device__ __forceinline__ void pk_bulk_g2s(void * dst, const void * src, unsigned bytes, PkBar * b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pk_smem_u32(dst)), "l"(src), "r"(bytes), "r"(pk_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void pk_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ unsigned pk_ld_acquire(const unsigned * p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void pk_red_release(unsigned * p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void pk_spin() {}
#endif
#ifdef B2EMU
static inline unsigned long long pk_now() { return 0ull; }
#else
__device__ __forceinline__ unsigned long long pk_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

// position in the weight ring: stage index and the parity of its current use (producer and consumers walk the same sequence of tiles)
struct PkRingPos { int s; unsigned ph; };
__device__ __forceinline__ void pk_ring_next(PkRingPos & p, int S) { if (++p.s == S) { p.s = 0; p.ph ^= 1u; } }
// 16-byte shared-memory load as ld.shared (the operand pointers are derived from the dynamic shared array through several inlined calls: left to itself the compiler
// emitted generic loads for part of them)
#ifdef B2EMU
static inline uint4 pk_lds128(const void * p) { uint4 v; memcpy(&v, p, 16); return v; }
#else
__device__ __forceinline__ uint4 pk_lds128(const void * p) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(pk_smem_u32(p)));
    return v;
}
#endif

#ifdef B2EMU
static inline uint2 pk_lds64(const void * p) { uint2 v; memcpy(&v, p, 8); return v; }
// functional model of mma.sync.m16n8k32.s8: a0 / a1 = rows g / g + 8, k slots 4t .. 4t+3; a2 / a3 the same rows, slots 16 + 4t ..; b0 / b1 = column g, the same slots
static inline void pk_imma16832(int * c, unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    unsigned mine[6] = {a0, a1, a2, a3, b0, b1}, all[32][6];
    b2emu::warp_exchange(mine, 6, &all[0][0]);
    const int lane = b2emu_lane(), g = lane >> 2, t = lane & 3;
    auto by = [](unsigned v, int i) { return (int) (signed char) ((v >> (8 * i)) & 0xffu); };
    for (int ci = 0; ci < 4; ci++) {
        const int row = g + (ci >> 1) * 8, col = 2 * t + (ci & 1);
        int acc = c[ci];
        for (int tp = 0; tp < 4; tp++)
            for (int hi = 0; hi < 2; hi++)
                for (int i = 0; i < 4; i++) acc += by(all[(row & 7) * 4 + tp][(row >> 3) + 2 * hi], i) * by(all[col * 4 + tp][4 + hi], i);
        c[ci] = acc;
    }
}
#else
__device__ __forceinline__ uint2 pk_lds64(const void * p) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(pk_smem_u32(p)));
    return v;
}
__device__ __forceinline__ void pk_imma16832(int * c, unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
#endif

// every consumer thread of every CTA calls it; orders all global stores before it against all loads after it, grid-wide (the cooperative-groups pattern: block
// barrier, one thread releases / acquires at gpu scope, block barrier).  The counter only grows: generation e is complete when it reaches e * gridDim.x.
__device__ __forceinline__ void pk_grid_sync(unsigned * ctr, unsigned & epoch) {
    pk_bar_sync(1, PK_CONS);
    epoch++;
    if (threadIdx.x == 0) {
        pk_red_release(ctr, 1u);                              // release at gpu scope: orders every store the block made before the bar.sync above (cumulativity); no extra fence
        const unsigned target = epoch * gridDim.x;
        while (pk_ld_acquire(ctr) < target) pk_spin();
    }
    pk_bar_sync(1, PK_CONS);
}

// 8 consecutive cache elements as floats (one 16-byte load for fp16 pages, two for fp32 stores); L2-coherent loads: other CTAs wrote them earlier in this launch
__device__ __forceinline__ void pk_load8(const __half * p, float * v) {
    const uint4 u = __ldcg(reinterpret_cast<const uint4 *>(p));
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; i++) { __half2 h; memcpy(&h, &w[i], 4); const float2 f = __half22float2(h); v[2 * i] = f.x; v[2 * i + 1] = f.y; }      // (memcpy: no type-punned reads)
}
__device__ __forceinline__ void pk_load8(const float * p, float * v) {
    const float4 a = __ldcg(reinterpret_cast<const float4 *>(p)), b = __ldcg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void pk_store1(__half * p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void pk_store1(float * p, float v) { *p = v; }

template <typename KVT> __device__ __forceinline__ KVT * pk_page_row(const PkParams & P, int layer, int r, int pos, int kv, int h) {
    const int pg = __ldg(P.page_table + (size_t) r * P.max_pages + (pos >> 5));
    KVT * base = reinterpret_cast<KVT *>(P.kv_pool + (size_t) layer * P.kv_layer_bytes);
    return base + ((((size_t) pg * 2 + kv) * P.kv_heads + h) * PK_PAGE + (pos & (PK_PAGE - 1))) * P.hd;      // h: kv head
}

// ---------------------------------------------------------------- the producer: this CTA's weight tiles of one op, in the order the consumers use them
// unit u of segment sg -> first output row of its primary tile (and, for paired units, of the partner tile)
__device__ __forceinline__ void pk_unit_rows(const PkSeg & sg, int u, int hd, int & n0, int & n1) {
    const int lu = u - sg.unit0;
    if (sg.pair == PKP_ROPE) { const int per = hd >> 4, head = lu / per, j = lu - head * per; n0 = head * hd + j * 8; n1 = n0 + (hd >> 1); }     // hd / 2 rows of a head in units of 8
    else { n0 = lu * 8; n1 = n0; }
}

// The whole producer warp runs it: lane 0 waits for the stage and posts the byte count, lanes 0-7 each copy one row of the tile.  (One lane issuing all eight copies
// spent ~200 dependent instructions per tile on addresses and the per-copy uniform-register hand-off: 0.6 us per 16 KB tile, a 2.7 TB/s cap on a B200 whose copy
// engines stream 6.9 TB/s in this shape -- scripts/probes/tma_stream_probe.cu.)
__device__ __forceinline__ void pk_produce_gemv(const PkOp & op, unsigned char * ring, PkBar * full, PkBar * empty, int S, PkRingPos & rp, int ak, int hd) {      // op: the producer's own shared-memory copy
    const int lane = threadIdx.x & 31, row = lane & 7;
    const int K = op.K, nA = (K + ak - 1) / ak;
    if (op.norm != PKN_NONE) {                                 // the norm's weight | bias (K floats each) travel through the ring like a tile: in shared memory long before the op starts
        if (lane == 0) {
            pk_mbar_wait(&empty[rp.s], rp.ph ^ 1u);
            pk_mbar_expect_tx(&full[rp.s], (unsigned) ((op.nb ? 2 : 1) * K * 4));
            pk_bulk_g2s(ring + (size_t) rp.s * PK_STAGE, op.nw, (unsigned) (K * 4), &full[rp.s]);
            if (op.nb) pk_bulk_g2s(ring + (size_t) rp.s * PK_STAGE + (size_t) K * 4, op.nb, (unsigned) (K * 4), &full[rp.s]);
        }
        pk_ring_next(rp, S);
    }
    for (int a = 0; a < nA; a++) {
        const int kA0 = a * ak, kAn = K - kA0 < ak ? K - kA0 : ak;
        for (int u = (int) blockIdx.x; u < op.n_units; u += (int) gridDim.x) {
            const int sj = (op.nseg > 2 && u >= op.seg[2].unit0) ? 2 : ((op.nseg > 1 && u >= op.seg[1].unit0) ? 1 : 0);
            const PkSeg & sg = op.seg[sj];
            int n0, n1;
            pk_unit_rows(sg, u, hd, n0, n1);
            const int N = sg.N, pair = sg.pair;
            if (op.q8) {                                        // Q8_0: a tile = 8 rows x <= 1 024 int8 values (lanes 0-7) + their <= 32 fp16 block scales per row (lanes 8-15)
                const int r0q = n0 + row < N ? n0 + row : N - 1, r1q = pair ? (n1 + row < N ? n1 + row : N - 1) : r0q;
                const unsigned char * v0 = reinterpret_cast<const unsigned char *>(sg.W) + (size_t) r0q * K + kA0, * v1 = reinterpret_cast<const unsigned char *>(pair ? sg.Wp : sg.W) + (size_t) r1q * K + kA0;
                const __half * s0 = sg.Ws + (size_t) r0q * (K >> 5) + (kA0 >> 5), * s1 = (pair ? sg.Wps : sg.Ws) + (size_t) r1q * (K >> 5) + (kA0 >> 5);
                for (int kl = 0; kl < kAn; kl += PK_TK) {
                    const unsigned bytes = (unsigned) (kAn - kl < PK_TK ? kAn - kl : PK_TK), sbytes = bytes >> 4;      // (bytes / 32 blocks x 2 bytes)
                    for (int part = 0; part < (pair ? 2 : 1); part++) {
                        if (lane == 0) {
                            pk_mbar_wait(&empty[rp.s], rp.ph ^ 1u);
                            pk_mbar_expect_tx(&full[rp.s], 8u * (bytes + sbytes));
                        }
                        __syncwarp();
                        unsigned char * st = ring + (size_t) rp.s * PK_STAGE;
                        if (lane < 8) pk_bulk_g2s(st + (size_t) row * PK_ROWQ, (part ? v1 : v0) + kl, bytes, &full[rp.s]);
                        else if (lane < 16) pk_bulk_g2s(st + PK_QSC + (size_t) row * 64, (part ? s1 : s0) + (kl >> 5), sbytes, &full[rp.s]);
                        pk_ring_next(rp, S);
                    }
                }
                continue;
            }
            const __half * Wl = sg.Wl;
            const int nparts = pair ? 2 : (Wl ? 2 : 1);         // paired unit: primary + partner tile; split matrix: high + low plane
            // this lane's row of the two tiles (rows past N re-read the last row and are never stored)
            const int r0 = n0 + row < N ? n0 + row : N - 1, r1 = pair ? (n1 + row < N ? n1 + row : N - 1) : r0;
            const __half * p0 = sg.W + (size_t) r0 * K + kA0, * p1 = (pair ? sg.Wp : (Wl ? Wl : sg.W)) + (size_t) r1 * K + kA0;
            for (int kl = 0; kl < kAn; kl += PK_TK) {
                const unsigned bytes = (unsigned) ((kAn - kl < PK_TK ? kAn - kl : PK_TK) * 2);      // within the chunk the consumers have staged
                for (int part = 0; part < nparts; part++) {
                    if (lane == 0) {
                        pk_mbar_wait(&empty[rp.s], rp.ph ^ 1u);
                        pk_mbar_expect_tx(&full[rp.s], 8u * bytes);
                    }
                    __syncwarp();                               // the stage is free and its byte count posted before any row lands
                    if (lane < 8) pk_bulk_g2s(ring + (size_t) rp.s * PK_STAGE + (size_t) row * PK_ROWB, (part ? p1 : p0) + kl, bytes, &full[rp.s]);
                    pk_ring_next(rp, S);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- consumers: GEMV phase
// stage rows [0, 16) x columns [k0, k0 + kn) of X as fp16 into sA (pitch halves per row; rows >= R are zero), optionally LayerNorm'd (mean / rstd per row in
// registers of the owning warp: warp w owns rows w and w + 8) and, for split matrices, their scaled low halves into sAl.
__device__ __forceinline__ void pk_stage_rows(const PkOp & op, const float * X, const float * snw, int R, int k0, int kn, __half * sA, __half * sAl, int pitch, const float * mean, const float * rstd) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float * snb = snw ? snw + op.K : nullptr;
    const int n4 = kn >> 2;                                   // float4 per row (kn % 256 == 0 -> n4 % 64 == 0)
    for (int j0 = 0; j0 < n4; j0 += 32 * 8) {                 // 8 float4 per lane and row in flight, for both rows
        float4 v[2][8];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const int r = warp + 8 * rr;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int j = j0 + u * 32 + lane;
                v[rr][u] = (r < R && j < n4) ? __ldcg(reinterpret_cast<const float4 *>(X + (size_t) r * op.ldx + k0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const int r = warp + 8 * rr;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int j = j0 + u * 32 + lane;
                if (j >= n4) continue;
                float4 x = v[rr][u];
                if (op.norm == PKN_RMS && r < R) {             // ggml_rms_norm then * weight: (x * scale) * w
                    const float4 w = reinterpret_cast<const float4 *>(snw + k0)[j];
                    x.x = (x.x * rstd[rr]) * w.x; x.y = (x.y * rstd[rr]) * w.y; x.z = (x.z * rstd[rr]) * w.z; x.w = (x.w * rstd[rr]) * w.w;
                } else if (op.norm == PKN_LAYER && r < R) {    // ggml_norm then * weight + bias (parler build_norm): ((x - mean) * scale) * w + b
                    const float4 w = reinterpret_cast<const float4 *>(snw + k0)[j], b = reinterpret_cast<const float4 *>(snb + k0)[j];
                    x.x = ((x.x - mean[rr]) * rstd[rr]) * w.x + b.x; x.y = ((x.y - mean[rr]) * rstd[rr]) * w.y + b.y;
                    x.z = ((x.z - mean[rr]) * rstd[rr]) * w.z + b.z; x.w = ((x.w - mean[rr]) * rstd[rr]) * w.w + b.w;
                }
                const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
                __half2 * d = reinterpret_cast<__half2 *>(sA + (size_t) r * pitch + 4 * j);
                d[0] = h01; d[1] = h23;
                if (sAl) {
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    __half2 * dl = reinterpret_cast<__half2 *>(sAl + (size_t) r * pitch + 4 * j);
                    dl[0] = __floats2half2_rn((x.x - f01.x) * GM_LO_SCALE, (x.y - f01.y) * GM_LO_SCALE);
                    dl[1] = __floats2half2_rn((x.z - f23.x) * GM_LO_SCALE, (x.w - f23.y) * GM_LO_SCALE);
                }
            }
        }
    }
}


// The fast path of the two functions below for K <= 1 024 (every normalised GEMV of the models here): rows warp and warp + 8 are loaded ONCE (8 float4 per lane and row,
// all in flight together with the norm's weight / bias), the statistics come from the registers, the normalised fp16 rows go to shared memory -- one L2 round trip
// instead of three.  Statistics: ggml_norm accumulates float values in double; here the four values of a float4 are added in float first (error <= 2^-23 of the
// partial), then double -- 4x fewer FP64 operations, the float mean / variance come out the same except in rare last-bit cases.
__device__ __forceinline__ void pk_stage_fast(const PkOp & op, const float * X, const float * snw, int R, __half * sA, __half * sAl, int pitch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, K = op.K, n4 = K >> 2;
    const float * snb = snw ? snw + K : nullptr;
    const int r0 = warp, r1 = warp + 8;
    const bool norm = op.norm == PKN_LAYER;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[2][8], w[4], b[4];
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const int j = u * 32 + lane;
        const bool in = j < n4;
        v[0][u] = (in && r0 < R) ? __ldcg(reinterpret_cast<const float4 *>(X + (size_t) r0 * op.ldx) + j) : z4;
        v[1][u] = (in && r1 < R) ? __ldcg(reinterpret_cast<const float4 *>(X + (size_t) r1 * op.ldx) + j) : z4;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {                              // the norm's weight / bias: from the ring stage the producer filled ahead of time
        const int j = u * 32 + lane;
        w[u] = (norm && j < n4) ? reinterpret_cast<const float4 *>(snw)[j] : z4;
        b[u] = (norm && j < n4) ? reinterpret_cast<const float4 *>(snb)[j] : z4;
    }
    float m0 = 0.f, m1 = 0.f, i0 = 1.f, i1 = 1.f;
    if (norm) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++) { s0 += (double) ((v[0][u].x + v[0][u].y) + (v[0][u].z + v[0][u].w)); s1 += (double) ((v[1][u].x + v[1][u].y) + (v[1][u].z + v[1][u].w)); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        m0 = (float) (s0 / (double) K); m1 = (float) (s1 / (double) K);
        double q0 = 0.0, q1 = 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (u * 32 + lane < n4) {
                float a = v[0][u].x - m0, c = v[0][u].y - m0, d = v[0][u].z - m0, e = v[0][u].w - m0;
                q0 += (double) ((a * a + c * c) + (d * d + e * e));
                a = v[1][u].x - m1; c = v[1][u].y - m1; d = v[1][u].z - m1; e = v[1][u].w - m1;
                q1 += (double) ((a * a + c * c) + (d * d + e * e));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o); }
        i0 = 1.0f / sqrtf((float) (q0 / (double) K) + op.eps); i1 = 1.0f / sqrtf((float) (q1 / (double) K) + op.eps);
    }
#pragma unroll
    for (int half = 0; half < 2; half++) {
        if (half) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = (4 + u) * 32 + lane;
                w[u] = (norm && j < n4) ? reinterpret_cast<const float4 *>(snw)[j] : z4;
                b[u] = (norm && j < n4) ? reinterpret_cast<const float4 *>(snb)[j] : z4;
            }
        }
#pragma unroll
        for (int uu = 0; uu < 4; uu++) {
            const int u = half * 4 + uu, j = u * 32 + lane;
            if (j >= n4) continue;
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int r = warp + 8 * rr;
                float4 x = v[rr][u];
                if (norm) {                                    // ggml_norm then * weight + bias (parler build_norm): ((x - mean) * scale) * w + b
                    const float m = rr ? m1 : m0, is = rr ? i1 : i0;
                    x.x = ((x.x - m) * is) * w[uu].x + b[uu].x; x.y = ((x.y - m) * is) * w[uu].y + b[uu].y;
                    x.z = ((x.z - m) * is) * w[uu].z + b[uu].z; x.w = ((x.w - m) * is) * w[uu].w + b[uu].w;
                }
                if (r >= R) x = z4;
                const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
                __half2 * d = reinterpret_cast<__half2 *>(sA + (size_t) r * pitch + 4 * j);
                d[0] = h01; d[1] = h23;
                if (sAl) {
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    __half2 * dl = reinterpret_cast<__half2 *>(sAl + (size_t) r * pitch + 4 * j);
                    dl[0] = __floats2half2_rn((x.x - f01.x) * GM_LO_SCALE, (x.y - f01.y) * GM_LO_SCALE);
                    dl[1] = __floats2half2_rn((x.z - f23.x) * GM_LO_SCALE, (x.w - f23.y) * GM_LO_SCALE);
                }
            }
        }
    }
}

// RMSNorm'd rows of up to NV * 128 columns (llama-style models: the whole hidden row), one row of the warp at a time: the row's NV float4 per lane are all in
// flight at once and stay in registers for the statistics and the normalisation -- one L2 round trip per row.  ggml_rms_norm: float squares accumulated in double
// (here: the four squares of a float4 are added in float first, as in pk_stage_fast), scale = 1 / sqrtf(mean + eps), then (x * scale) * weight (orpheus model.cpp:122-125).
template <int NV>
__device__ __forceinline__ void pk_stage_rms(const PkOp & op, const float * X, const float * snw, int R, __half * sA, int pitch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, K = op.K, n4 = K >> 2;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int rr = 0; rr < 2; rr++) {
        const int r = warp + 8 * rr;
        float4 v[NV];
#pragma unroll
        for (int u = 0; u < NV; u++) { const int j = u * 32 + lane; v[u] = (j < n4 && r < R) ? __ldcg(reinterpret_cast<const float4 *>(X + (size_t) r * op.ldx) + j) : z4; }
        double ss = 0.0;
#pragma unroll
        for (int u = 0; u < NV; u++) ss += (double) ((v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w));      // four squares in float, then double (see pk_stage_fast)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float sc = 1.0f / sqrtf((float) (ss / (double) K) + op.eps);
#pragma unroll
        for (int u = 0; u < NV; u++) {
            const int j = u * 32 + lane;
            if (j >= n4) continue;
            const float4 w = reinterpret_cast<const float4 *>(snw)[j];
            const __half2 h01 = __floats2half2_rn((v[u].x * sc) * w.x, (v[u].y * sc) * w.y), h23 = __floats2half2_rn((v[u].z * sc) * w.z, (v[u].w * sc) * w.w);
            uint2 pk; memcpy(&pk.x, &h01, 4); memcpy(&pk.y, &h23, 4);
            *reinterpret_cast<uint2 *>(sA + (size_t) r * pitch + 4 * j) = pk;      // one 8-byte store per lane: consecutive lanes, no bank conflict
        }
    }
}

// fp16 input rows (written by the previous phase): straight 16-byte copies into the operand buffer; warp w copies rows w and w + 8, up to 12 loads per row and lane
// all in flight (no index arithmetic per load: the row / column split by division was a quarter of the staging pass' instructions)
__device__ __forceinline__ void pk_stage_h16(const PkOp & op, const __half * X16, int R, int k0, int kn, __half * sA, int pitch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n8 = kn >> 3;      // uint4 (8 halves) per row
    const uint4 * src0 = reinterpret_cast<const uint4 *>(X16 + (size_t) warp * op.ldx + k0), * src1 = reinterpret_cast<const uint4 *>(X16 + (size_t) (warp + 8) * op.ldx + k0);
    uint4 * dst0 = reinterpret_cast<uint4 *>(sA + (size_t) warp * pitch), * dst1 = reinterpret_cast<uint4 *>(sA + (size_t) (warp + 8) * pitch);
    const bool in0 = warp < R, in1 = warp + 8 < R;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int j0 = lane; j0 < n8; j0 += 32 * 12) {
        uint4 v0[12], v1[12];
#pragma unroll
        for (int u = 0; u < 12; u++) { const int j = j0 + 32 * u; v0[u] = (in0 && j < n8) ? __ldcg(src0 + j) : z; v1[u] = (in1 && j < n8) ? __ldcg(src1 + j) : z; }
#pragma unroll
        for (int u = 0; u < 12; u++) { const int j = j0 + 32 * u; if (j < n8) { dst0[j] = v0[u]; dst1[j] = v1[u]; } }
    }
}

// mean and 1/sqrt(var + eps) of rows warp, warp + 8 over all K columns, ggml_norm's way: float values, double accumulators (ggml-cpu.c:7114-7163)
__device__ __forceinline__ void pk_row_stats(const PkOp & op, const float * X, int R, float * mean, float * rstd) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n4 = op.K >> 2;
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const int r = warp + 8 * rr;
        mean[rr] = 0.f; rstd[rr] = 0.f;
        if (r >= R) continue;                                   // warp-uniform
        const float4 * row = reinterpret_cast<const float4 *>(X + (size_t) r * op.ldx);
        if (op.norm == PKN_RMS) {                               // ggml_rms_norm: float squares accumulated in double, scale = 1 / sqrtf(mean + eps)
            double ss = 0.0;
            for (int j0 = lane; j0 < n4; j0 += 32 * 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = j0 + 32 * u < n4 ? __ldcg(row + j0 + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 8; u++) ss += (double) (v[u].x * v[u].x) + (double) (v[u].y * v[u].y) + (double) (v[u].z * v[u].z) + (double) (v[u].w * v[u].w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            mean[rr] = 0.f; rstd[rr] = 1.0f / sqrtf((float) (ss / (double) op.K) + op.eps);
            continue;
        }
        double s = 0.0;
        for (int j0 = lane; j0 < n4; j0 += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = j0 + 32 * u < n4 ? __ldcg(row + j0 + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; u++) s += (double) v[u].x + (double) v[u].y + (double) v[u].z + (double) v[u].w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float m = (float) (s / (double) op.K);
        double s2 = 0.0;
        for (int j0 = lane; j0 < n4; j0 += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = j0 + 32 * u < n4 ? __ldcg(row + j0 + 32 * u) : make_float4(m, m, m, m);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float a = v[u].x - m, b = v[u].y - m, c = v[u].z - m, d = v[u].w - m;
                s2 += (double) (a * a) + (double) (b * b) + (double) (c * c) + (double) (d * d);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        const float var = (float) (s2 / (double) op.K);
        mean[rr] = m; rstd[rr] = 1.0f / sqrtf(var + op.eps);
    }
}

template <typename KVT>
__device__ __forceinline__ void pk_epilogue(const PkParams & P, const PkOp & op, const PkSeg & sg, int r, int n, float a, float resv, int step_abs, const unsigned long long * skv) {
    const int nrep = sg.yrep ? PK_REP : 1;
    switch (sg.epi) {
        case PKE_GELU: { const __half hv = __float2half_rn(gelu_f16lut(a)); for (int c = 0; c < nrep; c++) sg.Y16[sg.yrep * c + (size_t) r * sg.ldy + n] = hv; break; }
        case PKE_RES:  { const float v = a + resv; for (int c = 0; c < nrep; c++) sg.Y[sg.yrep * c + (size_t) r * sg.ldy + n] = v; break; }
        case PKE_KV: {                                         // skv[r]: element offset of (this row's page, its slot in the page), looked up once per op
            const int h = n / P.hd, d = n - h * P.hd;
            KVT * base = reinterpret_cast<KVT *>(P.kv_pool + (size_t) op.layer * P.kv_layer_bytes);
            pk_store1(base + skv[r] + ((size_t) sg.kv * P.kv_heads + h) * PK_PAGE * P.hd + d, a);
            break;
        }
        case PKE_LOGITS:
            sg.Y[(size_t) r * sg.ldy + n] = a;
            if (P.logits_all && P.model != PKM_DIA) P.logits_all[((size_t) step_abs * P.R + r) * sg.ldy + n] = a;      // (Dia keeps the cfg-combined logits: pk_argmax_dia)
            break;
        default: sg.Y[(size_t) r * sg.ldy + n] = a; break;
    }
}

// joint epilogue of a paired unit.  PKP_ROPE: a0 = the row's value at in-head index i < hd / 2 (output n_lo), a1 = its partner at i + hd / 2 (output n_hi): NeoX rotation with
// the step's (cos, sin) table (ggml_rope_ext mode 2: rope_append_kernel), q to Y, k into the row's cache slot.  PKP_SWIGLU: a0 = gate, a1 = up: ggml_silu(gate) * up as fp16.
template <typename KVT>
__device__ __forceinline__ void pk_epilogue_pair(const PkParams & P, const PkOp & op, const PkSeg & sg, int r, int n_lo, int n_hi, float a0, float a1, const unsigned long long * skv) {
    if (sg.epi == PKE_SWIGLU) {
        const __half hv = __float2half_rn((a0 / (1.0f + expf(-a0))) * a1);
        for (int c = 0; c < (sg.yrep ? PK_REP : 1); c++) sg.Y16[sg.yrep * c + (size_t) r * sg.ldy + n_lo] = hv;
        return;
    }
    const int hd = P.hd, half = hd >> 1, h = n_lo / hd, i = n_lo - h * hd;
    const float2 cs = __ldcg(P.rope_cs + (size_t) r * half + i);
    const float y0 = a0 * cs.x - a1 * cs.y, y1 = a0 * cs.y + a1 * cs.x;
    if (sg.epi == PKE_ROPE_Q) { sg.Y[(size_t) r * sg.ldy + n_lo] = y0; sg.Y[(size_t) r * sg.ldy + n_hi] = y1; }
    else {
        KVT * dst = reinterpret_cast<KVT *>(P.kv_pool + (size_t) op.layer * P.kv_layer_bytes) + skv[r] + (size_t) h * PK_PAGE * hd;      // K plane (kv = 0) of the row's page slot
        pk_store1(dst + i, y0); pk_store1(dst + i + half, y1);
    }
}

// ---- one tile: this warp's k-slice (ks halves) of 8 output rows against the staged activation rows.  Fragments: 16 bytes per lane of the weight row g = lane / 4 and of
// the activation rows g, g + 8 feed two m16n8k16 MMAs (see gemv_mma_body in ar_kernels.cuh for the layout argument).
// plain: the two MMAs of a 32-column step go to two accumulators (half the dependent chain), added when the unit is done
// HI = false: at most 8 activation rows (rows 8 .. 15 of the M = 16 operand are zero): their fragments are not loaded -- a third less shared-memory traffic per tile
template <bool HI>
__device__ __forceinline__ void pk_mma_plain(float * cA, float * cB, const __half * wt, const __half * xa, const __half * xb, int ks) {
#pragma unroll 4
    for (int k = 0; k < ks; k += 32) {
        const uint4 wv = pk_lds128(wt + k), a = pk_lds128(xa + k), b = HI ? pk_lds128(xb + k) : make_uint4(0u, 0u, 0u, 0u);
        const unsigned f0[4] = {a.x, b.x, a.y, b.y}, f1[4] = {a.z, b.z, a.w, b.w};
        mma16816_f16f32(cA, f0, wv.x, wv.y);
        mma16816_f16f32(cB, f1, wv.z, wv.w);
    }
}
// split matrices (W = hi + lo / 2^11, x = xh + xl / 2^11), high plane: c += xh.Wh, cl += xl.Wh; the low plane's cl += xh.Wl goes through pk_mma_one
template <bool HI>
__device__ __forceinline__ void pk_mma_split_hi(float * c, float * cl, const __half * wt, const __half * xa, const __half * xb, const __half * xla, const __half * xlb, int ks) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 2
    for (int k = 0; k < ks; k += 32) {
        const uint4 wv = pk_lds128(wt + k), a = pk_lds128(xa + k), b = HI ? pk_lds128(xb + k) : z, la = pk_lds128(xla + k), lb = HI ? pk_lds128(xlb + k) : z;
        const unsigned f0[4] = {a.x, b.x, a.y, b.y}, f1[4] = {a.z, b.z, a.w, b.w}, l0[4] = {la.x, lb.x, la.y, lb.y}, l1[4] = {la.z, lb.z, la.w, lb.w};
        mma16816_f16f32(c, f0, wv.x, wv.y);
        mma16816_f16f32(cl, l0, wv.x, wv.y);
        mma16816_f16f32(c, f1, wv.z, wv.w);
        mma16816_f16f32(cl, l1, wv.z, wv.w);
    }
}
template <bool HI>
__device__ __forceinline__ void pk_mma_one(float * c, const __half * wt, const __half * xa, const __half * xb, int ks) {
#pragma unroll 4
    for (int k = 0; k < ks; k += 32) {
        const uint4 wv = pk_lds128(wt + k), a = pk_lds128(xa + k), b = HI ? pk_lds128(xb + k) : make_uint4(0u, 0u, 0u, 0u);
        const unsigned f0[4] = {a.x, b.x, a.y, b.y}, f1[4] = {a.z, b.z, a.w, b.w};
        mma16816_f16f32(c, f0, wv.x, wv.y);
        mma16816_f16f32(c, f1, wv.z, wv.w);
    }
}

enum { PKT_PLAIN = 0, PKT_SPLIT = 1, PKT_PAIR = 2 };
// all tiles of one unit within the staged activation chunk of kAn columns.  c / cl: PLAIN two partial accumulators of the same sums; SPLIT main and cross terms; PAIR the
// unit's primary and partner tile.  Two tiles are worked on at a time (PLAIN: two consecutive k tiles; SPLIT / PAIR: the two tiles of one k tile): with 2 consumer
// warps per scheduler a single tile's chain of shared-memory loads and dependent MMAs left the tensor pipe idle most of the time (issue slots 22 % busy, ncu r2h).
template <int MODE, bool HI>
__device__ __forceinline__ void pk_unit_tiles(float * c, float * cl, unsigned char * ring, const __half * sA, const __half * sAl, int pitch, PkBar * full, PkBar * empty, PkRingPos & rp, int S, int kAn, unsigned long long * pr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t8 = (lane & 3) * 8;
    auto wait_full = [&](const PkRingPos & q) {
        if (pr) { const unsigned long long t0 = pk_now(); pk_mbar_wait(&full[q.s], q.ph); pr[4] += pk_now() - t0; }      // (timeline: ns thread 0 waited for weight tiles)
        else pk_mbar_wait(&full[q.s], q.ph);
    };
    auto wtile = [&](const PkRingPos & q, int ks) { return reinterpret_cast<const __half *>(ring + (size_t) q.s * PK_STAGE) + (size_t) g * (PK_TK + PK_PAD) + t8 + warp * ks; };
    auto release = [&](const PkRingPos & q) { if (lane == 0) pk_mbar_arrive(&empty[q.s]); };      // (after a __syncwarp: this warp is done reading the stage)
    if (MODE == PKT_PLAIN) {
        float c2[4] = {0.f, 0.f, 0.f, 0.f}, cl2[4] = {0.f, 0.f, 0.f, 0.f};
        int kt0 = 0;
        for (; kt0 + PK_TK < kAn; kt0 += 2 * PK_TK) {           // two k tiles at a time: the first is a full tile
            const int ktn1 = kAn - kt0 - PK_TK < PK_TK ? kAn - kt0 - PK_TK : PK_TK, ks0 = PK_TK >> 3, ks1 = ktn1 >> 3;
            const PkRingPos q0 = rp; pk_ring_next(rp, S);
            const PkRingPos q1 = rp; pk_ring_next(rp, S);
            const __half * xa0 = sA + (size_t) g * pitch + kt0 + warp * ks0 + t8, * xa1 = sA + (size_t) g * pitch + kt0 + PK_TK + warp * ks1 + t8;
            wait_full(q0); wait_full(q1);
            pk_mma_plain<HI>(c, cl, wtile(q0, ks0), xa0, xa0 + (size_t) 8 * pitch, ks0);
            pk_mma_plain<HI>(c2, cl2, wtile(q1, ks1), xa1, xa1 + (size_t) 8 * pitch, ks1);
            __syncwarp();
            release(q0); release(q1);
        }
        if (kt0 < kAn) {                                        // an odd tile left
            const int ktn = kAn - kt0 < PK_TK ? kAn - kt0 : PK_TK, ks = ktn >> 3;
            const PkRingPos q0 = rp; pk_ring_next(rp, S);
            const __half * xa = sA + (size_t) g * pitch + kt0 + warp * ks + t8;
            wait_full(q0);
            pk_mma_plain<HI>(c, cl, wtile(q0, ks), xa, xa + (size_t) 8 * pitch, ks);
            __syncwarp();
            release(q0);
        }
#pragma unroll
        for (int e = 0; e < 4; e++) { c[e] += c2[e]; cl[e] += cl2[e]; }
    } else {
        for (int kt0 = 0; kt0 < kAn; kt0 += PK_TK) {
            const int ktn = kAn - kt0 < PK_TK ? kAn - kt0 : PK_TK, ks = ktn >> 3, ko = kt0 + warp * ks + t8;
            const PkRingPos q0 = rp; pk_ring_next(rp, S);
            const PkRingPos q1 = rp; pk_ring_next(rp, S);
            const __half * xa = sA + (size_t) g * pitch + ko, * xb = xa + (size_t) 8 * pitch;
            wait_full(q0); wait_full(q1);
            if (MODE == PKT_PAIR) { pk_mma_one<HI>(c, wtile(q0, ks), xa, xb, ks); pk_mma_one<HI>(cl, wtile(q1, ks), xa, xb, ks); }
            else { pk_mma_split_hi<HI>(c, cl, wtile(q0, ks), xa, xb, sAl + (size_t) g * pitch + ko, sAl + (size_t) (g + 8) * pitch + ko, ks); pk_mma_one<HI>(cl, wtile(q1, ks), xa, xb, ks); }
            __syncwarp();
            release(q0); release(q1);
        }
    }
}

// the eight warps' partial 16 x 8 tiles of a unit summed in warp order, then the epilogue.  c0, c1 = row g, columns 2t, 2t + 1; c2, c3 = row g + 8.  The scratch
// alternates between two buffers, so one block barrier per unit is enough: a warp that writes buffer b again has passed the next unit's barrier, which every warp
// reaches only after its reads of b.
template <typename KVT>
__device__ __forceinline__ void pk_unit_finish(const PkParams & P, const PkOp & op, const PkSeg & sg, float * c, float * cl, int mode, float * red, unsigned & rb, int n0, int n1, float resv, int step_abs,
                                               const unsigned long long * skv) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    if (mode == PKT_PLAIN) { for (int e = 0; e < 4; e++) c[e] += cl[e]; }
    else if (mode == PKT_SPLIT) { for (int e = 0; e < 4; e++) c[e] += cl[e] * (1.0f / GM_LO_SCALE); }
    float * buf = red + rb * PK_RED_FLOATS;
    rb ^= 1u;
    float * my = buf + warp * 128;
    *reinterpret_cast<float2 *>(my + g * 8 + 2 * tq) = make_float2(c[0], c[1]);
    *reinterpret_cast<float2 *>(my + (g + 8) * 8 + 2 * tq) = make_float2(c[2], c[3]);
    if (mode == PKT_PAIR) {
        float * my2 = my + 8 * 128;
        *reinterpret_cast<float2 *>(my2 + g * 8 + 2 * tq) = make_float2(cl[0], cl[1]);
        *reinterpret_cast<float2 *>(my2 + (g + 8) * 8 + 2 * tq) = make_float2(cl[2], cl[3]);
    }
    pk_bar_sync(1, PK_CONS);
    if (tid < 128) {
        const int r = tid >> 3, col = tid & 7;
        float sum = 0.f, sum2 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) sum += buf[w * 128 + tid];
        if (mode == PKT_PAIR) {
#pragma unroll
            for (int w = 0; w < 8; w++) sum2 += buf[8 * 128 + w * 128 + tid];
        }
        if (r < P.R && n0 + col < sg.N) {
            if (mode == PKT_PAIR) pk_epilogue_pair<KVT>(P, op, sg, r, n0 + col, n1 + col, sum, sum2, skv);
            else pk_epilogue<KVT>(P, op, sg, r, n0 + col, sum, resv, step_abs, skv);
        }
    }
}

template <typename KVT> __device__ __forceinline__ void pk_prefetch_kv(const PkParams & P, int layer, int step, const int * sfp, const int * spt);

template <typename KVT>
__device__ __forceinline__ void pk_gemv(const PkParams & P, const PkOp & op, unsigned char * ring, __half * sA, float * red, PkBar * full, PkBar * empty, PkRingPos & rp, int step_abs,
                                        unsigned long long * pr, unsigned long long * skv, const int * sfp, const int * spt) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int K = op.K, R = P.R, S = P.n_stages, ak = P.ak;
    const bool split = op.seg[0].Wl != nullptr;
    const int nA = (K + ak - 1) / ak;
    const int pitch = (K < ak ? K : ak) + PK_PAD;
    __half * sAl = split ? sA + (size_t) 16 * pitch : nullptr;
    float mean[2] = {0.f, 0.f}, rstd[2] = {0.f, 0.f};
    unsigned long long kvoff = 0ull;                           // the cache slot of row tid (q|k|v phase): requested now, parked in shared memory behind the staging pass
    if (op.kv_prefetch && tid < R) {
        const int pos = sfp[tid] + step_abs - P.pos_off;        // the position this step appends (the page table and the first positions sit in shared memory: see pdk_kernel)
        kvoff = (unsigned long long) spt[tid * P.max_pages + (pos >> 5)] * ((size_t) 2 * P.kv_heads * PK_PAGE * P.hd) + (size_t) (pos & (PK_PAGE - 1)) * P.hd;
    }
    if (op.kv_prefetch) pk_prefetch_kv<KVT>(P, op.layer, step_abs, sfp, spt);
    const bool fast = K <= 1024 && op.norm != PKN_RMS;         // one chunk, rows in registers: load + statistics + normalise in one pass
    const size_t xoff = op.xrep * (size_t) (blockIdx.x % PK_REP);      // this CTA's copy of the input rows
    const float * X = op.X ? op.X + xoff : nullptr; const __half * X16 = op.X16 ? op.X16 + xoff : nullptr;
    const float * snw = nullptr; int norm_stage = -1;
    if (op.norm != PKN_NONE) {                                 // the op's first "tile": the norm's weight | bias
        norm_stage = rp.s;
        pk_mbar_wait(&full[norm_stage], rp.ph);
        snw = reinterpret_cast<const float *>(ring + (size_t) norm_stage * PK_STAGE);
        pk_ring_next(rp, S);
        if (pr) pr[5] = pk_now();
    }
    const bool fast_rms = op.norm == PKN_RMS && !op.X16 && !split && K <= 3072 && K <= ak;      // the whole row in one chunk, in registers
    if (op.norm != PKN_NONE && !fast && !fast_rms && !op.X16) pk_row_stats(op, X, R, mean, rstd);
    auto stage = [&](int a) {
        const int kA0 = a * ak, kAn = K - kA0 < ak ? K - kA0 : ak;
        if (op.X16) pk_stage_h16(op, X16, R, kA0, kAn, sA, pitch);
        else if (fast) pk_stage_fast(op, X, snw, R, sA, sAl, pitch);
        else if (fast_rms) { if (K <= 1024) pk_stage_rms<8>(op, X, snw, R, sA, pitch); else pk_stage_rms<24>(op, X, snw, R, sA, pitch); }
        else pk_stage_rows(op, X, snw, R, kA0, kAn, sA, sAl, pitch, mean, rstd);
        if (a == 0 && op.kv_prefetch && tid < R) skv[tid] = kvoff;
        if (norm_stage >= 0 && a + 1 == nA) { __syncwarp(); if (lane == 0) pk_mbar_arrive(&empty[norm_stage]); }      // this warp is done with the norm's weights
        pk_bar_sync(1, PK_CONS);
        if (pr && a == 0) pr[1] = pk_now();                    // (GEMV phases: activations staged)
        return kAn;
    };
    auto seg_of = [&](int u) -> const PkSeg & { return op.seg[(op.nseg > 2 && u >= op.seg[2].unit0) ? 2 : ((op.nseg > 1 && u >= op.seg[1].unit0) ? 1 : 0)]; };
    auto res_of = [&](const PkSeg & sg, int n0) -> float {      // the epilogue's residual element of thread tid < 128, requested before the tiles are consumed
        if (sg.epi != PKE_RES || tid >= 128) return 0.f;
        const int r = tid >> 3, n = n0 + (tid & 7);
        return (r < R && n < sg.N) ? __ldcg(sg.res + sg.yrep * (size_t) (blockIdx.x % PK_REP) + (size_t) r * sg.ldy + n) : 0.f;
    };
    unsigned rb = 0;
    const bool hi = R > 8;                                     // rows 8 .. 15 of the MMA's M = 16 exist
    if (nA == 1) {                                             // the whole k extent staged at once: units one after the other
        const int kAn = stage(0);
        for (int u = (int) blockIdx.x; u < op.n_units; u += (int) gridDim.x) {
            const PkSeg & sg = seg_of(u);
            int n0, n1;
            pk_unit_rows(sg, u, P.hd, n0, n1);
            const float resv = res_of(sg, n0);
            float c[4] = {0.f, 0.f, 0.f, 0.f}, cl[4] = {0.f, 0.f, 0.f, 0.f};
            const int mode = sg.pair != PKP_NONE ? PKT_PAIR : (split ? PKT_SPLIT : PKT_PLAIN);
            if (hi) {
                if (mode == PKT_PAIR) pk_unit_tiles<PKT_PAIR, true>(c, cl, ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
                else if (mode == PKT_SPLIT) pk_unit_tiles<PKT_SPLIT, true>(c, cl, ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
                else pk_unit_tiles<PKT_PLAIN, true>(c, cl, ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
            } else {
                if (mode == PKT_PAIR) pk_unit_tiles<PKT_PAIR, false>(c, cl, ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
                else if (mode == PKT_SPLIT) pk_unit_tiles<PKT_SPLIT, false>(c, cl, ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
                else pk_unit_tiles<PKT_PLAIN, false>(c, cl, ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
            }
            pk_unit_finish<KVT>(P, op, sg, c, cl, mode, red, rb, n0, n1, resv, step_abs, skv);
        }
    } else {                                                   // several k chunks (down projections): at most 3 units per CTA (host-checked), their accumulators live across the chunks
        float acc[3][4], accl[3][4];
#pragma unroll
        for (int i = 0; i < 3; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; accl[i][0] = accl[i][1] = accl[i][2] = accl[i][3] = 0.f; }
        for (int a = 0; a < nA; a++) {
            if (a) pk_bar_sync(1, PK_CONS);                    // every warp is done with the previous chunk
            const int kAn = stage(a);
#pragma unroll
            for (int ui = 0; ui < 3; ui++) {
                const int u = (int) blockIdx.x + ui * (int) gridDim.x;
                if (u < op.n_units) {
                    const PkSeg & sg = seg_of(u);
                    const int n0 = (u - sg.unit0) * 8;          // (paired units only in single-chunk phases: host-checked)
                    const float resv = a + 1 == nA ? res_of(sg, n0) : 0.f;
                    if (split) pk_unit_tiles<PKT_SPLIT, true>(acc[ui], accl[ui], ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
                    else if (hi) pk_unit_tiles<PKT_PLAIN, true>(acc[ui], accl[ui], ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
                    else pk_unit_tiles<PKT_PLAIN, false>(acc[ui], accl[ui], ring, sA, sAl, pitch, full, empty, rp, S, kAn, pr);
                    if (a + 1 == nA) pk_unit_finish<KVT>(P, op, sg, acc[ui], accl[ui], split ? PKT_SPLIT : PKT_PLAIN, red, rb, n0, n0, resv, step_abs, skv);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- consumers: GEMV phase over Q8_0 matrices
// ggml's Q8_0 x Q8_0 product (ggml_vec_dot_q8_0_q8_0; gemv_rows_q_body in ar_kernels.cuh is the launch-per-op form): the activation row is quantised per 32-value block
// (quantize_row_q8_0: d = fp16(amax / 127), q = round(x * 127 / amax)), the block's integer dot product is scaled by d_w * d_x in fp32.  Here the integer dot products of
// 16 rows x 8 output rows x one block are ONE mma.sync.m16n8k32.s8 (every lane loads 8 consecutive bytes of its row: k slots permuted the same way on both operands).
// RMSNorm'd fp32 rows -> int8 rows + block scales in shared memory; the row stays in registers (see pk_stage_rms)
template <int NV>
__device__ __forceinline__ void pk_stage_q8_rms(const PkOp & op, const float * X, const float * snw, int R, unsigned char * sQ, float * sD, int pitchq, int nbk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, K = op.K, n4 = K >> 2;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int rr = 0; rr < 2; rr++) {
        const int r = warp + 8 * rr;
        float4 v[NV];
#pragma unroll
        for (int u = 0; u < NV; u++) { const int j = u * 32 + lane; v[u] = (j < n4 && r < R) ? __ldcg(reinterpret_cast<const float4 *>(X + (size_t) r * op.ldx) + j) : z4; }
        float sc = 1.f;
        if (op.norm == PKN_RMS) {
            double ss = 0.0;
#pragma unroll
            for (int u = 0; u < NV; u++) ss += (double) ((v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            sc = 1.0f / sqrtf((float) (ss / (double) K) + op.eps);
        }
#pragma unroll
        for (int u = 0; u < NV; u++) {
            if (u * 32 >= n4) continue;                         // (n4 is a multiple of 32: the whole warp takes the same branch)
            const int j = u * 32 + lane;
            float4 x = v[u];
            if (op.norm == PKN_RMS) { const float4 w = reinterpret_cast<const float4 *>(snw)[j]; x.x = (x.x * sc) * w.x; x.y = (x.y * sc) * w.y; x.z = (x.z * sc) * w.z; x.w = (x.w * sc) * w.w; }
            float amax = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));      // a block = the 8 float4 of lanes 8m .. 8m + 7
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
            const float id = amax != 0.f ? 127.f / amax : 0.f;
            const int a0 = __float2int_rn(x.x * id), a1 = __float2int_rn(x.y * id), a2 = __float2int_rn(x.z * id), a3 = __float2int_rn(x.w * id);
            reinterpret_cast<int *>(sQ + (size_t) r * pitchq)[j] = (a0 & 0xff) | ((a1 & 0xff) << 8) | ((a2 & 0xff) << 16) | ((a3 & 0xff) << 24);
            if ((lane & 7) == 0) sD[(size_t) r * nbk + (j >> 3)] = __half2float(__float2half_rn(amax / 127.f));
        }
    }
}
// fp16 rows written by the previous phase (attention output, SwiGLU output) -> int8 rows + block scales: a block = the 4 uint4 of lanes 4m .. 4m + 3
__device__ __forceinline__ void pk_stage_q8_h16(const PkOp & op, const __half * X16, int R, int k0, int kn, unsigned char * sQ, float * sD, int pitchq, int nbk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n8 = kn >> 3;
#pragma unroll 1
    for (int rr = 0; rr < 2; rr++) {
        const int r = warp + 8 * rr;
        const uint4 * src = reinterpret_cast<const uint4 *>(X16 + (size_t) r * op.ldx + k0);
        for (int j0 = 0; j0 < n8; j0 += 32 * 12) {
            uint4 v[12];
#pragma unroll
            for (int u = 0; u < 12; u++) { const int j = j0 + 32 * u + lane; v[u] = (r < R && j < n8) ? __ldcg(src + j) : make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll
            for (int u = 0; u < 12; u++) {
                if (j0 + 32 * u >= n8) continue;                // (n8 is a multiple of 32: warp-uniform)
                const int j = j0 + 32 * u + lane;
                float x[8];
                {
                    const unsigned wv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int i = 0; i < 4; i++) { __half2 h; memcpy(&h, &wv[i], 4); const float2 f = __half22float2(h); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
                }
                float amax = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(x[i]));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
                const float id = amax != 0.f ? 127.f / amax : 0.f;
                unsigned w[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int a0 = __float2int_rn(x[4 * h] * id), a1 = __float2int_rn(x[4 * h + 1] * id), a2 = __float2int_rn(x[4 * h + 2] * id), a3 = __float2int_rn(x[4 * h + 3] * id);
                    w[h] = (unsigned) ((a0 & 0xff) | ((a1 & 0xff) << 8) | ((a2 & 0xff) << 16) | ((a3 & 0xff) << 24));
                }
                *reinterpret_cast<uint2 *>(sQ + (size_t) r * pitchq + 8 * j) = make_uint2(w[0], w[1]);
                if ((lane & 3) == 0) sD[(size_t) r * nbk + (j >> 2)] = __half2float(__float2half_rn(amax / 127.f));
            }
        }
    }
}

// rows quantised by a PK_QUANT op -> shared memory: plain 16-byte copies (int8 values: K / 16 per row; scales: K / 128 per row)
__device__ __forceinline__ void pk_stage_q8_copy(const unsigned char * XQ, const float * XD, int K, int R, int k0, int kn, unsigned char * sQ, float * sD, int pitchq, int nbk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n16 = kn >> 4, nd4 = kn >> 7;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const int r = warp + 8 * rr;
        const uint4 * src = reinterpret_cast<const uint4 *>(XQ + (size_t) r * K + k0);
        uint4 * dst = reinterpret_cast<uint4 *>(sQ + (size_t) r * pitchq);
        uint4 v[6];
#pragma unroll
        for (int u = 0; u < 6; u++) { const int j = u * 32 + lane; v[u] = (r < R && j < n16) ? __ldcg(src + j) : z; }      // (kn <= 3 072: at most 6 per lane)
        const uint4 d = (r < R && lane < nd4) ? __ldcg(reinterpret_cast<const uint4 *>(XD + (size_t) r * (K >> 5) + (k0 >> 5)) + lane) : z;
#pragma unroll
        for (int u = 0; u < 6; u++) { const int j = u * 32 + lane; if (j < n16) dst[j] = v[u]; }
        if (lane < nd4) reinterpret_cast<uint4 *>(sD + (size_t) r * nbk)[lane] = d;
    }
}

// PK_QUANT: quantize_row_q8_0 of the phase's input rows (after ggml_rms_norm x weight when op.norm says so), once for the whole grid
__device__ __forceinline__ void pk_quant(const PkParams & P, const PkOp & op, float * red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, K = op.K, R = P.R, nrep = op.qrep ? PK_REP : 1;
    if (op.X) {                                                 // fp32 rows (the residual stream): one row per CTA turn, K <= 3 072 -> at most 3 float4 per thread
        const int n4 = K >> 2;
        double * redd = reinterpret_cast<double *>(red);
        for (int r = (int) blockIdx.x; r < R; r += (int) gridDim.x) {
            const float4 * row = reinterpret_cast<const float4 *>(op.X + (size_t) r * op.ldx);
            float4 v[3];
#pragma unroll
            for (int u = 0; u < 3; u++) { const int j = u * PK_CONS + tid; v[u] = j < n4 ? __ldcg(row + j) : make_float4(0.f, 0.f, 0.f, 0.f); }
            float sc = 1.f;
            if (op.norm == PKN_RMS) {
                double ss = 0.0;
#pragma unroll
                for (int u = 0; u < 3; u++) ss += (double) ((v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                if (lane == 0) redd[warp] = ss;
                pk_bar_sync(1, PK_CONS);
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < 8; w++) tot += redd[w];
                sc = 1.0f / sqrtf((float) (tot / (double) K) + op.eps);
                pk_bar_sync(1, PK_CONS);                        // (redd is reused by the CTA's next row)
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                if (u * PK_CONS >= n4) continue;                // (n4 is a multiple of 64 and the tail of a partial pass is handled per lane below)
                const int j = u * PK_CONS + tid;
                float4 x = v[u];
                if (op.norm == PKN_RMS && j < n4) { const float4 w = __ldg(reinterpret_cast<const float4 *>(op.nw) + j); x.x = (x.x * sc) * w.x; x.y = (x.y * sc) * w.y; x.z = (x.z * sc) * w.z; x.w = (x.w * sc) * w.w; }
                float amax = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));      // a block = the 8 float4 of lanes 8m .. 8m + 7 (whole or absent: n4 % 8 == 0)
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
                if (j >= n4) continue;
                const float id = amax != 0.f ? 127.f / amax : 0.f;
                const int a0 = __float2int_rn(x.x * id), a1 = __float2int_rn(x.y * id), a2 = __float2int_rn(x.z * id), a3 = __float2int_rn(x.w * id);
                const int word = (a0 & 0xff) | ((a1 & 0xff) << 8) | ((a2 & 0xff) << 16) | ((a3 & 0xff) << 24);
                const float d = __half2float(__float2half_rn(amax / 127.f));
                for (int c = 0; c < nrep; c++) {
                    reinterpret_cast<int *>(op.QY + op.qrep * c + (size_t) r * K)[j] = word;
                    if ((lane & 7) == 0) op.QD[op.qdrep * c + (size_t) r * (K >> 5) + (j >> 3)] = d;
                }
            }
        }
    } else {                                                    // fp16 rows written by the previous phase: item = (row, 2 048 columns), one uint4 (8 halves) per thread
        const int nch = (K + 2047) >> 11;
        for (int it = (int) blockIdx.x; it < R * nch; it += (int) gridDim.x) {
            const int r = it / nch, k0 = (it - r * nch) << 11, j = tid, col = k0 + 8 * j;
            const bool in = col < K;
            const uint4 v = in ? __ldcg(reinterpret_cast<const uint4 *>(op.X16 + (size_t) r * op.ldx + k0) + j) : make_uint4(0u, 0u, 0u, 0u);
            float x[8];
            {
                const unsigned wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) { __half2 h; memcpy(&h, &wv[i], 4); const float2 f = __half22float2(h); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
            }
            float amax = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(x[i]));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));      // a block = 4 threads (K % 32 == 0: whole or absent)
            if (!in) continue;
            const float id = amax != 0.f ? 127.f / amax : 0.f;
            unsigned w[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int a0 = __float2int_rn(x[4 * h] * id), a1 = __float2int_rn(x[4 * h + 1] * id), a2 = __float2int_rn(x[4 * h + 2] * id), a3 = __float2int_rn(x[4 * h + 3] * id);
                w[h] = (unsigned) ((a0 & 0xff) | ((a1 & 0xff) << 8) | ((a2 & 0xff) << 16) | ((a3 & 0xff) << 24));
            }
            const float d = __half2float(__float2half_rn(amax / 127.f));
            for (int c = 0; c < nrep; c++) {
                *reinterpret_cast<uint2 *>(op.QY + op.qrep * c + (size_t) r * K + col) = make_uint2(w[0], w[1]);
                if ((lane & 3) == 0) op.QD[op.qdrep * c + (size_t) r * (K >> 5) + (col >> 5)] = d;
            }
        }
    }
}

// all tiles of one unit within the staged chunk; c: the unit's sums, cl: the partner tile's (paired units).  A warp's k-slice of a tile is ks / 32 blocks (4 for a full
// tile): their scales come in one 8-byte (weights, per output column) / 16-byte (activations, per row) load each.  HI = false: at most 8 activation rows -- the
// fragments, scale products and sums of rows 8 .. 15 are skipped.
template <bool PAIR, bool HI>
__device__ __forceinline__ void pk_unit_tiles_q8(float * c, float * cl, unsigned char * ring, const unsigned char * sQ, const float * sD, int pitchq, int nbk, PkBar * full, PkBar * empty, PkRingPos & rp, int S,
                                                 int kAn, unsigned long long * pr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    for (int kt0 = 0; kt0 < kAn; kt0 += PK_TK) {
        const int ktn = kAn - kt0 < PK_TK ? kAn - kt0 : PK_TK, ks = ktn >> 3, nb = ks >> 5, tb0 = (warp * ks) >> 5, cb0 = (kt0 >> 5) + tb0;      // this warp's blocks: tb0.. within the tile, cb0.. within the chunk
        const unsigned char * xa = sQ + (size_t) g * pitchq + kt0 + warp * ks + 8 * tq, * xb = xa + (size_t) 8 * pitchq;
        const float * da = sD + (size_t) g * nbk + cb0, * db = da + (size_t) 8 * nbk;
#pragma unroll
        for (int part = 0; part < (PAIR ? 2 : 1); part++) {
            if (pr) { const unsigned long long t0 = pk_now(); pk_mbar_wait(&full[rp.s], rp.ph); pr[4] += pk_now() - t0; }
            else pk_mbar_wait(&full[rp.s], rp.ph);
            const unsigned char * st = ring + (size_t) rp.s * PK_STAGE, * wr = st + (size_t) g * PK_ROWQ + warp * ks + 8 * tq;
            const __half * ws = reinterpret_cast<const __half *>(st + PK_QSC) + tb0;
            float * acc = part ? cl : c;
            if (nb == 4) {                                      // a full tile: 4 blocks per warp, vector scale loads
                const uint2 s0 = pk_lds64(ws + (2 * tq) * 32), s1 = pk_lds64(ws + (2 * tq + 1) * 32);
                const uint4 ua = pk_lds128(da), ub = HI ? pk_lds128(db) : make_uint4(0u, 0u, 0u, 0u);
                float dw0[4], dw1[4], dxa[4], dxb[4];
                {
                    const unsigned w0[2] = {s0.x, s0.y}, w1[2] = {s1.x, s1.y}, fa[4] = {ua.x, ua.y, ua.z, ua.w}, fb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        __half2 h; float2 f;
                        memcpy(&h, &w0[i], 4); f = __half22float2(h); dw0[2 * i] = f.x; dw0[2 * i + 1] = f.y;
                        memcpy(&h, &w1[i], 4); f = __half22float2(h); dw1[2 * i] = f.x; dw1[2 * i + 1] = f.y;
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) { memcpy(&dxa[i], &fa[i], 4); memcpy(&dxb[i], &fb[i], 4); }
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint2 w = pk_lds64(wr + 32 * i), a = pk_lds64(xa + 32 * i), b = HI ? pk_lds64(xb + 32 * i) : make_uint2(0u, 0u);
                    int ci[4] = {0, 0, 0, 0};
                    pk_imma16832(ci, a.x, b.x, a.y, b.y, w.x, w.y);
                    acc[0] = fmaf((float) ci[0], dw0[i] * dxa[i], acc[0]); acc[1] = fmaf((float) ci[1], dw1[i] * dxa[i], acc[1]);
                    if (HI) { acc[2] = fmaf((float) ci[2], dw0[i] * dxb[i], acc[2]); acc[3] = fmaf((float) ci[3], dw1[i] * dxb[i], acc[3]); }
                }
            } else {
                for (int i = 0; i < nb; i++) {
                    const uint2 w = pk_lds64(wr + 32 * i), a = pk_lds64(xa + 32 * i), b = HI ? pk_lds64(xb + 32 * i) : make_uint2(0u, 0u);
                    int ci[4] = {0, 0, 0, 0};
                    pk_imma16832(ci, a.x, b.x, a.y, b.y, w.x, w.y);
                    const float dxa = da[i], dxb = HI ? db[i] : 0.f, dw0 = __half2float(ws[(2 * tq) * 32 + i]), dw1 = __half2float(ws[(2 * tq + 1) * 32 + i]);
                    acc[0] = fmaf((float) ci[0], dw0 * dxa, acc[0]); acc[1] = fmaf((float) ci[1], dw1 * dxa, acc[1]);
                    if (HI) { acc[2] = fmaf((float) ci[2], dw0 * dxb, acc[2]); acc[3] = fmaf((float) ci[3], dw1 * dxb, acc[3]); }
                }
            }
            __syncwarp();
            if (lane == 0) pk_mbar_arrive(&empty[rp.s]);
            pk_ring_next(rp, S);
        }
    }
}

template <typename KVT> __device__ __forceinline__ void pk_prefetch_kv(const PkParams & P, int layer, int step, const int * sfp, const int * spt);
template <typename KVT>
__device__ __forceinline__ void pk_unit_finish(const PkParams & P, const PkOp & op, const PkSeg & sg, float * c, float * cl, int mode, float * red, unsigned & rb, int n0, int n1, float resv, int step_abs,
                                               const unsigned long long * skv);

template <typename KVT>
__device__ __forceinline__ void pk_gemv_q8(const PkParams & P, const PkOp & op, unsigned char * ring, unsigned char * sQ, float * red, PkBar * full, PkBar * empty, PkRingPos & rp, int step_abs,
                                           unsigned long long * pr, unsigned long long * skv, const int * sfp, const int * spt) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int K = op.K, R = P.R, S = P.n_stages, ak = P.ak;
    const int nA = (K + ak - 1) / ak, kc = K < ak ? K : ak, pitchq = kc + 32, nbk = kc >> 5;
    float * sD = reinterpret_cast<float *>(sQ + (size_t) 16 * pitchq);
    unsigned long long kvoff = 0ull;
    if (op.kv_prefetch && tid < R) {
        const int pos = sfp[tid] + step_abs - P.pos_off;
        kvoff = (unsigned long long) spt[tid * P.max_pages + (pos >> 5)] * ((size_t) 2 * P.kv_heads * PK_PAGE * P.hd) + (size_t) (pos & (PK_PAGE - 1)) * P.hd;
    }
    if (op.kv_prefetch) pk_prefetch_kv<KVT>(P, op.layer, step_abs, sfp, spt);
    const size_t xoff = op.xrep * (size_t) (blockIdx.x % PK_REP);
    const float * X = op.X ? op.X + xoff : nullptr; const __half * X16 = op.X16 ? op.X16 + xoff : nullptr;
    const float * snw = nullptr; int norm_stage = -1;
    if (op.norm != PKN_NONE) {
        norm_stage = rp.s;
        pk_mbar_wait(&full[norm_stage], rp.ph);
        snw = reinterpret_cast<const float *>(ring + (size_t) norm_stage * PK_STAGE);
        pk_ring_next(rp, S);
        if (pr) pr[5] = pk_now();
    }
    auto stage = [&](int a) {
        const int kA0 = a * ak, kAn = K - kA0 < ak ? K - kA0 : ak;
        if (op.XQ) pk_stage_q8_copy(op.XQ + op.qrep * (size_t) (blockIdx.x % PK_REP), op.XD + op.qdrep * (size_t) (blockIdx.x % PK_REP), K, R, kA0, kAn, sQ, sD, pitchq, nbk);
        else if (X16) pk_stage_q8_h16(op, X16, R, kA0, kAn, sQ, sD, pitchq, nbk);
        else if (K <= 1024) pk_stage_q8_rms<8>(op, X, snw, R, sQ, sD, pitchq, nbk);      // (fp32 rows are staged whole: host-checked K <= ak <= 3 072)
        else pk_stage_q8_rms<24>(op, X, snw, R, sQ, sD, pitchq, nbk);
        if (a == 0 && op.kv_prefetch && tid < R) skv[tid] = kvoff;
        if (norm_stage >= 0 && a + 1 == nA) { __syncwarp(); if (lane == 0) pk_mbar_arrive(&empty[norm_stage]); }
        pk_bar_sync(1, PK_CONS);
        if (pr && a == 0) pr[1] = pk_now();
        return kAn;
    };
    auto seg_of = [&](int u) -> const PkSeg & { return op.seg[(op.nseg > 2 && u >= op.seg[2].unit0) ? 2 : ((op.nseg > 1 && u >= op.seg[1].unit0) ? 1 : 0)]; };
    auto res_of = [&](const PkSeg & sg, int n0) -> float {
        if (sg.epi != PKE_RES || tid >= 128) return 0.f;
        const int r = tid >> 3, n = n0 + (tid & 7);
        return (r < R && n < sg.N) ? __ldcg(sg.res + sg.yrep * (size_t) (blockIdx.x % PK_REP) + (size_t) r * sg.ldy + n) : 0.f;
    };
    unsigned rb = 0;
    const bool hi = R > 8;
    if (nA == 1) {
        const int kAn = stage(0);
        for (int u = (int) blockIdx.x; u < op.n_units; u += (int) gridDim.x) {
            const PkSeg & sg = seg_of(u);
            int n0, n1;
            pk_unit_rows(sg, u, P.hd, n0, n1);
            const float resv = res_of(sg, n0);
            float c[4] = {0.f, 0.f, 0.f, 0.f}, cl[4] = {0.f, 0.f, 0.f, 0.f};
            if (hi) {
                if (sg.pair != PKP_NONE) pk_unit_tiles_q8<true, true>(c, cl, ring, sQ, sD, pitchq, nbk, full, empty, rp, S, kAn, pr);
                else pk_unit_tiles_q8<false, true>(c, cl, ring, sQ, sD, pitchq, nbk, full, empty, rp, S, kAn, pr);
            } else {
                if (sg.pair != PKP_NONE) pk_unit_tiles_q8<true, false>(c, cl, ring, sQ, sD, pitchq, nbk, full, empty, rp, S, kAn, pr);
                else pk_unit_tiles_q8<false, false>(c, cl, ring, sQ, sD, pitchq, nbk, full, empty, rp, S, kAn, pr);
            }
            pk_unit_finish<KVT>(P, op, sg, c, cl, sg.pair != PKP_NONE ? PKT_PAIR : PKT_PLAIN, red, rb, n0, n1, resv, step_abs, skv);
        }
    } else {
        float acc[3][4], zero[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 3; i++) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        for (int a = 0; a < nA; a++) {
            if (a) pk_bar_sync(1, PK_CONS);
            const int kAn = stage(a);
#pragma unroll
            for (int ui = 0; ui < 3; ui++) {
                const int u = (int) blockIdx.x + ui * (int) gridDim.x;
                if (u < op.n_units) {
                    const PkSeg & sg = seg_of(u);
                    const int n0 = (u - sg.unit0) * 8;
                    const float resv = a + 1 == nA ? res_of(sg, n0) : 0.f;
                    if (hi) pk_unit_tiles_q8<false, true>(acc[ui], zero, ring, sQ, sD, pitchq, nbk, full, empty, rp, S, kAn, pr);
                    else pk_unit_tiles_q8<false, false>(acc[ui], zero, ring, sQ, sD, pitchq, nbk, full, empty, rp, S, kAn, pr);
                    if (a + 1 == nA) pk_unit_finish<KVT>(P, op, sg, acc[ui], zero, PKT_PLAIN, red, rb, n0, n0, resv, step_abs, skv);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- consumers: attention phase.  One (row, head) item per half-CTA (128 threads, named barrier 2 + grp)
// raw 8-element cache reads: issued in batches so that a thread has 8 independent 16-byte (fp32 store: 2 x 16-byte) loads in flight, converted on use
struct PkRawH { uint4 a; };
struct PkRawF { float4 a, b; };
__device__ __forceinline__ void pk_raw_load(const __half * p, PkRawH & r) { r.a = __ldcg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void pk_raw_load(const float * p, PkRawF & r) { r.a = __ldcg(reinterpret_cast<const float4 *>(p)); r.b = __ldcg(reinterpret_cast<const float4 *>(p) + 1); }
__device__ __forceinline__ void pk_raw_zero(PkRawH & r) { r.a = make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ void pk_raw_zero(PkRawF & r) { r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; }
__device__ __forceinline__ void pk_raw_f(const PkRawH & r, float * v) {
    const unsigned w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int i = 0; i < 4; i++) { __half2 h; memcpy(&h, &w[i], 4); const float2 f = __half22float2(h); v[2 * i] = f.x; v[2 * i + 1] = f.y; }      // (memcpy: no type-punned reads)
}
__device__ __forceinline__ void pk_raw_f(const PkRawF & r, float * v) { v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w; }
template <typename CT> struct PkRawOf { typedef PkRawF type; };
template <> struct PkRawOf<__half> { typedef PkRawH type; };

#ifdef B2EMU
static inline void pk_prefetch_l2(const void *) {}
#else
__device__ __forceinline__ void pk_prefetch_l2(const void * p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

template <typename CT> struct PkAttU { static constexpr int v = sizeof(CT) == 2 ? 16 : 8; };      // keys per thread in flight (16 bytes each for fp16 pages, 32 for fp32 stores)
constexpr int PK_ATT_HDR = 1024;                               // floats of per-group scratch before the P.V partials: q [128] | reduction [128] | (pad) | page offsets [256 x 8 bytes] at 512
static inline size_t pk_att_bytes(int T) { return (size_t) (PK_ATT_HDR + 1024) * 4 + (size_t) ((T + 3) & ~3) * 4; }      // per half-CTA group

// One item = one (row, query head); kh = the kv head it reads (the reference's repeat-interleaved GQA cache, orpheus model.cpp:196-228, by indexing).
// HD = head size (compile time: the per-key reduction over the HD / 8 threads of a key is three unrolled shuffles that the scheduler interleaves across the keys in
// flight; with a run-time head size it was a serial loop per key -- 20 % of the kernel's issue slots on a B200).
// (Tried: the first batch of K rows issued before q arrives and the first batch of V rows before the softmax reductions, to overlap the item's chain of L2 round trips:
// the load registers then live across the block barriers and the item became 1.5-2x slower -- r2h / r2k2 timelines.)
// (Tried: one item per kv head with its 3 query heads sharing the K / V loads -- 2.6x slower per item, the scores / P.V arithmetic of three heads per thread is not
// free at 128 threads; r2g timeline.)
template <typename KVT, typename CT, int HD>
__device__ __forceinline__ void pk_attn_item(const PkParams & P, const PkOp & op, float * base, int grp, int r, int h, int kh, int t0, int T, float * chunk, const int * spt) {      // positions [t0, T); chunk != null: one position chunk of a split item
    typedef typename PkRawOf<CT>::type Raw;
    constexpr int U = PkAttU<CT>::v, PARTS = HD / 8, KPP = 128 / PARTS;
    const int gt = threadIdx.x & 127, gw = gt >> 5, H = P.H, part = gt % PARTS, kq = gt / PARTS;
    float * qs = base; float * wredf = base + 128; double * wredd = reinterpret_cast<double *>(base + 136);
    unsigned long long * spo = reinterpret_cast<unsigned long long *>(base + 512);      // element offset of each of this sequence's pages within the layer's pool
    float * pvs = base + PK_ATT_HDR; float * sc = base + PK_ATT_HDR + 1024;
    const size_t page_elems = (size_t) 2 * P.kv_heads * PK_PAGE * HD;
    const CT * flat_k = reinterpret_cast<const CT *>(op.ck) + op.cross_row_stride * (size_t) r + (size_t) kh * HD + part * 8;
    const CT * flat_v = reinterpret_cast<const CT *>(op.cv) + op.cross_row_stride * (size_t) r + (size_t) kh * HD + part * 8;
    const CT * pool_k = reinterpret_cast<const CT *>(P.kv_pool + (size_t) op.layer * P.kv_layer_bytes) + (size_t) kh * PK_PAGE * HD + part * 8;
    const CT * pool_v = pool_k + (size_t) P.kv_heads * PK_PAGE * HD;
    const int * pt = spt + r * P.max_pages;
    auto krow = [&](int t, int kv) -> const CT * {
        if (op.cross) return (kv ? flat_v : flat_k) + (size_t) t * H;
        return (kv ? pool_v : pool_k) + spo[t >> 5] + (t & (PK_PAGE - 1)) * HD;
    };
    if (gt < HD) qs[gt] = __ldcg(op.q + (size_t) r * H + (size_t) h * HD + gt);
    if (!op.cross) for (int i = gt + (t0 >> 5); i * PK_PAGE < T; i += 128) spo[i] = (unsigned long long) pt[i] * page_elems;
    pk_bar_sync(2 + grp, 128);
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q8[i] = qs[part * 8 + i];
    // scores: PARTS threads per key (8 channels each), KPP keys per pass, U passes in flight
    float mloc = -INFINITY;
    for (int tb = t0; tb < T; tb += KPP * U) {
        Raw raw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int t = tb + u * KPP + kq;
            if (t < T) pk_raw_load(krow(t, 0), raw[u]); else pk_raw_zero(raw[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int t = tb + u * KPP + kq;
            float k8[8];
            pk_raw_f(raw[u], k8);
            float a = fmaf(q8[3], k8[3], fmaf(q8[2], k8[2], fmaf(q8[1], k8[1], q8[0] * k8[0]))) + fmaf(q8[7], k8[7], fmaf(q8[6], k8[6], fmaf(q8[5], k8[5], q8[4] * k8[4])));
#pragma unroll
            for (int o = PARTS >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            a *= op.scale;
            if (t < T) { if (part == 0) sc[t - t0] = a; mloc = fmaxf(mloc, a); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, o));
    if ((gt & 31) == 0) wredf[gw] = mloc;
    pk_bar_sync(2 + grp, 128);
    const float m = fmaxf(fmaxf(wredf[0], wredf[1]), fmaxf(wredf[2], wredf[3]));
    double sum = 0.0;                                           // ggml_soft_max: expf(s - max), the sum accumulated in double, scale by (float) (1 / sum)
    for (int t = gt; t < T - t0; t += 128) { const float e = expf(sc[t] - m); sc[t] = e; sum += (double) e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((gt & 31) == 0) wredd[gw] = sum;
    pk_bar_sync(2 + grp, 128);
    const double total = ((wredd[0] + wredd[1]) + wredd[2]) + wredd[3];
    const float inv = chunk ? 1.0f : (float) (1.0 / total);      // a chunk keeps its exps unnormalised: pk_attn_combine divides by the sum over all chunks
    // P.V: thread (slice kq, part) walks positions kq, kq + KPP, ... for its 8 channels with p = e * inv; slices summed in order afterwards
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    for (int tb = t0; tb < T; tb += KPP * U) {
        Raw raw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int t = tb + u * KPP + kq;
            if (t < T) pk_raw_load(krow(t, 1), raw[u]); else pk_raw_zero(raw[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int t = tb + u * KPP + kq;
            const float p = t < T ? sc[t - t0] * inv : 0.f;
            float v8[8];
            pk_raw_f(raw[u], v8);
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = fmaf(p, v8[i], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) pvs[(size_t) kq * HD + part * 8 + i] = acc[i];
    pk_bar_sync(2 + grp, 128);
    if (gt < HD) {
        float a = 0.f;
#pragma unroll 8
        for (int sl = 0; sl < KPP; sl++) a += pvs[(size_t) sl * HD + gt];
        if (chunk) {
            chunk[4 + gt] = a;
            if (gt == 0) { chunk[0] = m; chunk[1] = (float) total; }
        } else {
            const __half hv = __float2half_rn(a);
            for (int c = 0; c < (op.orep ? PK_REP : 1); c++) op.out16[op.orep * c + (size_t) r * H + (size_t) h * HD + gt] = hv;
        }
    }
    pk_bar_sync(2 + grp, 128);                                  // the scratch is free for the group's next item
}

// Cross-attention over a SHORT flat store (T <= 4 x 128 / (HD / 8) keys: Parler's text encoding): the same arithmetic in the same order as pk_attn_item, but the K and the
// V rows (at most 4 + 4 per thread) are requested together with q, so the item is one L2 round trip plus arithmetic instead of three dependent ones.
template <int HD>
__device__ __forceinline__ void pk_attn_cross_short(const PkParams & P, const PkOp & op, float * base, int grp, int r, int h, int T) {
    constexpr int US = 4, PARTS = HD / 8, KPP = 128 / PARTS;
    const int gt = threadIdx.x & 127, gw = gt >> 5, H = P.H, part = gt % PARTS, kq = gt / PARTS;
    float * qs = base; float * wredf = base + 128; double * wredd = reinterpret_cast<double *>(base + 136);
    float * pvs = base + PK_ATT_HDR; float * sc = base + PK_ATT_HDR + 1024;
    const float * flat_k = op.ck + op.cross_row_stride * (size_t) r + (size_t) h * HD + part * 8, * flat_v = op.cv + op.cross_row_stride * (size_t) r + (size_t) h * HD + part * 8;
    PkRawF rk[US], rv[US];
#pragma unroll
    for (int u = 0; u < US; u++) {
        const int t = u * KPP + kq;
        if (t < T) { pk_raw_load(flat_k + (size_t) t * H, rk[u]); pk_raw_load(flat_v + (size_t) t * H, rv[u]); } else { pk_raw_zero(rk[u]); pk_raw_zero(rv[u]); }
    }
    if (gt < HD) qs[gt] = __ldcg(op.q + (size_t) r * H + (size_t) h * HD + gt);
    pk_bar_sync(2 + grp, 128);
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q8[i] = qs[part * 8 + i];
    float mloc = -INFINITY;
#pragma unroll
    for (int u = 0; u < US; u++) {
        const int t = u * KPP + kq;
        float k8[8];
        pk_raw_f(rk[u], k8);
        float a = fmaf(q8[3], k8[3], fmaf(q8[2], k8[2], fmaf(q8[1], k8[1], q8[0] * k8[0]))) + fmaf(q8[7], k8[7], fmaf(q8[6], k8[6], fmaf(q8[5], k8[5], q8[4] * k8[4])));
#pragma unroll
        for (int o = PARTS >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        a *= op.scale;
        if (t < T) { if (part == 0) sc[t] = a; mloc = fmaxf(mloc, a); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, o));
    if ((gt & 31) == 0) wredf[gw] = mloc;
    pk_bar_sync(2 + grp, 128);
    const float m = fmaxf(fmaxf(wredf[0], wredf[1]), fmaxf(wredf[2], wredf[3]));
    double sum = 0.0;
    for (int t = gt; t < T; t += 128) { const float e = expf(sc[t] - m); sc[t] = e; sum += (double) e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((gt & 31) == 0) wredd[gw] = sum;
    pk_bar_sync(2 + grp, 128);
    const float inv = (float) (1.0 / (((wredd[0] + wredd[1]) + wredd[2]) + wredd[3]));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
#pragma unroll
    for (int u = 0; u < US; u++) {
        const int t = u * KPP + kq;
        const float p = t < T ? sc[t] * inv : 0.f;
        float v8[8];
        pk_raw_f(rv[u], v8);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fmaf(p, v8[i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) pvs[(size_t) kq * HD + part * 8 + i] = acc[i];
    pk_bar_sync(2 + grp, 128);
    if (gt < HD) {
        float a = 0.f;
#pragma unroll 8
        for (int sl = 0; sl < KPP; sl++) a += pvs[(size_t) sl * HD + gt];
        const __half hv = __float2half_rn(a);
        for (int c = 0; c < (op.orep ? PK_REP : 1); c++) op.out16[op.orep * c + (size_t) r * H + (size_t) h * HD + gt] = hv;
    }
    pk_bar_sync(2 + grp, 128);
}

template <typename KVT, int HD>
__device__ __forceinline__ void pk_attn(const PkParams & P, const PkOp & op, unsigned char * scratch, int step, const int * sfp, const int * spt) {
    const int grp = threadIdx.x >> 7;
    float * base = reinterpret_cast<float *>(scratch + (size_t) grp * (P.a_bytes / 2));
    const int rep = P.heads / P.kv_heads, ns = op.tsplit > 1 ? op.tsplit : 1;
    for (int it = (int) blockIdx.x * 2 + grp; it < P.R * P.heads * ns; it += 2 * (int) gridDim.x) {
        const int item = it / ns, c = it - item * ns, r = item / P.heads, h = item - r * P.heads;
        const int T = op.cross ? op.cross_len : sfp[r] + step - P.pos_off + 1;
        int t0 = 0, t1 = T;
        float * part = nullptr;
        if (ns > 1) {                                           // chunk c of the item's positions, in whole pages
            const int per = (((T + ns - 1) / ns) + PK_PAGE - 1) & ~(PK_PAGE - 1);
            t0 = c * per; t1 = t0 + per < T ? t0 + per : T;
            part = P.att_part + ((size_t) item * ns + c) * (HD + 4);
            if (t0 >= T) { if ((threadIdx.x & 127) == 0) { part[0] = -INFINITY; part[1] = 0.f; } continue; }      // an empty chunk (short context): weight 0 in the combination
        }
        if (op.cross && ns == 1 && T <= 4 * (128 / (HD / 8))) pk_attn_cross_short<HD>(P, op, base, grp, r, h, T);
        else if (op.cross) pk_attn_item<KVT, float, HD>(P, op, base, grp, r, h, h, t0, t1, part, spt);
        else pk_attn_item<KVT, KVT, HD>(P, op, base, grp, r, h, h / rep, t0, t1, part, spt);
    }
}

// the chunks of a split attention phase combined: out = sum_c w_c acc_c / sum_c w_c sum_c with w_c = exp(max_c - max) -- the reference's softmax(max, expf, double sum)
// followed by P.V, evaluated chunk by chunk
template <int HD>
__device__ __forceinline__ void pk_attn_combine(const PkParams & P, const PkOp & op) {
    const int grp = threadIdx.x >> 7, gt = threadIdx.x & 127, ns = op.tsplit, H = P.H;
    for (int item = (int) blockIdx.x * 2 + grp; item < P.R * P.heads; item += 2 * (int) gridDim.x) {
        const int r = item / P.heads, h = item - r * P.heads;
        const float * part = P.att_part + (size_t) item * ns * (HD + 4);
        float M = -INFINITY;
        for (int c = 0; c < ns; c++) M = fmaxf(M, __ldcg(part + (size_t) c * (HD + 4)));
        double den = 0.0; float a = 0.f;
        for (int c = 0; c < ns; c++) {
            const float * pc = part + (size_t) c * (HD + 4);
            const float mc = __ldcg(pc), w = mc == -INFINITY ? 0.f : expf(mc - M);
            den += (double) (w * __ldcg(pc + 1));
            if (gt < HD && w != 0.f) a = fmaf(w, __ldcg(pc + 4 + gt), a);
        }
        if (gt < HD) {
            const __half hv = __float2half_rn(a * (float) (1.0 / den));
            for (int cp = 0; cp < (op.orep ? PK_REP : 1); cp++) op.out16[op.orep * cp + (size_t) r * H + (size_t) h * HD + gt] = hv;
        }
    }
}

// ask L2 for the K / V rows this CTA's self-attention items of layer `layer` will read (all positions but the one this step appends): issued at the start of the layer's
// q|k|v phase, so that HBM serves them while that phase runs and the attention phase finds them in L2
template <typename KVT>
__device__ __forceinline__ void pk_prefetch_kv(const PkParams & P, int layer, int step, const int * sfp, const int * spt) {
    const int grp = threadIdx.x >> 7, gt = threadIdx.x & 127, hd = P.hd;
    const size_t page_elems = (size_t) 2 * P.kv_heads * PK_PAGE * hd, v_off = (size_t) P.kv_heads * PK_PAGE * hd;
    const int lines = (hd * (int) sizeof(KVT) + 127) / 128;     // 128-byte lines per cache row
    for (int it = (int) blockIdx.x * 2 + grp; it < P.R * P.kv_heads; it += 2 * (int) gridDim.x) {
        const int r = it / P.kv_heads, h = it - r * P.kv_heads, T = sfp[r] + step - P.pos_off;
        const KVT * pool = reinterpret_cast<const KVT *>(P.kv_pool + (size_t) layer * P.kv_layer_bytes) + (size_t) h * PK_PAGE * hd;
        for (int i = gt; i < T * lines; i += 128) {
            const int t = i / lines, ln = i - t * lines;
            const KVT * row = pool + (size_t) spt[r * P.max_pages + (t >> 5)] * page_elems + (size_t) (t & (PK_PAGE - 1)) * hd + ln * (128 / (int) sizeof(KVT));
            pk_prefetch_l2(row); pk_prefetch_l2(row + v_off);
        }
    }
}

// ---------------------------------------------------------------- consumers: rows + embedding, argmax
__device__ __forceinline__ void pk_rows(const PkParams & P, int step, int * sids) {
    const int tid = threadIdx.x, n_out = P.n_out, R = P.R, H = P.H;
    for (int b = (int) blockIdx.x; b < R; b += (int) gridDim.x) {
        const int pos = (P.first_pos ? P.first_pos[b] : 0) + step;
        if (tid == 0) {                                        // delay_rows_kernel for sequence b
            const int * last = (P.d_teacher ? P.d_teacher : P.d_out) + ((size_t) (step > 0 ? step - 1 : 0) * R + b) * n_out;
            if (P.seen && step >= 1 && P.stopped[b] < 0) {
                bool stop = pos >= P.max_gen;
                if (!stop) { stop = true; for (int i = 0; i < n_out; i++) stop = stop && (P.seen[b * n_out + i] || __ldcg(last + i) == P.eos); }
                if (stop) P.stopped[b] = step;
            }
            for (int i = 0; i < n_out; i++) {
                const bool s = P.seen && P.seen[b * n_out + i];
                const int lt = step > 0 ? __ldcg(last + i) : 0;
                const int id = step > i ? (s ? P.eos : lt) : P.bos;
                sids[i] = id; P.ids[b * n_out + i] = id;
                if (P.seen && step >= 1 && lt == P.eos) P.seen[b * n_out + i] = 1;
            }
            P.row_pos[b] = pos;
        }
        pk_bar_sync(1, PK_CONS);
        for (int c = tid; c < H; c += PK_CONS) {              // codebook_embed_kernel: the tables' rows summed in head order, then the positional row
            float a = P.tables[(size_t) sids[0] * H + c];
            for (int i = 1; i < n_out; i++) a = P.tables[(size_t) i * P.tab_stride + (size_t) sids[i] * H + c] + a;
            if (P.pos_embed) a = a + P.pos_embed[(size_t) pos * H + c];
            for (int cp = 0; cp < (P.x0rep ? PK_REP : 1); cp++) P.x0[P.x0rep * cp + (size_t) b * H + c] = a;
        }
        pk_bar_sync(1, PK_CONS);
    }
}

__device__ __forceinline__ void pk_argmax(const PkParams & P, int step, float * red) {      // sampler::max per (row, head): the first maximum wins
    float * sv = red; int * si = reinterpret_cast<int *>(red + 256);
    const int tid = threadIdx.x, V = P.vocab, rows = P.R * P.n_out;
    for (int b = (int) blockIdx.x; b < rows; b += (int) gridDim.x) {
        const float * lg = P.logits + (size_t) b * V;
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int i = tid; i < V; i += PK_CONS) { const float v = __ldcg(lg + i); if (v > best) { best = v; bi = i; } }
        sv[tid] = best; si[tid] = bi;
        pk_bar_sync(1, PK_CONS);
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) { if (sv[tid + o] > sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; } }
            pk_bar_sync(1, PK_CONS);
        }
        if (tid == 0) P.d_out[(size_t) step * rows + b] = si[0] == 0x7fffffff ? 0 : si[0];
        pk_bar_sync(1, PK_CONS);
    }
}

// ---- PKM_ORPHEUS (llama-3 style: orpheus_runner::build_orpheus_graph, reference src/models/orpheus/model.cpp:231-353)
// the token row b produced at step s: its amax_ch partial maxima combined (first maximum wins: sampler::max), written to d_out, checked against the stopping token
// (generate_from_batch's stop rule, model.cpp:389-398); every consumer thread returns it
__device__ __forceinline__ int pk_orpheus_token(const PkParams & P, int s, int b, float * red) {
    float * sv = red; int * si = reinterpret_cast<int *>(red + 256);
    const int tid = threadIdx.x;
    sv[tid] = tid < P.amax_ch ? __ldcg(P.amax_v + (size_t) b * P.amax_ch + tid) : -INFINITY;
    si[tid] = tid < P.amax_ch ? __ldcg(P.amax_i + (size_t) b * P.amax_ch + tid) : 0x7fffffff;
    pk_bar_sync(1, PK_CONS);
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { if (sv[tid + o] > sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; } }
        pk_bar_sync(1, PK_CONS);
    }
    const int tok = si[0] == 0x7fffffff ? 0 : si[0];
    pk_bar_sync(1, PK_CONS);                                   // (red is reused right away)
    if (tid == 0) {
        P.d_out[(size_t) b * P.n_steps_total + s] = tok;
        if (P.stopped && P.stopped[b] < 0 && tok == P.stop_token) P.stopped[b] = s + 1;
    }
    return tok;
}

// rows of a decode step: x0[b] = embed[token of step - 1] (ggml_get_rows), position = prompt length + step - 1, and the step's NeoX RoPE table for that position
// (ggml rope cache: theta starts at the position and is multiplied by theta_scale pair after pair; llama-3 frequency factors divide it)
__device__ __forceinline__ void pk_rows_orpheus(const PkParams & P, int step, bool first_in_launch, float * red) {
    const int tid = threadIdx.x, R = P.R, H = P.H, half = P.hd >> 1;
    for (int b = (int) blockIdx.x; b < R; b += (int) gridDim.x) {
        const int tok = first_in_launch ? __ldcg(P.d_out + (size_t) b * P.n_steps_total + step - 1) : pk_orpheus_token(P, step - 1, b, red);
        const int pos = P.first_pos[b] + step - P.pos_off;
        const float4 * src = reinterpret_cast<const float4 *>(P.embed + (size_t) tok * H);
        for (int c = tid; c < (H >> 2); c += PK_CONS) {
            const float4 v = __ldg(src + c);
            for (int cp = 0; cp < (P.x0rep ? PK_REP : 1); cp++) reinterpret_cast<float4 *>(P.x0 + P.x0rep * cp + (size_t) b * H)[c] = v;
        }
        for (int i = tid; i < half; i += PK_CONS) {
            float theta = (float) pos;
            for (int j = 0; j < i; j++) theta *= P.theta_scale;
            const float th = P.rope_ff ? theta / P.rope_ff[i] : theta;
            P.rope_cs[(size_t) b * half + i] = make_float2(cosf(th), sinf(th));
        }
        if (tid == 0) P.row_pos[b] = pos;
    }
}

// partial maxima of the step's logits: item (row b, chunk c) scans [c * len, (c + 1) * len) of the row
__device__ __forceinline__ void pk_argmax_partial(const PkParams & P, float * red) {
    float * sv = red; int * si = reinterpret_cast<int *>(red + 256);
    const int tid = threadIdx.x, V = P.vocab, nch = P.amax_ch, len = (V + nch - 1) / nch;
    for (int item = (int) blockIdx.x; item < P.R * nch; item += (int) gridDim.x) {
        const int b = item / nch, c = item - b * nch, lo = c * len, hi = lo + len < V ? lo + len : V;
        const float * lg = P.logits + (size_t) b * V;
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int i0 = lo + tid; i0 < hi; i0 += PK_CONS * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = i0 + u * PK_CONS < hi ? __ldcg(lg + i0 + u * PK_CONS) : -INFINITY;
#pragma unroll
            for (int u = 0; u < 8; u++) if (v[u] > best) { best = v[u]; bi = i0 + u * PK_CONS; }      // ascending indices per thread: the first maximum stays
        }
        sv[tid] = best; si[tid] = bi;
        pk_bar_sync(1, PK_CONS);
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) { if (sv[tid + o] > sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; } }
            pk_bar_sync(1, PK_CONS);
        }
        if (tid == 0) { P.amax_v[item] = sv[0]; P.amax_i[item] = si[0]; }
        pk_bar_sync(1, PK_CONS);
    }
}

// ---- PKM_DIA (dia_runner::generate_from_batch / check_stopping, reference src/models/dia/model.cpp:806-858; build_dia_decoder :580-700)
// rows of a step: utterance u's ids under the delay pattern with the end-of-stream injection (dia_step_rows_kernel), the summed codebook embeddings for its two rows
// (codebook_embed_kernel) and the NeoX RoPE table of position `step` (no frequency factors)
__device__ __forceinline__ void pk_rows_dia(const PkParams & P, int step, int * sids) {
    const int tid = threadIdx.x, n_out = P.n_out, Bu = P.R >> 1, H = P.H, half = P.hd >> 1;
    for (int u = (int) blockIdx.x; u < Bu; u += (int) gridDim.x) {
        if (tid == 0) {
            const int * last = (P.d_teacher ? P.d_teacher : P.d_out) + ((size_t) (step > 0 ? step - 1 : 0) * Bu + u) * n_out;
            const int pattern[9] = {0, 8, 9, 10, 11, 12, 13, 14, 15};     // dia_model::delay_pattern (model.h:85)
            int audio[9];
            for (int i = 0; i < n_out; i++) audio[i] = step > i ? __ldcg(last + i) : P.bos;
            int d = P.delay[u];
            if (d == -1 && (audio[0] == P.eos || step >= P.max_gen - P.max_delay)) d = P.max_delay;
            if (d > 0) {
                const int after = P.max_delay - d;
                for (int i = 0; i < n_out; i++) {
                    if (after == pattern[i]) audio[i] = P.eos;
                    else if (after > pattern[i]) audio[i] = P.pad;
                }
                d -= 1;
            }
            P.delay[u] = d;
            if (d == 0 && P.stopped[u] < 0) P.stopped[u] = step;
            for (int i = 0; i < n_out; i++) { sids[i] = audio[i]; P.ids[(2 * u) * n_out + i] = audio[i]; P.ids[(2 * u + 1) * n_out + i] = audio[i]; }
            P.row_pos[2 * u] = step; P.row_pos[2 * u + 1] = step;
        }
        pk_bar_sync(1, PK_CONS);
        for (int c = tid; c < H; c += PK_CONS) {              // the tables' rows summed in head order
            float a = P.tables[(size_t) sids[0] * H + c];
            for (int i = 1; i < n_out; i++) a = P.tables[(size_t) i * P.tab_stride + (size_t) sids[i] * H + c] + a;
            for (int cp = 0; cp < (P.x0rep ? PK_REP : 1); cp++) { P.x0[P.x0rep * cp + (size_t) (2 * u) * H + c] = a; P.x0[P.x0rep * cp + (size_t) (2 * u + 1) * H + c] = a; }
        }
        for (int i = tid; i < half; i += PK_CONS) {
            float theta = (float) step;
            for (int j = 0; j < i; j++) theta *= P.theta_scale;
            const float2 cs = make_float2(cosf(theta), sinf(theta));
            P.rope_cs[(size_t) (2 * u) * half + i] = cs; P.rope_cs[(size_t) (2 * u + 1) * half + i] = cs;
        }
        pk_bar_sync(1, PK_CONS);
    }
}

// cfg_scale then sampler::max per (utterance, head): out = cond + scale * (cond - uncond); the first maximum wins
__device__ __forceinline__ void pk_argmax_dia(const PkParams & P, int step, float * red) {
    float * sv = red; int * si = reinterpret_cast<int *>(red + 256);
    const int tid = threadIdx.x, V = P.vocab, n_out = P.n_out, NV = n_out * V, Bu = P.R >> 1;
    for (int b = (int) blockIdx.x; b < Bu * n_out; b += (int) gridDim.x) {
        const int u = b / n_out, i = b - u * n_out;
        const float * lc = P.logits + (size_t) (2 * u) * NV + (size_t) i * V, * lu = lc + NV;
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int j = tid; j < V; j += PK_CONS) {
            const float cr = __ldcg(lc + j), ur = __ldcg(lu + j), v = cr + P.cfg * (cr - ur);
            P.logits_cfg[(size_t) u * NV + (size_t) i * V + j] = v;
            if (P.logits_all) P.logits_all[((size_t) step * Bu + u) * NV + (size_t) i * V + j] = v;
            if (v > best) { best = v; bi = j; }
        }
        sv[tid] = best; si[tid] = bi;
        pk_bar_sync(1, PK_CONS);
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) { if (sv[tid + o] > sv[tid] || (sv[tid + o] == sv[tid] && si[tid + o] < si[tid])) { sv[tid] = sv[tid + o]; si[tid] = si[tid + o]; } }
            pk_bar_sync(1, PK_CONS);
        }
        if (tid == 0) P.d_out[(size_t) step * Bu * n_out + b] = si[0] == 0x7fffffff ? 0 : si[0];
        pk_bar_sync(1, PK_CONS);
    }
}

// ---------------------------------------------------------------- the kernel
template <typename KVT, int HD>
__global__ void __launch_bounds__(PK_THREADS, 1) pdk_kernel(const PkParams P) {
    extern __shared__ __align__(128) unsigned char pk_smem[];
    unsigned char * ring = pk_smem;
    unsigned char * areg = pk_smem + (size_t) P.n_stages * PK_STAGE;           // activation rows (GEMV phases) / attention scratch
    float * red = reinterpret_cast<float *>(areg + P.a_bytes);
    int * sids = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(red) + PK_RED_BYTES);      // [16] ids of the row being embedded
    PkBar * full = reinterpret_cast<PkBar *>(sids + 16);
    PkBar * empty = full + PK_MAXSTAGES;
    PkOp * sops = reinterpret_cast<PkOp *>((reinterpret_cast<uintptr_t>(empty + PK_MAXSTAGES) + 15) & ~(uintptr_t) 15);      // [2]: the running op's descriptor and the next one's
    unsigned long long * skv = reinterpret_cast<unsigned long long *>(sops + 3);      // (sops[2] is the producer warp's copy)                                             // [16]: cache slots of the rows' new k / v
    // first decode position of every sequence and the page table: constant for the launch, read by every CTA in every attention phase -- from shared memory (as
    // global reads they were 148 requesters on a handful of L2 lines right after each barrier, and ld.acquire's L1 invalidation defeats caching them)
    int * sfp = reinterpret_cast<int *>(skv + 16); int * spt = sfp + 16;
    for (int i = (int) threadIdx.x; i < 16; i += PK_THREADS) sfp[i] = (i < P.R && P.first_pos) ? P.first_pos[i] : 0;
    for (int i = (int) threadIdx.x; i < P.R * P.max_pages; i += PK_THREADS) spt[i] = P.page_table[i];
    const int tid = threadIdx.x, warp = tid >> 5, S = P.n_stages;
    if (tid == 0) {
        for (int s = 0; s < S; s++) { pk_mbar_init(&full[s], 1); pk_mbar_init(&empty[s], 8); }
        pk_fence_init();
    }
    __syncthreads();
    PkRingPos rp; rp.s = 0; rp.ph = 0u;
    if (warp == 8) {                                           // ---- producer warp: streams this CTA's weight tiles in program order, as far ahead as the ring allows
        // the descriptor of the op being produced is copied to shared memory first (17 independent 16-byte loads: one L2 round trip per op instead of one per field
        // and tile -- the mbarrier / bulk-copy instructions are compiler barriers, so every field read from global memory was re-issued after each of them)
        PkOp * pop = sops + 2;
        const int lane = tid & 31;
        for (int st = 0; st < P.n_steps; st++)
            for (int oi = 0; oi < P.n_ops; oi++) {
                __syncwarp();
                for (int w = lane; w < (int) (sizeof(PkOp) / 16); w += 32) reinterpret_cast<uint4 *>(pop)[w] = __ldg(reinterpret_cast<const uint4 *>(&P.ops[oi]) + w);
                __syncwarp();
                if (pop->kind == PK_GEMV) pk_produce_gemv(*pop, ring, full, empty, S, rp, P.ak, P.hd);
            }
        return;
    }
    unsigned epoch = 0;
    // op descriptors are read from shared memory: the next op's descriptor is requested from global memory when an op starts and parked in the other slot before the
    // grid barrier (a descriptor read at op start would put one more L2 round trip on every phase's critical path)
    constexpr int OPW = (int) (sizeof(PkOp) / 16);
    if (tid < OPW) reinterpret_cast<uint4 *>(&sops[0])[tid] = __ldg(reinterpret_cast<const uint4 *>(&P.ops[0]) + tid);
    pk_bar_sync(1, PK_CONS);
    unsigned opn = 0;
    for (int st = 0; st < P.n_steps; st++) {
        const int step = P.step_begin + st;
        const bool prof = P.prof && step == P.prof_step && tid == 0;
        for (int oi = 0; oi < P.n_ops; oi++, opn++) {
            const PkOp & op = sops[opn & 1u];
            uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
            if (tid < OPW) nxt = __ldg(reinterpret_cast<const uint4 *>(&P.ops[oi + 1 < P.n_ops ? oi + 1 : 0]) + tid);
            unsigned long long * pr = prof ? P.prof + ((size_t) oi * gridDim.x + blockIdx.x) * 8 : nullptr;
            if (pr) pr[0] = pk_now();
            switch (op.kind) {
                case PK_ROWS:   if (P.model == PKM_ORPHEUS) pk_rows_orpheus(P, step, st == 0, red); else if (P.model == PKM_DIA) pk_rows_dia(P, step, sids); else pk_rows(P, step, sids); break;
                case PK_GEMV:   if (op.q8) pk_gemv_q8<KVT>(P, op, ring, areg, red, full, empty, rp, step, pr, skv, sfp, spt);
                                else pk_gemv<KVT>(P, op, ring, reinterpret_cast<__half *>(areg), red, full, empty, rp, step, pr, skv, sfp, spt);
                                break;
                case PK_ATTN:   pk_attn<KVT, HD>(P, op, areg, step, sfp, spt); break;
                case PK_ATTNC:  pk_attn_combine<HD>(P, op); break;
                case PK_QUANT:  pk_quant(P, op, red); break;
                case PK_ARGMAX: if (P.model == PKM_ORPHEUS) pk_argmax_partial(P, red); else if (P.model == PKM_DIA) pk_argmax_dia(P, step, red); else pk_argmax(P, step, red); break;
            }
            if (pr) pr[2] = pk_now();
            if (tid < OPW) reinterpret_cast<uint4 *>(&sops[(opn & 1u) ^ 1u])[tid] = nxt;
            pk_grid_sync(P.bar, epoch);
            if (pr) pr[3] = pk_now();
        }
    }
    if (P.model == PKM_ORPHEUS)                                // the launch's last token (the other steps' tokens were combined by the following step's rows phase)
        for (int b = (int) blockIdx.x; b < P.R; b += (int) gridDim.x) pk_orpheus_token(P, P.step_begin + P.n_steps - 1, b, red);
    if (blockIdx.x == 0 && tid == 0) *P.d_step = P.step_begin + P.n_steps;
}

// prompt-pass K / V rows (fp32, [R0 rows][H] per layer, row r of sequence seq at position pos) -> the pages
template <typename KVT>
__global__ void pk_kv_import_kernel(const float * __restrict__ Kc, const float * __restrict__ Vc, size_t layer_stride, const int * __restrict__ row_src, const int * __restrict__ row_seq,
                                    const int * __restrict__ row_pos, const PkParams P) {
    const int r = blockIdx.x, l = blockIdx.y, seq = row_seq[r], pos = row_pos[r];
    const int KVW = P.kv_heads * P.hd;                           // floats per cache row (compact: kv heads, not query heads)
    const float * k = Kc + (size_t) l * layer_stride + (size_t) row_src[r] * KVW, * v = Vc + (size_t) l * layer_stride + (size_t) row_src[r] * KVW;
    for (int c = threadIdx.x; c < KVW; c += blockDim.x) {
        const int h = c / P.hd, d = c - h * P.hd;
        pk_store1(pk_page_row<KVT>(P, l, seq, pos, 0, h) + d, k[c]);
        pk_store1(pk_page_row<KVT>(P, l, seq, pos, 1, h) + d, v[c]);
    }
}

}  // namespace

static inline size_t pk_smem_bytes(int n_stages, int a_bytes, int pt_ints) { return (size_t) (16 + pt_ints) * 4 + (size_t) n_stages * PK_STAGE + (size_t) a_bytes + PK_RED_BYTES + 64 + 2 * PK_MAXSTAGES * sizeof(PkBar) + 3 * sizeof(PkOp) + 16 + 128 + 128; }

// ---------------------------------------------------------------- host side shared by the models' generate() functions

// shared-memory layout for a program whose widest activation chunk has KA columns (KAs: widest chunk of a split-matrix phase, 0 = none) and whose attention sees at
// most Tscore positions; picks the kernel instantiation for (cache element type, head size).  Returns 1 when the shape does not fit.
int pk_configure(PkParams & Pk, bool kv_f32, int KA, int KAs, int Tscore, PkLaunch & L) {
    size_t a = (size_t) 16 * (KA + PK_PAD) * 2;
    if (KAs) a = std::max(a, (size_t) 2 * 16 * (KAs + PK_PAD) * 2);
    a = std::max(a, 2 * pk_att_bytes(Tscore));
    a = (a + 255) & ~(size_t) 255;
    const int pt = Pk.R * Pk.max_pages;
    if (Pk.max_pages > 256) return 1;                          // page offsets of a sequence sit in a 256-entry scratch (pk_attn_item)
    int ns = PK_MAXSTAGES;
    while (ns > 2 && pk_smem_bytes(ns, (int) a, pt) > (size_t) 227 * 1024) ns--;
    if (pk_smem_bytes(ns, (int) a, pt) > (size_t) 227 * 1024) return 1;
    Pk.n_stages = ns; Pk.a_bytes = (int) a;
    L.smem = pk_smem_bytes(ns, (int) a, pt);
#ifdef B2EMU
#define PK_PICK(T, D) { L.kemu = [](const PkParams & q) { pdk_kernel<T, D>(q); }; }
#else
#define PK_PICK(T, D) { L.kfn = (const void *) pdk_kernel<T, D>; }
#endif
    const int hd = Pk.hd;
    if (hd != 8 && hd != 64 && hd != 128) return 1;
    if (kv_f32) { if (hd == 8) PK_PICK(float, 8) else if (hd == 64) PK_PICK(float, 64) else PK_PICK(float, 128) }
    else        { if (hd == 8) PK_PICK(__half, 8) else if (hd == 64) PK_PICK(__half, 64) else PK_PICK(__half, 128) }
#undef PK_PICK
#ifndef B2EMU
    if (cudaFuncSetAttribute(L.kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) L.smem) != cudaSuccess) { cudaGetLastError(); return 1; }
#endif
    return 0;
}
// one cooperative launch of steps [Pk.step_begin, Pk.step_begin + Pk.n_steps)
cudaError_t pk_launch(const PkLaunch & L, const PkParams & Pk, int grid, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(Pk.bar, 0, 256, st);
    if (e != cudaSuccess) return e;
#ifdef B2EMU
    const PkParams q = Pk;
    b2emu::launch_coop(dim3(grid), dim3(PK_THREADS), L.smem, [=]() { L.kemu(q); });
    return cudaSuccess;
#else
    void * args[] = {(void *) &Pk};
    return cudaLaunchCooperativeKernel(L.kfn, dim3(grid), dim3(PK_THREADS), args, L.smem, st);
#endif
}
// B2TTS_PDK_PROF=<step>: %globaltimer timeline of that decode step (every op x every CTA), written as raw uint64 to $B2TTS_PDK_PROF_FILE after the run
void pk_prof_begin(PkParams & Pk, size_t n_ops, int grid, cudaStream_t st) {
    const char * e = getenv("B2TTS_PDK_PROF");
    if (!e) return;
    const size_t words = n_ops * (size_t) grid * 8;
    if (cudaMalloc(&Pk.prof, words * 8) != cudaSuccess) { cudaGetLastError(); Pk.prof = nullptr; return; }
    cudaMemsetAsync(Pk.prof, 0, words * 8, st);
    Pk.prof_step = atoi(e);
}
void pk_prof_end(PkParams & Pk, const std::vector<PkOp> & ops, int grid, cudaStream_t st) {
    if (!Pk.prof) return;
    std::vector<unsigned long long> hp(ops.size() * (size_t) grid * 8);
    cudaMemcpyAsync(hp.data(), Pk.prof, hp.size() * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    if (const char * pf = getenv("B2TTS_PDK_PROF_FILE")) {
        if (FILE * f = fopen(pf, "wb")) {
            const int hdr[4] = {(int) ops.size(), grid, 8, Pk.prof_step};
            fwrite(hdr, 4, 4, f);
            for (const PkOp & o : ops) { const int k[4] = {o.kind, o.layer, o.K, o.n_units}; fwrite(k, 4, 4, f); }
            fwrite(hp.data(), 8, hp.size(), f); fclose(f);
        }
    }
    cudaFree(Pk.prof); Pk.prof = nullptr;
}


void pk_kv_import(bool kv_f32, const float * Kc, const float * Vc, size_t layer_stride, const int * row_src, const int * row_seq, const int * row_pos, const PkParams & P, int n_rows, int n_layers, cudaStream_t st) {
    dim3 grid(n_rows, n_layers);
    if (kv_f32) pk_kv_import_kernel<float><<<grid, 256, 0, st>>>(Kc, Vc, layer_stride, row_src, row_seq, row_pos, P);
    else pk_kv_import_kernel<__half><<<grid, 256, 0, st>>>(Kc, Vc, layer_stride, row_src, row_seq, row_pos, P);
}

}  // namespace b2
#endif  // B2_PDK_IMPLEMENTATION
