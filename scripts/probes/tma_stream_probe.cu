// tma_stream_probe.cu -- how fast can one elected thread per SM stream weight tiles HBM -> shared memory with cp.async.bulk, as a function of the copy shape?
//   mode 0: a tile = 8 row segments of 2 KB, rows K * 2 bytes apart (the row-major [N][K] fp16 matrix of the persistent decode kernel, K = 3072)
//   mode 1: a tile = ONE contiguous 16 KB copy (a pre-tiled matrix)
//   mode 2: a tile = 2 contiguous 8 KB copies
// consumers (8 warps) only wait for the tile and release the stage.  Prints GB/s per (mode, stages).   nvcc -arch=sm_100a -O3 -o tma_probe tma_stream_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t * b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_arrive(uint64_t * b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t * b, unsigned n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t * b, unsigned ph) {
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void * d, const void * s, unsigned n, uint64_t * b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}

constexpr int TILE = 16384, ROWB = 2048 + 64;

__global__ void __launch_bounds__(288, 1) probe(const unsigned char * W, size_t K2, int tiles_per_cta, int S, int mode, unsigned * sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t * full = reinterpret_cast<uint64_t *>(sm), * empty = full + 16;
    unsigned char * ring = sm + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int s = 0; s < S; s++) { mb_init(&full[s], 1); mb_init(&empty[s], 8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    int s = 0; unsigned ph = 0;
    if (warp == 8) {
        if (lane == 0) {
            for (int i = 0; i < tiles_per_cta; i++) {
                const size_t tile = (size_t) blockIdx.x + (size_t) i * gridDim.x;
                mb_wait(&empty[s], ph ^ 1u);
                mb_expect(&full[s], TILE);
                unsigned char * dst = ring + (size_t) s * (8 * ROWB);
                if (mode == 0) {
                    // unit u = tile / 3, k tile t = tile % 3 of a [N][3072] fp16 matrix: rows 8u .. 8u+7, columns 1024 t ..
                    const size_t u = tile / 3, t = tile % 3;
                    const unsigned char * src = W + (u * 8) * K2 + t * 2048;
#pragma unroll
                    for (int r = 0; r < 8; r++) bulk(dst + r * ROWB, src + r * K2, 2048, &full[s]);
                } else if (mode == 1) {
                    bulk(dst, W + tile * TILE, TILE, &full[s]);
                } else {
                    bulk(dst, W + tile * TILE, TILE / 2, &full[s]);
                    bulk(dst + TILE / 2, W + tile * TILE + TILE / 2, TILE / 2, &full[s]);
                }
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }
    unsigned acc = 0;
    for (int i = 0; i < tiles_per_cta; i++) {
        mb_wait(&full[s], ph);
        acc += *reinterpret_cast<const unsigned *>(ring + (size_t) s * (8 * ROWB) + lane * 4);
        __syncwarp();
        if (lane == 0) mb_arrive(&empty[s]);
        if (++s == S) { s = 0; ph ^= 1u; }
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const size_t bytes = (size_t) 4 << 30;                      // 4 GB of "weights": far beyond L2
    unsigned char * W; unsigned * sink;
    cudaMalloc(&W, bytes); cudaMalloc(&sink, 4); cudaMemset(W, 1, bytes);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int tiles_per_cta = 1600;                             // 148 x 1600 x 16 KB = 3.9 GB
    for (int mode = 0; mode < 3; mode++)
        for (int S : {2, 4, 6, 9, 12}) {
            const size_t smem = 256 + (size_t) S * 8 * ROWB;
            float best = 1e9f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                probe<<<sms, 288, smem>>>(W, 6144, tiles_per_cta, S, mode, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const cudaError_t err = cudaGetLastError();
            printf("mode %d (%s) stages %2d: %.3f ms  %.0f GB/s  %s\n", mode, mode == 0 ? "8 x 2 KB rows, 6 KB apart" : mode == 1 ? "1 x 16 KB contiguous" : "2 x 8 KB contiguous", S, best,
                   (double) sms * tiles_per_cta * TILE / (best * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
    return 0;
}
