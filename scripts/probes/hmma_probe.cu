// hmma_probe.cu -- issue rate of the legacy warp-level tensor-core instruction (mma.sync.m16n8k16 f16 x f16 -> f32) on sm_100a, for the persistent decode kernel's GEMV
// (8 consumer warps per SM, 2 per scheduler).  Prints TFLOP/s and cycles per MMA per scheduler for 1..4 warps per scheduler and 1, 2, 4 independent accumulators per warp.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float * c, const unsigned * a, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int CH>
__global__ void k(float * out, int iters, long long * cyc) {
    float c[CH][4];
    for (int j = 0; j < CH; j++) for (int e = 0; e < 4; e++) c[j][e] = 0.f;
    unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}; unsigned b0 = threadIdx.x * 5u, b1 = 11u;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < CH; j++) mma(c[j], a, b0, b1);
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < CH; j++) for (int e = 0; e < 4; e++) s += c[j][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH> void run(int warps_per_sm, float * out, long long * cyc) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH><<<148, warps_per_sm * 32>>>(out, 100, cyc);
    cudaEventRecord(e0);
    k<CH><<<148, warps_per_sm * 32>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double n_mma = 148.0 * warps_per_sm * iters * CH;
    printf("warps/SM %2d  chains/warp %d: %.1f TFLOP/s   %.1f cycles per MMA per scheduler (warp 0: %.1f cycles per its MMA)\n", warps_per_sm, CH, n_mma * 4096 / (ms * 1e-3) / 1e12,
           (double) h / ((double) iters * CH * (warps_per_sm / 4.0)), (double) h / ((double) iters * CH));
}
int main() {
    float * out; long long * cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int w : {4, 8, 16}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); }
    return 0;
}
