// tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A functional CPU emulation of the small CUDA subset the plain (CUDA-core) kernels of this library use, so that their LOGIC -- indexing,
// reductions, barriers, host-side orchestration -- can be checked against the golden vectors in the GPU-less build container
// (tests/test_emu_cpu.py).  A kernel launch runs its blocks one after another; the threads of a block are cooperative fibers, so
// __syncthreads() and warp shuffles have their real meaning.  Nothing in the product links or includes this; it says nothing about
// performance, memory coalescing or hardware-only features (tcgen05, TMA, mbarriers, clusters are not emulated).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() {}
    dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
    dim3(int a) : x((unsigned) a) {}
    dim3(long a) : x((unsigned) a) {}
    dim3(size_t a) : x((unsigned) a) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
typedef struct b2emu_stream * cudaStream_t;
typedef struct b2emu_event * cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };

static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char * cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMalloc(void ** p, size_t n) { *p = nullptr; return posix_memalign(p, 256, n ? n : 1) ? 2 : 0; }   // exact size (sanitizer red zones)
template <class T> static inline cudaError_t cudaMalloc(T ** p, size_t n) { return cudaMalloc((void **) p, n); }
static inline cudaError_t cudaFree(void * p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void ** p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T ** p, size_t n) { return cudaMalloc((void **) p, n); }
static inline cudaError_t cudaFreeHost(void * p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void * d, const void * s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void * d, const void * s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void * d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void * d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t * e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float * ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

// stream capture: launches issued between Begin/EndCapture are recorded (closures own their arguments, like kernel parameters) and replayed by cudaGraphLaunch
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
typedef struct b2emu_graph * cudaGraph_t;
typedef struct b2emu_graph * cudaGraphExec_t;
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode);
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t * g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t * e, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e);
cudaError_t cudaGraphDestroy(cudaGraph_t g);

namespace b2emu {
struct Fiber { uint3 tid; void * sp; char * stack; bool done; dim3 bidx; int blk; };     // bidx: blockIdx of the thread's block; blk: its slot among the blocks alive at once
extern Fiber * g_cur;
extern dim3 g_blockDim, g_gridDim;
extern uint64_t g_launches, g_blocks, g_replays;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> & body);
void launch_coop(dim3 grid, dim3 block, size_t smem, const std::function<void()> & body);   // every block of the grid alive at once (cudaLaunchCooperativeKernel)
void sync_block();
void named_bar(int id, int nthreads);      // bar.sync id, nthreads
void yield_spin();                         // inside a polling loop on memory another thread writes
void note_progress();                      // by the writer of such memory
uint64_t shfl(uint64_t v, int src_lane);   // every live lane of the warp calls it; returns the value lane src_lane passed
void * dyn_smem();
void warp_exchange(const unsigned * mine, int nwords, unsigned * all /* [32][nwords] */);   // every live lane publishes nwords and receives the whole warp's
}  // namespace b2emu

#define threadIdx (b2emu::g_cur->tid)
#define blockIdx  (b2emu::g_cur->bidx)
#define blockDim  (b2emu::g_blockDim)
#define gridDim   (b2emu::g_gridDim)

static inline void __syncthreads() { b2emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { (void) b2emu::shfl(0, 0); }
static inline int b2emu_lane() { return (int) ((threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y) & 31); }
template <class T> static inline T b2emu_shfl_t(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of a wide type");
    uint64_t u = 0; memcpy(&u, &v, sizeof(T));
    u = b2emu::shfl(u, src & 31);
    T r; memcpy(&r, &u, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return b2emu_shfl_t(v, b2emu_lane() ^ o); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return b2emu_shfl_t(v, src); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { const int l = b2emu_lane(); return b2emu_shfl_t(v, l + d < 32 ? l + d : l); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { const int l = b2emu_lane(); return b2emu_shfl_t(v, l - d >= 0 ? l - d : l); }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
template <class T> static inline T __ldcs(const T * p) { return *p; }
static inline int __dp4a(int a, int b, int c) { for (int i = 0; i < 4; i++) c += (int) (int8_t) (a >> (8 * i)) * (int) (int8_t) (b >> (8 * i)); return c; }
static inline unsigned __vsub4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; i++) r |= (((a >> (8 * i)) - (b >> (8 * i))) & 0xffu) << (8 * i); return r; }
// the explicitly rounded forms: one IEEE operation each, never contracted into an FMA by the host compiler (the volatile result pins the rounding)
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline int __float2int_rn(float x) { return (int) nearbyintf(x); }      // round-to-nearest-even (the default rounding mode)
// fibers are cooperative (a thread runs until its next barrier / shuffle), so a read-modify-write is atomic as it stands
static inline unsigned atomicAdd(unsigned * p, unsigned v) { const unsigned o = *p; *p = o + v; b2emu::note_progress(); return o; }
static inline int atomicAdd(int * p, int v) { const int o = *p; *p = o + v; b2emu::note_progress(); return o; }
static inline void __threadfence() {}
static inline int __clz(int x) { return x ? __builtin_clz((unsigned) x) : 32; }
template <class T> static inline T __ldcg(const T * p) { return *p; }
template <class T> static inline void __stcg(T * p, T v) { *p = v; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T * p) { return *p; }
