// tests/emu/include/cuda_fp16.h -- TEST INFRASTRUCTURE ONLY: __half on the host through the compiler's _Float16 (see cuda_runtime.h here).
#pragma once
#include "cuda_runtime.h"
struct __half_raw { unsigned short x; };
struct alignas(2) __half {
    unsigned short x = 0;
    __half() {}
    __half(const __half_raw & r) : x(r.x) {}
    __half(float f) { _Float16 h = (_Float16) f; memcpy(&x, &h, 2); }
    operator float() const { _Float16 h; memcpy(&h, &x, 2); return (float) h; }
};
struct alignas(4) __half2 { __half x, y; };
static inline float __half2float(__half h) { return (float) h; }
static inline __half __float2half(float f) { return __half(f); }
static inline __half __float2half_rn(float f) { return __half(f); }
static inline __half2 __floats2half2_rn(float a, float b) { __half2 r; r.x = __half(a); r.y = __half(b); return r; }
static inline float2 __half22float2(__half2 h) { return float2{(float) h.x, (float) h.y}; }
