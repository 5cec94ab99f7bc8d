// common.cuh — shared helpers for libb200gan (error reporting, launch accounting, reductions).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200gan.h"

namespace b200 {

int set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define B200_CHECK_LAUNCH()                                                                   \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return b200::set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
        b200::count_launch();                                                                 \
    } while (0)

#define B200_REQUIRE(cond, ...)                                      \
    do {                                                             \
        if (!(cond)) return b200::set_error(__VA_ARGS__);            \
    } while (0)

constexpr int kNumSMs = 148;

static inline int grid_for(int64_t work_items, int block, int max_blocks_per_sm = 8) {
    int64_t g = (work_items + block - 1) / block;
    int64_t cap = (int64_t)kNumSMs * max_blocks_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

static inline cudaStream_t as_stream(b200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// ---- activation storage types: channel-last activations are fp32 (B200_F32) or bf16 (B200_BF16) ----------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive elements (pointer must be 4-element aligned: 16 B for fp32, 8 B for bf16)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

// ---------------------------------------------------------------------------------------------------------
// 16-byte vector access: V = 4 fp32 or 8 bf16 consecutive channels per thread
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct VecIO;
template <> struct VecIO<float> {
    static constexpr int V = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct VecIO<bf16> {
    static constexpr int V = 8;
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {            // bf16 -> fp32 is a 16-bit shift
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
    static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
        uint4 q;
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]),
                       c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
        q.x = *reinterpret_cast<uint32_t*>(&a); q.y = *reinterpret_cast<uint32_t*>(&b);
        q.z = *reinterpret_cast<uint32_t*>(&c); q.w = *reinterpret_cast<uint32_t*>(&d);
        *reinterpret_cast<uint4*>(p) = q;
    }
};
template <int V>
__device__ __forceinline__ void ldp(const float* p, float (&v)[V]) {       // V fp32 parameters (16-byte aligned)
#pragma unroll
    for (int i = 0; i < V; i += 4) {
        const float4 q = *reinterpret_cast<const float4*>(p + i);
        v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
    }
}
// dispatch helper: run `BODY` with `T` bound to the storage type selected by an int dtype code
#define B200_DISPATCH_DT(dt, T, ...)                          \
    do {                                                      \
        if ((dt) == B200_BF16) { typedef b200::bf16 T; __VA_ARGS__; } \
        else { typedef float T; __VA_ARGS__; }                \
    } while (0)

}  // namespace b200
