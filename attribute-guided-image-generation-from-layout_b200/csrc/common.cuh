// common.cuh — shared helpers for libb200gan (error reporting, launch accounting, reductions).
#pragma once
#include <cuda_runtime.h>
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

}  // namespace b200
