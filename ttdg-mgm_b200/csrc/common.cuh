// common.cuh - shared helpers for libttdg_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/ttdg_b200.h"

#define TTDG_WARP 32
#define TTDG_FULL 0xffffffffu

#define TTDG_CHECK_ARG(cond) do { if (!(cond)) return TTDG_E_ARG; } while (0)
#define TTDG_LAUNCH_RET() do { cudaError_t e__ = cudaGetLastError(); return (int)e__; } while (0)

namespace ttdg {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TTDG_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TTDG_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(TTDG_FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(TTDG_FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TTDG_FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(TTDG_FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(TTDG_FULL, v, o));
    return v;
}
// named barrier for a sub-group of warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

void count_launches(int n);      // api.cu: kernels launched through the C ABI (bench.py reports it)
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace ttdg

#define TTDG_STR2(x) #x
#define TTDG_STR(x) TTDG_STR2(x)
