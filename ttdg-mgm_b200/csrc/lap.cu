// lap.cu - batched hungarian(): one warp per score matrix.  Reference: utils/hungarian.py:8-65.
#include "lap.cuh"

namespace ttdg {

constexpr int LAP_WARPS = 4;

__global__ void __launch_bounds__(LAP_WARPS * 32)
lap_kernel(const float *__restrict__ s, float *__restrict__ perm, const int64_t *__restrict__ items, int n_items) {
    extern __shared__ __align__(16) unsigned char lap_smem[];
    LapWork *work = reinterpret_cast<LapWork *>(lap_smem);
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * LAP_WARPS + warp;
    if (b >= n_items) return;
    const int64_t *d = items + (size_t)b * 6;
    const int n1 = (int)d[2], n2 = (int)d[3];
    if (n1 > LAP_MAX_DIM || n2 > LAP_MAX_DIM) return;
    if ((threadIdx.x & 31) == 0) { work[warp].stat_steps = 0; work[warp].stat_hops = 0; }
    hungarian_warp(s + d[0], perm + d[1], n1, n2, (int)d[4], (int)d[5], work[warp]);
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_lap_solve(const float *s, float *perm, const int64_t *items, int n_items, void *stream) {
    TTDG_CHECK_ARG(s && perm && items && n_items >= 0);
    if (n_items == 0) return 0;
    const size_t smem = sizeof(LapWork) * LAP_WARPS;
    cudaError_t e = cudaFuncSetAttribute(lap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ttdg::count_launches(1);
    lap_kernel<<<ceil_div(n_items, LAP_WARPS), LAP_WARPS * 32, smem, (cudaStream_t)stream>>>(s, perm, items, n_items);
    TTDG_LAUNCH_RET();
}
