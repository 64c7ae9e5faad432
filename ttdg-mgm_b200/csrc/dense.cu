// dense.cu - small dense layers of the matching head with fp64 accumulation.
// Reference call sites: attention q/k projections (utils/attentions.py:66-69), affinity projections and
// MLP (utils/affinity.py:48-54), universe init (multi_graph_matching.py:531-532), and their autograd.
// These are tiny (M <= ~800 rows, 256..1024 features): one generic tiled kernel, fp64 FMA, one rounding
// at the output, so the results are independent of tiling / summation order to ~1e-16.
#include "common.cuh"

namespace ttdg {

constexpr int GK = 16;      // k-slab
// C tile is GT x GT, 256 threads, (GT / 16)^2 outputs per thread.  GT = 64 normally; GT = 32 when 64-wide tiles would leave
// most SMs without a CTA (the matching head's GEMMs are ~280 x 512 x 256: 40 tiles of 64 x 64 on 148 SMs).  Every output
// is accumulated over k in ascending order whatever the tile: results do not depend on GT.

__device__ __forceinline__ double ld_any(const void *p, int is_f64, size_t idx) {
    return is_f64 ? reinterpret_cast<const double *>(p)[idx] : (double)reinterpret_cast<const float *>(p)[idx];
}

// C (m x n) = op(A) (m x k) * op(B) (k x n) [+ bias(n)] [+ C];  op(A)(i,kk) = transA ? A[kk*lda+i] : A[i*lda+kk]
template <int GT>
__global__ void __launch_bounds__(256)
gemm_f64acc_kernel(int transA, int transB, int m, int n, int k, const void *__restrict__ A, int a64, int lda,
                   const void *__restrict__ B, int b64, int ldb, void *__restrict__ C, int c64, int ldc,
                   const float *__restrict__ bias, int accumulate) {
    __shared__ double As[GK][GT + 1];
    __shared__ double Bs[GK][GT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row0 = blockIdx.y * GT, col0 = blockIdx.x * GT;
    constexpr int TM = GT / 16;
    double acc[TM][TM];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < k; k0 += GK) {
        for (int e = threadIdx.x; e < GK * GT; e += 256) {
            // pick the index order that walks the contiguous dimension of the operand
            int kk, i;
            if (transA) { i = e % GT; kk = e / GT; } else { kk = e % GK; i = e / GK; }
            const int gi = row0 + i, gk = k0 + kk;
            As[kk][i] = (gi < m && gk < k) ? ld_any(A, a64, transA ? (size_t)gk * lda + gi : (size_t)gi * lda + gk) : 0.0;
        }
        for (int e = threadIdx.x; e < GK * GT; e += 256) {
            int kk, j;
            if (transB) { kk = e % GK; j = e / GK; } else { j = e % GT; kk = e / GT; }
            const int gj = col0 + j, gk = k0 + kk;
            Bs[kk][j] = (gj < n && gk < k) ? ld_any(B, b64, transB ? (size_t)gj * ldb + gk : (size_t)gk * ldb + gj) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double a[TM], b[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < TM; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TM; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gi = row0 + ty + 16 * i;
        if (gi >= m) continue;
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            const int gj = col0 + tx + 16 * j;
            if (gj >= n) continue;
            double v = acc[i][j];
            if (bias) v += (double)bias[gj];
            const size_t idx = (size_t)gi * ldc + gj;
            if (c64) {
                double *c = reinterpret_cast<double *>(C);
                c[idx] = accumulate ? c[idx] + v : v;
            } else {
                float *c = reinterpret_cast<float *>(C);
                c[idx] = accumulate ? (float)((double)c[idx] + v) : (float)v;
            }
        }
    }
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_gemm_f64acc(int transA, int transB, int m, int n, int k, const void *A, int a_is_f64, int lda,
                                const void *B, int b_is_f64, int ldb, void *C, int c_is_f64, int ldc, int accumulate,
                                void *stream) {
    TTDG_CHECK_ARG(A && B && C && m >= 0 && n >= 0 && k >= 0);
    if (m == 0 || n == 0) return 0;
    ttdg::count_launches(1);
    if (ceil_div(n, 64) * ceil_div(m, 64) >= 148)
        gemm_f64acc_kernel<64><<<dim3(ceil_div(n, 64), ceil_div(m, 64)), 256, 0, (cudaStream_t)stream>>>(
            transA, transB, m, n, k, A, a_is_f64, lda, B, b_is_f64, ldb, C, c_is_f64, ldc, nullptr, accumulate);
    else
        gemm_f64acc_kernel<32><<<dim3(ceil_div(n, 32), ceil_div(m, 32)), 256, 0, (cudaStream_t)stream>>>(
            transA, transB, m, n, k, A, a_is_f64, lda, B, b_is_f64, ldb, C, c_is_f64, ldc, nullptr, accumulate);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_linear_f64acc(const float *X, int ldx, const float *W, int ldw, const float *b, float *Y, int ldy,
                                  int m, int n, int k, void *stream) {
    TTDG_CHECK_ARG(X && W && Y && m >= 0 && n >= 0 && k >= 0);
    if (m == 0 || n == 0) return 0;
    ttdg::count_launches(1);
    if (ceil_div(n, 64) * ceil_div(m, 64) >= 148)
        gemm_f64acc_kernel<64><<<dim3(ceil_div(n, 64), ceil_div(m, 64)), 256, 0, (cudaStream_t)stream>>>(0, 1, m, n, k, X, 0, ldx, W, 0, ldw, Y,
                                                                                                           0, ldy, b, 0);
    else
        gemm_f64acc_kernel<32><<<dim3(ceil_div(n, 32), ceil_div(m, 32)), 256, 0, (cudaStream_t)stream>>>(0, 1, m, n, k, X, 0, ldx, W, 0, ldw, Y,
                                                                                                           0, ldy, b, 0);
    TTDG_LAUNCH_RET();
}
