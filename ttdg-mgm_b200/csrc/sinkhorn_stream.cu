// sinkhorn_stream.cu - log-space Sinkhorn for LARGE matrices (the N = 256 / 512 / 1024 micro-benchmark of
// BASELINE.json configs[4]); fp32.  Same operator as utils/sinkhorn.py:58-87 -> pygmtools.sinkhorn
//  without dummy rows.
//
// B200 design: one thread-block CLUSTER of R CTAs per matrix; CTA r keeps a slab of rows in shared memory for
// ALL iterations (rows that do not fit are re-read through L2, never re-written).  The matrix is never
// rewritten: with row potentials f and column potentials g the normalised log-matrix is
//       z_ij = x_ij / tau - f_i - g_j,
//   row step   f_i = logsumexp_j (x_ij / tau - g_j)        (local to the CTA owning row i)
//   col step   g_j = logsumexp_i (x_ij / tau - f_i)        (per-CTA partial (max, sum) -> reduced across the
//                                                           cluster through distributed shared memory)
// which is algebraically the reference's "z -= logsumexp(z)" alternation.  HBM traffic is one read of x and
// one write of exp(z) in total instead of one read + one write per half-iteration.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace ttdg {

constexpr int SKS_THREADS = 1024;
constexpr int SKS_WARPS = SKS_THREADS / 32;
constexpr int SKS_SLAB_BYTES = 192 * 1024;

struct SksParams {
    const float *s;
    float *out;
    int batch, n1, n2, R, rows_per_cta, res_rows, max_iter;
    float inv_tau;
};

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

__global__ void __launch_bounds__(SKS_THREADS, 1)
sinkhorn_stream_kernel(const SksParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int cid = blockIdx.x / p.R, ncl = gridDim.x / p.R;
    extern __shared__ __align__(16) unsigned char sks_smem[];
    const int n2 = p.n2;
    float *slab = reinterpret_cast<float *>(sks_smem);            // res_rows x n2 (already scaled by 1/tau)
    float *g = slab + (size_t)p.res_rows * n2;                    // n2      column potentials (replicated per CTA)
    float *f = g + n2;                                            // rows_per_cta
    float *cm = f + p.rows_per_cta;                               // n2      this CTA's partial column max
    float *cs = cm + n2;                                          // n2      this CTA's partial column sum
    float *pm = cs + n2;                                          // RG x n2 per-row-group partials
    const int CW = n2 < SKS_THREADS ? n2 : SKS_THREADS;           // threads across columns in the column step
    const int RG = SKS_THREADS / CW;                              // row groups
    float *ps = pm + (size_t)RG * n2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = r * p.rows_per_cta;
    const int nrows = max(0, min(p.rows_per_cta, p.n1 - row0));
    const int cols_per_cta = (n2 + p.R - 1) / p.R;

    for (int b = cid; b < p.batch; b += ncl) {
        const float *x = p.s + (size_t)b * p.n1 * n2 + (size_t)row0 * n2;
        // ---- load the resident slab (scaled)
        const int nres = min(nrows, p.res_rows);
        for (int e = tid * 4; e < nres * n2; e += SKS_THREADS * 4) {
            float4 v = ld4(x + e);
            v.x *= p.inv_tau; v.y *= p.inv_tau; v.z *= p.inv_tau; v.w *= p.inv_tau;
            *reinterpret_cast<float4 *>(slab + e) = v;
        }
        for (int j = tid; j < n2; j += SKS_THREADS) g[j] = 0.f;
        for (int i = tid; i < p.rows_per_cta; i += SKS_THREADS) f[i] = 0.f;
        cluster.sync();                                            // peers may still read g / cm / cs of the previous matrix

        for (int it = 0; it < p.max_iter; ++it) {
            if ((it & 1) == 0) {
                // ---------------- row step: f_i = lse_j (t_ij - g_j), warp per row
                for (int i = warp; i < nrows; i += SKS_WARPS) {
                    const bool res = i < p.res_rows;
                    const float *row = res ? slab + (size_t)i * n2 : x + (size_t)i * n2;
                    const float sc = res ? 1.f : p.inv_tau;
                    float m = -INFINITY;
                    for (int j = lane * 4; j < n2; j += 128) {
                        const float4 v = ld4(row + j), gv = ld4(g + j);
                        m = fmaxf(m, fmaxf(fmaxf(v.x * sc - gv.x, v.y * sc - gv.y), fmaxf(v.z * sc - gv.z, v.w * sc - gv.w)));
                    }
                    m = warp_max(m);
                    float sum = 0.f;
                    for (int j = lane * 4; j < n2; j += 128) {
                        const float4 v = ld4(row + j), gv = ld4(g + j);
                        sum += __expf(v.x * sc - gv.x - m) + __expf(v.y * sc - gv.y - m) + __expf(v.z * sc - gv.z - m) +
                               __expf(v.w * sc - gv.w - m);
                    }
                    sum = warp_sum(sum);
                    if (lane == 0) f[i] = m + __logf(sum);
                }
                __syncthreads();
            } else {
                // ---------------- column step: partial (max, sum) over this CTA's rows
                const int cx = tid % CW, ry = tid / CW;
                for (int j = cx; j < n2; j += CW) {
                    float m = -INFINITY;
                    for (int i = ry; i < nrows; i += RG) {
                        const float t = i < p.res_rows ? slab[(size_t)i * n2 + j] : __ldg(x + (size_t)i * n2 + j) * p.inv_tau;
                        m = fmaxf(m, t - f[i]);
                    }
                    float sum = 0.f;
                    if (m > -INFINITY)
                        for (int i = ry; i < nrows; i += RG) {
                            const float t = i < p.res_rows ? slab[(size_t)i * n2 + j] : __ldg(x + (size_t)i * n2 + j) * p.inv_tau;
                            sum += __expf(t - f[i] - m);
                        }
                    pm[(size_t)ry * n2 + j] = m; ps[(size_t)ry * n2 + j] = sum;
                }
                __syncthreads();
                for (int j = tid; j < n2; j += SKS_THREADS) {
                    float m = pm[j];
                    for (int q = 1; q < RG; ++q) m = fmaxf(m, pm[(size_t)q * n2 + j]);
                    float sum = 0.f;
                    if (m > -INFINITY)
                        for (int q = 0; q < RG; ++q) sum += ps[(size_t)q * n2 + j] * __expf(pm[(size_t)q * n2 + j] - m);
                    cm[j] = m; cs[j] = sum;
                }
                cluster.sync();                                    // partials of every CTA are published
                // CTA r reduces its slice of columns over the cluster and broadcasts g_j to every CTA
                for (int jj = tid; jj < cols_per_cta; jj += SKS_THREADS) {
                    const int j = r * cols_per_cta + jj;
                    if (j < n2) {
                        float m = -INFINITY;
                        float mq[16], sq[16];
                        for (int q = 0; q < p.R; ++q) {
                            mq[q] = *cluster.map_shared_rank(cm + j, q);
                            sq[q] = *cluster.map_shared_rank(cs + j, q);
                            m = fmaxf(m, mq[q]);
                        }
                        float sum = 0.f;
                        for (int q = 0; q < p.R; ++q) sum += sq[q] * __expf(mq[q] - m);
                        const float gj = m + __logf(sum);
                        for (int q = 0; q < p.R; ++q) *cluster.map_shared_rank(g + j, q) = gj;
                    }
                }
                cluster.sync();                                    // new g visible everywhere
            }
        }
        // ---- out = exp(t - f - g)
        float *o = p.out + (size_t)b * p.n1 * n2 + (size_t)row0 * n2;
        for (int e = tid * 4; e < nrows * n2; e += SKS_THREADS * 4) {
            const int i = e / n2, j = e - i * n2;
            float4 v;
            if (i < p.res_rows) v = *reinterpret_cast<const float4 *>(slab + e);
            else { v = ld4(x + e); v.x *= p.inv_tau; v.y *= p.inv_tau; v.z *= p.inv_tau; v.w *= p.inv_tau; }
            const float4 gv = ld4(g + j);
            const float fi = f[i];
            v.x = __expf(v.x - fi - gv.x); v.y = __expf(v.y - fi - gv.y); v.z = __expf(v.z - fi - gv.z); v.w = __expf(v.w - fi - gv.w);
            *reinterpret_cast<float4 *>(o + e) = v;
        }
        __syncthreads();
    }
    cluster.sync();                                                // no CTA exits while peers may still touch its smem
}

static void sks_plan(int n1, int n2, SksParams &p, size_t &smem) {
    int R = 8;
    while (R > 1 && n1 / R < 32) R >>= 1;
    p.R = R;
    p.rows_per_cta = (n1 + R - 1) / R;
    int res = SKS_SLAB_BYTES / (n2 * (int)sizeof(float));
    p.res_rows = res < p.rows_per_cta ? res : p.rows_per_cta;
    const int CW = n2 < SKS_THREADS ? n2 : SKS_THREADS;
    const int RG = SKS_THREADS / CW;
    smem = ((size_t)p.res_rows * n2 + n2 + p.rows_per_cta + 2 * (size_t)n2 + 2 * (size_t)RG * n2) * sizeof(float);
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int64_t ttdg_sinkhorn_stream_scratch_bytes(int batch, int n1, int n2) {
    (void)batch; (void)n1; (void)n2;
    return 0;      // the matrix lives in distributed shared memory; no global scratch
}

extern "C" int ttdg_sinkhorn_stream_fwd(const float *s, float *out, int batch, int n1, int n2, float tau, int max_iter,
                                        int dummy_row, void *scratch, void *stream) {
    (void)scratch;
    TTDG_CHECK_ARG(s && out && batch >= 0 && n1 >= 1 && n2 >= 1 && tau > 0.f && max_iter >= 0);
    if (n1 > n2 || (n2 & 3) || n2 > 4096 || (dummy_row && n1 != n2)) return TTDG_E_LIMIT;
    if (batch == 0) return 0;
    SksParams p;
    size_t smem;
    sks_plan(n1, n2, p, smem);
    if (smem > 227 * 1024) return TTDG_E_LIMIT;
    p.s = s; p.out = out; p.batch = batch; p.n1 = n1; p.n2 = n2; p.max_iter = max_iter; p.inv_tau = 1.0f / tau;
    cudaError_t e = cudaFuncSetAttribute(sinkhorn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(SKS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.R; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(p.R);
    int max_clusters = 0;
    e = cudaOccupancyMaxActiveClusters(&max_clusters, sinkhorn_stream_kernel, &cfg);
    if (e != cudaSuccess || max_clusters < 1) { cudaGetLastError(); max_clusters = 148 / p.R; }
    const int ncl = batch < max_clusters ? batch : max_clusters;
    cfg.gridDim = dim3(ncl * p.R);
    e = cudaLaunchKernelEx(&cfg, sinkhorn_stream_kernel, p);
    return (int)e;
}
