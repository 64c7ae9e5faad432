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

constexpr int SKS_THREADS = 512;        // 16 warps: 128 registers per thread for the register-resident row
constexpr int SKS_WARPS = SKS_THREADS / 32;

struct SksParams {
    const float *s;
    float *out;
    int batch, n1, n2, R, rows_per_cta, res_rows, max_iter;
    float inv_tau;
};

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max4(const float4 &v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }

// Everything lives in the log2 domain: slab = x * log2(e) / tau, potentials f, g in log2 units, exp = ex2.approx
// (one MUFU op, no multiply), out = 2^(t - f - g).
__global__ void __launch_bounds__(SKS_THREADS, 1)
sinkhorn_stream_kernel(const SksParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int cid = blockIdx.x / p.R, ncl = gridDim.x / p.R;
    extern __shared__ __align__(16) unsigned char sks_smem[];
    const int n2 = p.n2;
    float *slab = reinterpret_cast<float *>(sks_smem);            // res_rows x n2 (scaled)
    float *g = slab + (size_t)p.res_rows * n2;                    // n2      column potentials (replicated per CTA)
    float *f = g + n2;                                            // rows_per_cta (padded to a multiple of 4)
    float *cm = f + ((p.rows_per_cta + 3) & ~3);                  // n2      this CTA's partial column max
    float *cs = cm + n2;                                          // n2      this CTA's partial column sum
    float *pm = cs + n2;                                          // RG x n2 per-row-group partials
    const int n2q = n2 >> 2;                                      // float4 columns
    const int CW = n2q < SKS_THREADS ? n2q : SKS_THREADS;         // threads across float4 columns in the column step
    const int RG = SKS_THREADS / CW;                              // row groups
    float *ps = pm + (size_t)RG * n2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = r * p.rows_per_cta;
    const int nrows = max(0, min(p.rows_per_cta, p.n1 - row0));
    const int cols_per_cta = (n2 + p.R - 1) / p.R;
    const float sc = p.inv_tau;                                   // log2(e) / tau

    for (int b = cid; b < p.batch; b += ncl) {
        const float *x = p.s + (size_t)b * p.n1 * n2 + (size_t)row0 * n2;
        const int nres = min(nrows, p.res_rows);
        for (int e = tid * 4; e < nres * n2; e += SKS_THREADS * 4) {
            float4 v = ld4(x + e);
            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
            *reinterpret_cast<float4 *>(slab + e) = v;
        }
        for (int j = tid; j < n2; j += SKS_THREADS) g[j] = 0.f;
        for (int i = tid; i < p.rows_per_cta; i += SKS_THREADS) f[i] = 0.f;
        cluster.sync();                                            // peers may still read g / cm / cs of the previous matrix

        for (int it = 0; it < p.max_iter; ++it) {
            if ((it & 1) == 0) {
                // ---------------- row step: f_i = lse_j (t_ij - g_j); warp per row, the row is held in registers
                for (int i = warp; i < nrows; i += SKS_WARPS) {
                    const bool res = i < p.res_rows;
                    const float *row = res ? slab + (size_t)i * n2 : x + (size_t)i * n2;
                    const float rs = res ? 1.f : sc;
                    float m = -INFINITY, sum = 0.f;
                    if (n2 <= 1024) {
                        float4 v[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const int j = c * 128 + lane * 4;
                            if (j < n2) {
                                const float4 t = ld4(row + j), gv = ld4(g + j);
                                v[c] = make_float4(fmaf(t.x, rs, -gv.x), fmaf(t.y, rs, -gv.y), fmaf(t.z, rs, -gv.z), fmaf(t.w, rs, -gv.w));
                                m = fmaxf(m, max4(v[c]));
                            }
                        }
                        m = warp_max(m);
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (c * 128 + lane * 4 < n2) sum += ex2(v[c].x - m) + ex2(v[c].y - m) + ex2(v[c].z - m) + ex2(v[c].w - m);
                    } else {
                        for (int j = lane * 4; j < n2; j += 128) {
                            const float4 t = ld4(row + j), gv = ld4(g + j);
                            m = fmaxf(m, fmaxf(fmaxf(fmaf(t.x, rs, -gv.x), fmaf(t.y, rs, -gv.y)), fmaxf(fmaf(t.z, rs, -gv.z), fmaf(t.w, rs, -gv.w))));
                        }
                        m = warp_max(m);
                        for (int j = lane * 4; j < n2; j += 128) {
                            const float4 t = ld4(row + j), gv = ld4(g + j);
                            sum += ex2(fmaf(t.x, rs, -gv.x) - m) + ex2(fmaf(t.y, rs, -gv.y) - m) + ex2(fmaf(t.z, rs, -gv.z) - m) +
                                   ex2(fmaf(t.w, rs, -gv.w) - m);
                        }
                    }
                    sum = warp_sum(sum);
                    if (lane == 0) f[i] = m + lg2(sum);
                }
                __syncthreads();
            } else {
                // ---------------- column step: partial (max, sum) over this CTA's rows; a thread owns 4 columns
                const int cx = tid % CW, ry = tid / CW;
                if (ry < RG) for (int jq = cx; jq < n2q; jq += CW) {
                    const int j = jq * 4;
                    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll 4
                    for (int i = ry; i < nrows; i += RG) {
                        const bool res = i < p.res_rows;
                        const float4 t = ld4((res ? slab : x) + (size_t)i * n2 + j);
                        const float rs = res ? 1.f : sc, fi = f[i];
                        m.x = fmaxf(m.x, fmaf(t.x, rs, -fi)); m.y = fmaxf(m.y, fmaf(t.y, rs, -fi));
                        m.z = fmaxf(m.z, fmaf(t.z, rs, -fi)); m.w = fmaxf(m.w, fmaf(t.w, rs, -fi));
                    }
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (nrows > ry) {
#pragma unroll 4
                        for (int i = ry; i < nrows; i += RG) {
                            const bool res = i < p.res_rows;
                            const float4 t = ld4((res ? slab : x) + (size_t)i * n2 + j);
                            const float rs = res ? 1.f : sc, fi = f[i];
                            sum.x += ex2(fmaf(t.x, rs, -fi) - m.x); sum.y += ex2(fmaf(t.y, rs, -fi) - m.y);
                            sum.z += ex2(fmaf(t.z, rs, -fi) - m.z); sum.w += ex2(fmaf(t.w, rs, -fi) - m.w);
                        }
                    }
                    *reinterpret_cast<float4 *>(pm + (size_t)ry * n2 + j) = m;
                    *reinterpret_cast<float4 *>(ps + (size_t)ry * n2 + j) = sum;
                }
                __syncthreads();
                for (int j = tid; j < n2; j += SKS_THREADS) {
                    float m = pm[j];
                    for (int q = 1; q < RG; ++q) m = fmaxf(m, pm[(size_t)q * n2 + j]);
                    float sum = 0.f;
                    if (m > -INFINITY)
                        for (int q = 0; q < RG; ++q) sum += ps[(size_t)q * n2 + j] * ex2(pm[(size_t)q * n2 + j] - m);
                    cm[j] = m; cs[j] = sum;
                }
                cluster.sync();                                    // partials of every CTA are published
                // CTA r reduces its slice of columns over the cluster and broadcasts g_j to every CTA
                for (int jj = tid; jj < cols_per_cta; jj += SKS_THREADS) {
                    const int j = r * cols_per_cta + jj;
                    if (j < n2) {
                        float m = -INFINITY;
                        float mq[16], sq[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (q < p.R) {
                                mq[q] = *cluster.map_shared_rank(cm + j, q);
                                sq[q] = *cluster.map_shared_rank(cs + j, q);
                                m = fmaxf(m, mq[q]);
                            }
                        float sum = 0.f;
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (q < p.R && mq[q] > -INFINITY) sum += sq[q] * ex2(mq[q] - m);
                        const float gj = m + lg2(sum);
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (q < p.R) *cluster.map_shared_rank(g + j, q) = gj;
                    }
                }
                cluster.sync();                                    // new g visible everywhere
            }
        }
        // ---- out = 2^(t - f - g)
        float *o = p.out + (size_t)b * p.n1 * n2 + (size_t)row0 * n2;
        for (int e = tid * 4; e < nrows * n2; e += SKS_THREADS * 4) {
            const int i = e / n2, j = e - i * n2;
            float4 v;
            if (i < p.res_rows) v = *reinterpret_cast<const float4 *>(slab + e);
            else { v = ld4(x + e); v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
            const float4 gv = ld4(g + j);
            const float fi = f[i];
            v.x = ex2(v.x - fi - gv.x); v.y = ex2(v.y - fi - gv.y); v.z = ex2(v.z - fi - gv.z); v.w = ex2(v.w - fi - gv.w);
            __stcs(reinterpret_cast<float4 *>(o + e), v);
        }
        __syncthreads();
    }
    cluster.sync();                                                // no CTA exits while peers may still touch its smem
}

static size_t sks_need(int n2, int rows_per_cta, int res_rows) {
    const int n2q = n2 / 4;
    const int CW = n2q < SKS_THREADS ? n2q : SKS_THREADS;
    const int RG = SKS_THREADS / CW;
    return ((size_t)res_rows * n2 + n2 + ((rows_per_cta + 3) & ~3) + 2 * (size_t)n2 + 2 * (size_t)RG * n2) * sizeof(float);
}

// Smallest cluster whose slabs hold the whole matrix (fewer CTAs per matrix = cheaper cluster barriers and more
// matrices in flight: N = 256 -> 2 CTAs, N = 512 -> 8); if even 8 slabs are too small (N = 1024) use 8 and re-read
// the rows that do not fit through L2.
static void sks_plan(int n1, int n2, SksParams &p, size_t &smem) {
    const size_t budget = 226 * 1024;
    for (int R = 1; R <= 8; R <<= 1) {
        const int rows = (n1 + R - 1) / R;
        if (sks_need(n2, rows, rows) <= budget || R == 8) {
            p.R = R; p.rows_per_cta = rows; p.res_rows = rows;
            while (p.res_rows > 0 && sks_need(n2, rows, p.res_rows) > budget) --p.res_rows;
            smem = sks_need(n2, rows, p.res_rows);
            return;
        }
    }
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
    cudaError_t e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    int max_clusters = 0;
    sks_plan(n1, n2, p, smem);
    if (smem > 227 * 1024) return TTDG_E_LIMIT;
    e = cudaFuncSetAttribute(sinkhorn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cfg.blockDim = dim3(SKS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.R; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(p.R);
    e = cudaOccupancyMaxActiveClusters(&max_clusters, sinkhorn_stream_kernel, &cfg);
    if (e != cudaSuccess || max_clusters < 1) { cudaGetLastError(); max_clusters = 148 / p.R; }
    p.s = s; p.out = out; p.batch = batch; p.n1 = n1; p.n2 = n2; p.max_iter = max_iter;
    p.inv_tau = 1.4426950408889634f / tau;     // log2(e) / tau
    const int ncl = batch < max_clusters ? batch : max_clusters;
    cfg.gridDim = dim3(ncl * p.R);
    ttdg::count_launches(1);
    e = cudaLaunchKernelEx(&cfg, sinkhorn_stream_kernel, p);
    return (int)e;
}
