// mgm_ops.cu - attention adjacency, separable learned affinity (fwd/bwd) and the matching loss (fwd/bwd).
// Reference: multi_graph_matching.py:487-569 (MGM3_unsup.forward) and the modules it calls.
#include "common.cuh"

namespace ttdg {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG) for production-mode dropout; parity tests inject explicit masks.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t ctr) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c0, c1, c2, c3, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return (float)(c0 >> 8) * (1.0f / 16777216.0f);      // [0, 1)
}

// ------------------------------------------------------------------------------------------------
// Attention adjacency (utils/attentions.py:25-42, 60-86; mgm:497-502).  One warp per row of A.
// S = q k^T (M x M fp32, from ttdg_gemm_f64acc); only the diagonal blocks are read.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_adjacency_kernel(const float *__restrict__ S, const int32_t *__restrict__ node_off, int G, int M, float scale,
                      const float *__restrict__ keep_mask, const int64_t *__restrict__ mask_off, float p_drop,
                      uint64_t seed, uint64_t offset, float *__restrict__ A) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    int g = 0;
    while (g + 1 < G && node_off[g + 1] <= row) ++g;
    const int c0 = node_off[g], c1 = node_off[g + 1], n = c1 - c0, i = row - c0;
    const float *srow = S + (size_t)row * M;
    float *arow = A + (size_t)row * M;
    double mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmax(mx, (double)(srow[c0 + j] * scale));
    mx = warp_max(mx);
    double sum = 0.0;
    for (int j = lane; j < n; j += 32) sum += exp((double)(srow[c0 + j] * scale) - mx);
    sum = warp_sum(sum);
    const float keep_scale = 1.0f / (1.0f - p_drop);
    for (int j = lane; j < M; j += 32) {
        float v = 0.0f;
        if (j >= c0 && j < c1 && j != row) {
            const int jj = j - c0;
            v = (float)(exp((double)(srow[j] * scale) - mx) / sum);
            if (keep_mask) v = v * (keep_mask[mask_off[g] + (size_t)i * n + jj] * keep_scale);
            else if (p_drop > 0.0f)
                v = philox_uniform(seed, offset + (uint64_t)row * (uint64_t)M + (uint64_t)j) >= p_drop ? v * keep_scale : 0.0f;
        }
        arow[j] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Affinity, separable form (utils/affinity.py:44-57).
//   hidden stage: ac[node] = [ Xp W0a^T | Yp W0b^T + b0 ]  (fp64, 2*hidden per node) - via gemm + this bias add
//   pair stage:   out[i, j] = sum_k w1[k] relu(a[src_i, k] + c[tgt_j, k]) + b1
// ------------------------------------------------------------------------------------------------
__global__ void affinity_bias_kernel(double *__restrict__ ac, const float *__restrict__ b0, int M, int hidden) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)M * hidden) return;
    const size_t node = e / hidden;
    const int k = (int)(e - node * hidden);
    ac[node * 2 * hidden + hidden + k] += (double)b0[k];
}

constexpr int AFF_TI = 8;      // src rows per CTA
// grid: (ceil(max_n_src / AFF_TI), n_pairs); block 256 threads = 8 warps; warp w handles src row i0 + w, lanes over j
__global__ void __launch_bounds__(256)
affinity_pairs_fwd_kernel(const double *__restrict__ ac, const float *__restrict__ w1, const float *__restrict__ b1,
                          const int64_t *__restrict__ pairs, const int64_t *__restrict__ out_off, int hidden,
                          float *__restrict__ out) {
    extern __shared__ double sm[];
    double *a_s = sm;                       // AFF_TI x hidden
    double *w_s = a_s + AFF_TI * hidden;    // hidden
    const int64_t *pd = pairs + (size_t)blockIdx.y * 4;
    const int src0 = (int)pd[0], ns = (int)pd[1], tgt0 = (int)pd[2], nt = (int)pd[3];
    const int i0 = blockIdx.x * AFF_TI;
    if (i0 >= ns) return;
    const int rows = min(AFF_TI, ns - i0);
    for (int e = threadIdx.x; e < rows * hidden; e += 256) {
        const int r = e / hidden, k = e - r * hidden;
        a_s[e] = ac[(size_t)(src0 + i0 + r) * 2 * hidden + k];
    }
    for (int k = threadIdx.x; k < hidden; k += 256) w_s[k] = (double)w1[k];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w >= rows) return;
    const double *a = a_s + w * hidden;
    float *orow = out + out_off[blockIdx.y] + (size_t)(i0 + w) * nt;
    const double bias = (double)b1[0];
    for (int j = lane; j < nt; j += 32) {
        const double *c = ac + (size_t)(tgt0 + j) * 2 * hidden + hidden;
        double acc0 = 0.0, acc1 = 0.0;
        for (int k = 0; k < hidden; k += 2) {
            const double2 cv = *reinterpret_cast<const double2 *>(c + k);
            acc0 = fma(w_s[k], fmax(a[k] + cv.x, 0.0), acc0);
            acc1 = fma(w_s[k + 1], fmax(a[k + 1] + cv.y, 0.0), acc1);
        }
        orow[j] = (float)(acc0 + acc1 + bias);
    }
}

// Backward of the pair stage, deterministic gather form.
//   g_a[node i, k] = w1[k] * sum over pairs p with src graph containing i, over j: gout_p[i, j] * [a_ik + c_jk > 0]
//   g_c[node j, k] = w1[k] * sum over pairs p with tgt graph containing j, over i: gout_p[i, j] * [a_ik + c_jk > 0]
//   g_w1[k]        = sum_p sum_ij gout_p[i, j] relu(a_ik + c_jk)      (partials per CTA -> scratch, then reduced)
// grid: (M nodes, 2 roles); block = hidden/2 threads, each thread owns 2 consecutive k.
__global__ void __launch_bounds__(256)
affinity_pairs_bwd_kernel(const double *__restrict__ ac, const float *__restrict__ w1, const int64_t *__restrict__ pairs,
                          const int64_t *__restrict__ out_off, int n_pairs, int hidden, const float *__restrict__ gout,
                          double *__restrict__ g_ac, double *__restrict__ w1_partial) {
    const int node = blockIdx.x, role = blockIdx.y;       // role 0: node acts as src row (a), 1: as tgt column (c)
    extern __shared__ float go_s[];                        // one row / column of gout for the current pair
    double acc[4] = {0, 0, 0, 0}, wacc[4] = {0, 0, 0, 0};
    const int kpt = hidden / blockDim.x;                   // 2 (hidden 512, 256 threads)
    const int k0 = threadIdx.x * kpt;
    double mine[4];
    for (int t = 0; t < kpt; ++t) mine[t] = ac[(size_t)node * 2 * hidden + role * hidden + k0 + t];
    for (int p = 0; p < n_pairs; ++p) {
        const int64_t *pd = pairs + (size_t)p * 4;
        const int src0 = (int)pd[0], ns = (int)pd[1], tgt0 = (int)pd[2], nt = (int)pd[3];
        const int my0 = role == 0 ? src0 : tgt0, myn = role == 0 ? ns : nt;
        if (node < my0 || node >= my0 + myn) continue;     // block-uniform
        const int idx = node - my0;
        const int oth0 = role == 0 ? tgt0 : src0, othn = role == 0 ? nt : ns;
        const float *gp = gout + out_off[p];
        __syncthreads();
        for (int t = threadIdx.x; t < othn; t += blockDim.x)
            go_s[t] = role == 0 ? gp[(size_t)idx * nt + t] : gp[(size_t)t * nt + idx];
        __syncthreads();
        for (int t = 0; t < othn; ++t) {
            const double gv = (double)go_s[t];
            if (gv == 0.0) continue;                       // self-pair blocks carry no gradient (mgm:543-564)
            const double *o = ac + (size_t)(oth0 + t) * 2 * hidden + (1 - role) * hidden + k0;
            for (int u = 0; u < kpt; ++u) {
                const double h = mine[u] + o[u];
                if (h > 0.0) { acc[u] += gv; if (role == 0) wacc[u] = fma(gv, h, wacc[u]); }
            }
        }
    }
    for (int u = 0; u < kpt; ++u) {
        g_ac[(size_t)node * 2 * hidden + role * hidden + k0 + u] = acc[u] * (double)w1[k0 + u];
        if (role == 0) w1_partial[(size_t)node * hidden + k0 + u] = wacc[u];
    }
}

// g_w1[k] = sum_node w1_partial[node][k];  g_b1 = sum of all gout over all pairs
__global__ void __launch_bounds__(256)
affinity_reduce_kernel(const double *__restrict__ w1_partial, int M, int hidden, const float *__restrict__ gout,
                       int64_t gout_total, float *__restrict__ g_w1, float *__restrict__ g_b1) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x < gridDim.x - 1) {
        if (k < hidden) {
            double s = 0.0;
            for (int nd = 0; nd < M; ++nd) s += w1_partial[(size_t)nd * hidden + k];
            g_w1[k] = (float)s;
        }
    } else {                                                // last CTA: bias gradient
        __shared__ double red[256];
        double s = 0.0;
        for (int64_t e = threadIdx.x; e < gout_total; e += 256) s += (double)gout[e];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) g_b1[0] = (float)red[0];
    }
}

// ------------------------------------------------------------------------------------------------
// Matching loss (mgm:543-564, 594-633; utils/losses.py:83-103, 419-455).
// pair (i1 < i2): S'(b, a) = Wds[off[i2] + b, off[i1] + a]  (block written by the pairwise Sinkhorn, src = i2, tgt = i1),
// target y(a, b) = sum_u U[off[i1] + a, u] U[off[i2] + b, u];  focal BCE, mean over the block, mean over pairs.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double focal_term(double p, double y) {
    const double eps = (double)1e-6f;           // torch.clamp(min=1e-6, max=1 - 1e-6) on an fp32 tensor
    const double hi = (double)0.999999f;        // python 1 - 1e-6 cast to fp32
    p = fmin(fmax(p, eps), hi);
    return -0.25 * (1.0 - p) * (1.0 - p) * y * log(p) - 0.75 * p * p * (1.0 - y) * log(1.0 - p);
}
__device__ __forceinline__ double focal_grad(double p, double y) {
    const double eps = (double)1e-6f, hi = (double)0.999999f;
    if (p < eps || p > hi) return 0.0;          // clamp has zero gradient outside (torch: inclusive bounds pass grad)
    const double q = 1.0 - p;
    // d/dp [ -a q^2 y log p - (1-a) p^2 (1-y) log q ]
    return -0.25 * y * (-2.0 * q * log(p) + q * q / p) - 0.75 * (1.0 - y) * (2.0 * p * log(q) - p * p / q);
}

__device__ __forceinline__ void pair_from_index(int p, int G, int &i1, int &i2) {
    int c = 0;
    for (i1 = 0; i1 < G - 1; ++i1) {
        const int cnt = G - 1 - i1;
        if (p < c + cnt) { i2 = i1 + 1 + (p - c); return; }
        c += cnt;
    }
    i1 = 0; i2 = 1;
}

// grid: n_pairs CTAs; writes pair_loss[p] (fp64); a second tiny kernel averages (deterministic)
template <bool BWD>
__global__ void __launch_bounds__(256)
matching_loss_kernel(const float *__restrict__ Wds, const float *__restrict__ U, const int32_t *__restrict__ node_off,
                     int G, int M, int n_univ, double *__restrict__ pair_loss, int32_t *__restrict__ flags,
                     const float *__restrict__ grad_loss, float *__restrict__ grad_Wds) {
    int i1, i2;
    pair_from_index(blockIdx.x, G, i1, i2);
    const int o1 = node_off[i1], n1 = node_off[i1 + 1] - o1, o2 = node_off[i2], n2 = node_off[i2 + 1] - o2;
    const int npairs = G * (G - 1) / 2;
    __shared__ double red[256];
    double s = 0.0;
    const double gscale = BWD ? (double)grad_loss[0] / ((double)npairs * (double)n1 * (double)n2) : 0.0;
    bool bad = false;
    for (int e = threadIdx.x; e < n1 * n2; e += 256) {
        const int b = e / n1, a = e - b * n1;                  // a fastest: coalesced over Wds columns
        const size_t widx = (size_t)(o2 + b) * M + (o1 + a);
        const double p = (double)Wds[widx];
        const float *ua = U + (size_t)(o1 + a) * n_univ, *ub = U + (size_t)(o2 + b) * n_univ;
        double y = 0.0;
        for (int u = 0; u < n_univ; ++u) y = fma((double)ua[u], (double)ub[u], y);
        if (BWD) grad_Wds[widx] = (float)(gscale * focal_grad(p, y));
        else {
            if (!(p >= 0.0 && p <= 1.0) || !(y >= 0.0 && y <= 1.0)) bad = true;
            s += focal_term(p, y);
        }
    }
    if (!BWD) {
        if (bad) atomicOr(flags, 1);
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) pair_loss[blockIdx.x] = red[0] / ((double)n1 * (double)n2);
    }
}

__global__ void matching_loss_finish_kernel(const double *__restrict__ pair_loss, int npairs, float *__restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        // the reference accumulates fp32 pair losses in order and divides by the count (mgm:560-564)
        float acc = 0.0f;
        for (int p = 0; p < npairs; ++p) acc += (float)pair_loss[p];
        loss[0] = acc / (float)npairs;
    }
}

__global__ void __launch_bounds__(256)
focal_bce_fwd_kernel(const float *__restrict__ p, const float *__restrict__ y, int64_t n, double *__restrict__ partial) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256)
        s += focal_term((double)p[e], (double)y[e]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void focal_bce_finish_kernel(const double *__restrict__ partial, int nparts, int64_t n, float *__restrict__ loss) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nparts; ++i) s += partial[i];
        loss[0] = (float)(s / (double)n);
    }
}
__global__ void __launch_bounds__(256)
focal_bce_bwd_kernel(const float *__restrict__ p, const float *__restrict__ y, int64_t n, const float *__restrict__ grad_loss,
                     float *__restrict__ grad_p) {
    const double g = (double)grad_loss[0] / (double)n;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256)
        grad_p[e] = (float)(g * focal_grad((double)p[e], (double)y[e]));
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_attn_adjacency(const float *S, const int32_t *node_off, int G, int M, float scale,
                                   const float *keep_mask, const int64_t *mask_off, float p_drop, uint64_t seed,
                                   uint64_t offset, float *A, void *stream) {
    TTDG_CHECK_ARG(S && node_off && A && G >= 1 && M >= 0 && p_drop >= 0.0f && p_drop < 1.0f);
    TTDG_CHECK_ARG(!keep_mask || mask_off);
    if (M == 0) return 0;
    ttdg::count_launches(1);
    attn_adjacency_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(S, node_off, G, M, scale, keep_mask, mask_off,
                                                                           p_drop, seed, offset, A);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_affinity_hidden(const float *Xp, const float *Yp, const float *w0, const float *b0, int M, int dim,
                                    int hidden, double *ac, void *stream) {
    TTDG_CHECK_ARG(Xp && Yp && w0 && b0 && ac && M >= 0 && dim > 0 && hidden > 0);
    if (M == 0) return 0;
    // a = Xp W0[:, :dim]^T   -> ac[:, 0:hidden];   c = Yp W0[:, dim:]^T (+ b0) -> ac[:, hidden:2 hidden]
    int rc = ttdg_gemm_f64acc(0, 1, M, hidden, dim, Xp, 0, dim, w0, 0, 2 * dim, ac, 1, 2 * hidden, 0, stream);
    if (rc) return rc;
    rc = ttdg_gemm_f64acc(0, 1, M, hidden, dim, Yp, 0, dim, w0 + dim, 0, 2 * dim, ac + hidden, 1, 2 * hidden, 0, stream);
    if (rc) return rc;
    const size_t n = (size_t)M * hidden;
    ttdg::count_launches(1);
    affinity_bias_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ac, b0, M, hidden);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_affinity_pairs_fwd(const double *ac, const float *w1, const float *b1, const int64_t *pairs,
                                       const int64_t *out_off, int n_pairs, int max_n_src, int hidden, float *out,
                                       void *stream) {
    TTDG_CHECK_ARG(ac && w1 && b1 && pairs && out_off && out && n_pairs >= 0 && hidden > 0 && hidden % 2 == 0);
    if (n_pairs == 0 || max_n_src == 0) return 0;
    const size_t smem = (size_t)(AFF_TI + 1) * hidden * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(affinity_pairs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(ceil_div(max_n_src, AFF_TI), n_pairs);
    ttdg::count_launches(1);
    affinity_pairs_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(ac, w1, b1, pairs, out_off, hidden, out);
    TTDG_LAUNCH_RET();
}

extern "C" int64_t ttdg_affinity_bwd_scratch_bytes(int M, int hidden) {
    return (int64_t)M * hidden * (int64_t)sizeof(double);
}

extern "C" int ttdg_affinity_pairs_bwd(const double *ac, const float *w1, const int64_t *pairs, const int64_t *out_off,
                                       int n_pairs, int hidden, int M, int max_n, const float *grad_out,
                                       int64_t grad_out_total, double *g_ac, float *g_w1, float *g_b1, void *scratch,
                                       void *stream) {
    TTDG_CHECK_ARG(ac && w1 && pairs && out_off && grad_out && g_ac && g_w1 && g_b1 && scratch);
    TTDG_CHECK_ARG(hidden % 256 == 0 && hidden / 256 <= 4 && M >= 0 && max_n >= 0);
    if (M == 0) return 0;
    dim3 grid(M, 2);
    ttdg::count_launches(2);
    affinity_pairs_bwd_kernel<<<grid, 256, (size_t)max_n * sizeof(float), (cudaStream_t)stream>>>(
        ac, w1, pairs, out_off, n_pairs, hidden, grad_out, g_ac, reinterpret_cast<double *>(scratch));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    affinity_reduce_kernel<<<ceil_div(hidden, 256) + 1, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double *>(scratch), M, hidden, grad_out, grad_out_total, g_w1, g_b1);
    TTDG_LAUNCH_RET();
}

extern "C" int64_t ttdg_matching_loss_scratch_bytes(int G) { return (int64_t)(G * (G - 1) / 2 + 1) * 8; }

extern "C" int ttdg_matching_loss_fwd(const float *Wds, const float *U, const int32_t *node_off, int G, int M, int n_univ,
                                      float *loss, int32_t *flags, void *scratch, void *stream) {
    TTDG_CHECK_ARG(Wds && U && node_off && loss && flags && scratch && G >= 2 && M >= 0 && n_univ > 0);
    const int npairs = G * (G - 1) / 2;
    ttdg::count_launches(2);
    matching_loss_kernel<false><<<npairs, 256, 0, (cudaStream_t)stream>>>(Wds, U, node_off, G, M, n_univ,
                                                                         reinterpret_cast<double *>(scratch), flags, nullptr, nullptr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    matching_loss_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double *>(scratch), npairs, loss);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_matching_loss_bwd(const float *Wds, const float *U, const int32_t *node_off, int G, int M, int n_univ,
                                      const float *grad_loss, float *grad_Wds, void *stream) {
    TTDG_CHECK_ARG(Wds && U && node_off && grad_loss && grad_Wds && G >= 2 && M >= 0 && n_univ > 0);
    cudaError_t e = cudaMemsetAsync(grad_Wds, 0, (size_t)M * M * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    ttdg::count_launches(1);
    matching_loss_kernel<true><<<G * (G - 1) / 2, 256, 0, (cudaStream_t)stream>>>(Wds, U, node_off, G, M, n_univ, nullptr,
                                                                                 nullptr, grad_loss, grad_Wds);
    TTDG_LAUNCH_RET();
}

extern "C" int64_t ttdg_focal_bce_scratch_bytes(void) { return 1024 * 8; }

extern "C" int ttdg_focal_bce_fwd(const float *p, const float *y, int64_t n, float *loss, void *scratch, void *stream) {
    TTDG_CHECK_ARG(p && y && loss && scratch && n > 0);
    const int nb = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);      // 4 CTAs per SM x 148
    ttdg::count_launches(2);
    focal_bce_fwd_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(p, y, n, reinterpret_cast<double *>(scratch));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    focal_bce_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double *>(scratch), nb, n, loss);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_focal_bce_bwd(const float *p, const float *y, int64_t n, const float *grad_loss, float *grad_p,
                                  void *stream) {
    TTDG_CHECK_ARG(p && y && grad_loss && grad_p && n > 0);
    const int nb = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
    ttdg::count_launches(1);
    focal_bce_bwd_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(p, y, n, grad_loss, grad_p);
    TTDG_LAUNCH_RET();
}
