// sinkhorn_small.cu - per-item log-space Sinkhorn (forward + backward) for the matching head.
// Reference: adapteacher/modeling/GModule/utils/sinkhorn.py:58-87 -> pygmtools.sinkhorn,
// call sites multi_graph_matching.py:467-468, 519-522 (pairwise, differentiable, dummy rows).
// One CTA per item; the working matrix (rows <= cols) lives in shared memory as fp64.
#include "sinkhorn_small.cuh"

namespace ttdg {

constexpr int SK_SMALL_MAX = 96;        // graphs have <= 95 nodes (19 per level x 5 levels, build_graph.py:189-195)
constexpr int SK_THREADS = 256;

struct SkItem { long long a_off, b_off, c_off; int n1, n2, lda, ldb, ldc, flags; };
constexpr int SK_ITEM_FIELDS = 9;

__device__ __forceinline__ SkItem load_item(const int64_t *items, int b) {
    const int64_t *d = items + (size_t)b * SK_ITEM_FIELDS;
    SkItem it;
    it.a_off = d[0]; it.b_off = d[1]; it.c_off = d[2];
    it.n1 = (int)d[3]; it.n2 = (int)d[4]; it.lda = (int)d[5]; it.ldb = (int)d[6]; it.ldc = (int)d[7]; it.flags = (int)d[8];
    return it;
}

__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_small_fwd_kernel(const float *__restrict__ s, float *__restrict__ out, const int64_t *__restrict__ items,
                          double tau, int max_iter, int dummy_row, int max_dim) {
    extern __shared__ double smem[];
    const SkItem it = load_item(items, blockIdx.x);
    const int n1 = it.n1, n2 = it.n2;
    if (n1 <= 0 || n2 <= 0 || n1 > max_dim || n2 > max_dim) return;
    const bool tr = n2 < n1 || (n1 == n2 && (it.flags & 1));
    const int nr = tr ? n2 : n1, nq = tr ? n1 : n2;
    const int pitch = nq | 1;
    double *z = smem;
    double *padv = z + (size_t)nr * pitch;
    const int mult = dummy_row ? nq - nr : 0;
    const float *src = s + it.a_off;
    for (int e = threadIdx.x; e < n1 * n2; e += SK_THREADS) {
        const int i = e / n2, j = e - i * n2;
        const double v = (double)src[(size_t)i * it.lda + j] / tau;
        if (tr) z[j * pitch + i] = v; else z[i * pitch + j] = v;
    }
    for (int q = threadIdx.x; q < nq; q += SK_THREADS) padv[q] = -100.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = 0; k < max_iter; ++k) {
        sinkhorn_step(z, pitch, 1, nr, nq, padv, mult, k, nullptr, warp, SK_THREADS / 32, lane);
        __syncthreads();
    }
    float *dst = out + it.b_off;
    for (int e = threadIdx.x; e < n1 * n2; e += SK_THREADS) {
        const int i = e / n2, j = e - i * n2;
        dst[(size_t)i * it.ldb + j] = (float)exp(tr ? z[j * pitch + i] : z[i * pitch + j]);
    }
    if (it.c_off >= 0) {
        float *mir = out + it.c_off;
        for (int e = threadIdx.x; e < n1 * n2; e += SK_THREADS) {
            const int j = e / n1, i = e - j * n1;
            mir[(size_t)j * it.ldc + i] = (float)exp(tr ? z[j * pitch + i] : z[i * pitch + j]);
        }
    }
}

// items: {s_off, gout_off, gin_off, n1, n2, ld_s, ld_gout, ld_gin}
__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_small_bwd_kernel(const float *__restrict__ s, const float *__restrict__ gout, float *__restrict__ gin,
                          const int64_t *__restrict__ items, double tau, int max_iter, int dummy_row, int max_dim) {
    extern __shared__ double smem[];
    const SkItem it = load_item(items, blockIdx.x);
    const int n1 = it.n1, n2 = it.n2;
    if (n1 <= 0 || n2 <= 0 || n1 > max_dim || n2 > max_dim) return;
    const bool tr = n2 < n1 || (n1 == n2 && (it.flags & 1));
    const int nr = tr ? n2 : n1, nq = tr ? n1 : n2;
    const int pitch = nq | 1;
    double *z = smem;
    double *g = z + (size_t)nr * pitch;
    double *padv = g + (size_t)nr * pitch;
    double *gp = padv + nq;
    double *L = gp + nq;                       // max_iter x (nq + 1)
    const int mult = dummy_row ? nq - nr : 0;
    const float *src = s + it.a_off;
    for (int e = threadIdx.x; e < n1 * n2; e += SK_THREADS) {
        const int i = e / n2, j = e - i * n2;
        const double v = (double)src[(size_t)i * it.lda + j] / tau;
        if (tr) z[j * pitch + i] = v; else z[i * pitch + j] = v;
    }
    for (int q = threadIdx.x; q < nq; q += SK_THREADS) { padv[q] = -100.0; gp[q] = 0.0; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = 0; k < max_iter; ++k) {
        sinkhorn_step(z, pitch, 1, nr, nq, padv, mult, k, L + (size_t)k * (nq + 1), warp, SK_THREADS / 32, lane);
        __syncthreads();
    }
    const float *go = gout + it.b_off;
    for (int e = threadIdx.x; e < n1 * n2; e += SK_THREADS) {
        const int i = e / n2, j = e - i * n2;
        const int w = tr ? j * pitch + i : i * pitch + j;
        g[w] = (double)go[(size_t)i * it.ldb + j] * exp(z[w]);      // out = exp(z)
    }
    __syncthreads();
    for (int k = max_iter - 1; k >= 0; --k) {
        sinkhorn_step_bwd(z, g, pitch, 1, nr, nq, padv, gp, mult, k, L + (size_t)k * (nq + 1), warp,
                          SK_THREADS / 32, lane);
        __syncthreads();
    }
    float *gi = gin + it.c_off;
    for (int e = threadIdx.x; e < n1 * n2; e += SK_THREADS) {
        const int i = e / n2, j = e - i * n2;
        gi[(size_t)i * it.ldc + j] = (float)(g[tr ? j * pitch + i : i * pitch + j] / tau);
    }
}

static size_t fwd_smem(int dim) { return ((size_t)dim * (dim | 1) + dim) * sizeof(double); }
static size_t bwd_smem(int dim, int max_iter) {
    return ((size_t)2 * dim * (dim | 1) + 2 * dim + (size_t)max_iter * (dim + 1)) * sizeof(double);
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_sinkhorn_small_fwd(const float *s, float *out, const int64_t *items, int n_items, int max_dim,
                                       double tau, int max_iter, int dummy_row, void *stream) {
    TTDG_CHECK_ARG(s && out && items && n_items >= 0 && tau > 0 && max_iter >= 0 && max_dim >= 1);
    if (max_dim > SK_SMALL_MAX) return TTDG_E_LIMIT;
    if (n_items == 0) return 0;
    const size_t smem = fwd_smem(max_dim);
    cudaError_t e = cudaFuncSetAttribute(sinkhorn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem(SK_SMALL_MAX));
    if (e != cudaSuccess) return (int)e;
    ttdg::count_launches(1);
    sinkhorn_small_fwd_kernel<<<n_items, SK_THREADS, smem, (cudaStream_t)stream>>>(s, out, items, tau, max_iter, dummy_row, max_dim);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_sinkhorn_small_bwd(const float *s, const float *grad_out, float *grad_in, const int64_t *items,
                                       int n_items, int max_dim, double tau, int max_iter, int dummy_row, void *stream) {
    TTDG_CHECK_ARG(s && grad_out && grad_in && items && n_items >= 0 && tau > 0 && max_iter >= 0 && max_dim >= 1);
    if (max_dim > SK_SMALL_MAX) return TTDG_E_LIMIT;
    if (n_items == 0) return 0;
    const size_t smem = bwd_smem(max_dim, max_iter);
    if (smem > 227 * 1024) return TTDG_E_LIMIT;
    cudaError_t e = cudaFuncSetAttribute(sinkhorn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ttdg::count_launches(1);
    sinkhorn_small_bwd_kernel<<<n_items, SK_THREADS, smem, (cudaStream_t)stream>>>(s, grad_out, grad_in, items, tau,
                                                                                  max_iter, dummy_row, max_dim);
    TTDG_LAUNCH_RET();
}
