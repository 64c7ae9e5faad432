// sampler.cu - graph-node sampler over the FPN pyramid.
// Reference: PrototypeComputation (adapteacher/modeling/GModule/build_graph.py): locations :133-157
// ((x, y) = stride * idx + stride // 2 for strides 4, 8, 16, 32, 64), per-location target assignment :70-115
// (inside the box, max(l, t, r, b) within the level's size range :28-39, smallest-area box wins, label =
// class + 1, 0 = background), per-level sub-sampling :160-250 (keep every step-th positive, step = P // 10,
// all of them if step <= 1).  All box arithmetic is fp32 like the reference's tensors.
#include "common.cuh"

namespace ttdg {

struct Pyramid {
    int h[5], w[5], base[6];     // base[l] = first flat location index of level l, base[5] = L
};

__device__ __forceinline__ void level_range(int l, float &lo, float &hi) {
    // build_graph.py:28-39 (INF = 1e8)
    const float los[5] = {-1.f, 64.f, 128.f, 256.f, 512.f};
    const float his[5] = {64.f, 128.f, 256.f, 512.f, 100000000.f};
    lo = los[l]; hi = his[l];
}

// grid: (ceil(L / 256), B)
__global__ void __launch_bounds__(256)
sampler_label_kernel(const float *__restrict__ boxes, const int64_t *__restrict__ classes,
                     const int32_t *__restrict__ box_off, Pyramid pyr, int32_t *__restrict__ label) {
    const int loc = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
    const int L = pyr.base[5];
    if (loc >= L) return;
    int l = 0;
    while (loc >= pyr.base[l + 1]) ++l;
    const int stride = 4 << l;
    const int pix = loc - pyr.base[l];
    const int iy = pix / pyr.w[l], ix = pix - iy * pyr.w[l];
    const float x = (float)(ix * stride) + (float)(stride / 2);
    const float y = (float)(iy * stride) + (float)(stride / 2);
    float lo, hi;
    level_range(l, lo, hi);
    const float INF = 100000000.f;
    float best = INF;
    int best_k = -1;
    for (int k = box_off[b]; k < box_off[b + 1]; ++k) {
        const float x0 = boxes[4 * k], y0 = boxes[4 * k + 1], x1 = boxes[4 * k + 2], y1 = boxes[4 * k + 3];
        const float dl = x - x0, dt = y - y0, dr = x1 - x, db = y1 - y;
        const float mn = fminf(fminf(dl, dt), fminf(dr, db));
        const float mx = fmaxf(fmaxf(dl, dt), fmaxf(dr, db));
        const bool ok = (mn > 0.f) && (mx >= lo) && (mx <= hi);
        const float area = __fmul_rn(__fadd_rn(__fsub_rn(x1, x0), 1.f), __fadd_rn(__fsub_rn(y1, y0), 1.f));
        const float a = ok ? area : INF;
        if (a < best) { best = a; best_k = k; }          // first minimum wins (torch.min)
    }
    label[(size_t)b * L + loc] = (best_k >= 0 && best != INF) ? (int32_t)(classes[best_k] + 1) : 0;
}

// grid: (5, B), 1024 threads.  Ranks the positives of one level in location order and keeps every step-th.
__global__ void __launch_bounds__(1024)
sampler_select_kernel(const int32_t *__restrict__ label, Pyramid pyr, int sample_dist, int max_per_level,
                      int32_t *__restrict__ counts, int32_t *__restrict__ sel_idx) {
    const int l = blockIdx.x, b = blockIdx.y;
    const int L = pyr.base[5], n = pyr.base[l + 1] - pyr.base[l];
    const int32_t *lab = label + (size_t)b * L + pyr.base[l];
    __shared__ int wsum[32];
    __shared__ int total_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pass 1: count positives
    int cnt = 0;
    for (int e = threadIdx.x; e < n; e += 1024) cnt += lab[e] > 0;
    cnt = warp_sum_i(cnt);
    if (lane == 0) wsum[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += wsum[w]; total_s = t; }
    __syncthreads();
    const int P = total_s;
    const int step = P / sample_dist;
    int32_t *out = sel_idx + (size_t)(b * 5 + l) * max_per_level;
    // pass 2: rank in order (chunks of 1024 locations)
    int running = 0;
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int e = c0 + threadIdx.x;
        const bool pos = e < n && lab[e] > 0;
        const unsigned m = __ballot_sync(TTDG_FULL, pos);
        __syncthreads();
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int before = 0, chunk = 0;
        for (int w = 0; w < 32; ++w) { const int v = wsum[w]; if (w < warp) before += v; chunk += v; }
        if (pos) {
            const int rank = running + before + __popc(m & ((1u << lane) - 1u));
            int slot = -1;
            if (step > 1) { if (rank % step == 0) slot = rank / step; } else slot = rank;
            if (slot >= 0 && slot < max_per_level) out[slot] = pyr.base[l] + e;
        }
        running += chunk;
    }
    if (threadIdx.x == 0) {
        int kept = step > 1 ? (P + step - 1) / step : P;
        counts[b * 5 + l] = kept < max_per_level ? kept : max_per_level;
    }
}

struct FeatPtrs { const float *p[5]; long long s_img[5], s_ch[5], s_px[5]; };
struct FeatPtrsW { float *p[5]; long long s_img[5], s_ch[5], s_px[5]; };

// grid: n_total nodes; block: 256 threads over channels
template <bool BWD>
__global__ void __launch_bounds__(256)
sampler_gather_kernel(FeatPtrsW f, Pyramid pyr, int B, int C, const int32_t *__restrict__ label,
                      const int32_t *__restrict__ sel_idx, int max_per_level, const int32_t *__restrict__ node_off,
                      float *__restrict__ nodes, int64_t *__restrict__ labels_out) {
    const int k = blockIdx.x;
    int seg = 0;                                           // seg = b * 5 + l
    const int nseg = B * 5;
    while (seg + 1 < nseg && node_off[seg + 1] <= k) ++seg;
    const int b = seg / 5, l = seg - 5 * b, slot = k - node_off[seg];
    const int loc = sel_idx[(size_t)seg * max_per_level + slot];
    const int pix = loc - pyr.base[l];
    float *base = f.p[l] + (size_t)b * f.s_img[l] + (size_t)pix * f.s_px[l];
    for (int c = threadIdx.x; c < C; c += 256) {
        if (BWD) base[(size_t)c * f.s_ch[l]] += nodes[(size_t)k * C + c];
        else nodes[(size_t)k * C + c] = base[(size_t)c * f.s_ch[l]];
    }
    if (!BWD && threadIdx.x == 0) labels_out[k] = (int64_t)label[(size_t)b * pyr.base[5] + loc];
}

static int make_pyramid(const int32_t *hw, Pyramid &p) {
    p.base[0] = 0;
    for (int l = 0; l < 5; ++l) {
        p.h[l] = hw[2 * l]; p.w[l] = hw[2 * l + 1];
        if (p.h[l] < 1 || p.w[l] < 1) return TTDG_E_ARG;
        p.base[l + 1] = p.base[l] + p.h[l] * p.w[l];
    }
    return 0;
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_sampler_select(const float *boxes, const int64_t *classes, const int32_t *box_off, int B,
                                   const int32_t *lvl_hw_h, int sample_dist, int max_per_level, int32_t *label,
                                   int32_t *counts, int32_t *sel_idx, void *stream) {
    TTDG_CHECK_ARG(boxes && classes && box_off && lvl_hw_h && label && counts && sel_idx && B >= 1 && sample_dist >= 1 &&
                   max_per_level >= 1);
    Pyramid pyr;
    if (make_pyramid(lvl_hw_h, pyr)) return TTDG_E_ARG;
    dim3 g1(ceil_div(pyr.base[5], 256), B);
    ttdg::count_launches(2);
    sampler_label_kernel<<<g1, 256, 0, (cudaStream_t)stream>>>(boxes, classes, box_off, pyr, label);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    dim3 g2(5, B);
    sampler_select_kernel<<<g2, 1024, 0, (cudaStream_t)stream>>>(label, pyr, sample_dist, max_per_level, counts, sel_idx);
    TTDG_LAUNCH_RET();
}

static int fill_feat(const void *const *ptrs_h, const int64_t *strides_h, FeatPtrsW &f) {
    for (int l = 0; l < 5; ++l) {
        if (!ptrs_h[l]) return TTDG_E_ARG;
        f.p[l] = (float *)ptrs_h[l];
        f.s_img[l] = strides_h[3 * l]; f.s_ch[l] = strides_h[3 * l + 1]; f.s_px[l] = strides_h[3 * l + 2];
    }
    return 0;
}

extern "C" int ttdg_sampler_gather(const float *const *feat_ptrs_h, const int64_t *feat_strides_h, const int32_t *lvl_hw_h,
                                   int B, int C, int n_total, const int32_t *label, const int32_t *sel_idx,
                                   int max_per_level, const int32_t *node_off, float *nodes, int64_t *labels_out,
                                   void *stream) {
    TTDG_CHECK_ARG(feat_ptrs_h && feat_strides_h && lvl_hw_h && label && sel_idx && node_off && nodes && labels_out);
    if (n_total == 0) return 0;
    Pyramid pyr;
    FeatPtrsW f;
    if (make_pyramid(lvl_hw_h, pyr) || fill_feat((const void *const *)feat_ptrs_h, feat_strides_h, f)) return TTDG_E_ARG;
    ttdg::count_launches(1);
    sampler_gather_kernel<false><<<n_total, 256, 0, (cudaStream_t)stream>>>(f, pyr, B, C, label, sel_idx, max_per_level,
                                                                           node_off, nodes, labels_out);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_sampler_scatter_bwd(const float *grad_nodes, float *const *gfeat_ptrs_h, const int64_t *feat_strides_h,
                                        const int32_t *lvl_hw_h, int B, int C, int n_total, const int32_t *sel_idx,
                                        int max_per_level, const int32_t *node_off, void *stream) {
    TTDG_CHECK_ARG(grad_nodes && gfeat_ptrs_h && feat_strides_h && lvl_hw_h && sel_idx && node_off);
    if (n_total == 0) return 0;
    Pyramid pyr;
    FeatPtrsW f;
    if (make_pyramid(lvl_hw_h, pyr) || fill_feat((const void *const *)gfeat_ptrs_h, feat_strides_h, f)) return TTDG_E_ARG;
    ttdg::count_launches(1);
    sampler_gather_kernel<true><<<n_total, 256, 0, (cudaStream_t)stream>>>(f, pyr, B, C, nullptr, sel_idx, max_per_level,
                                                                          node_off, const_cast<float *>(grad_nodes), nullptr);
    TTDG_LAUNCH_RET();
}
