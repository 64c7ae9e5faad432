// evalmetric.cu - on-device half of DiceEvaluator (adapteacher/evaluation/dice_metric.py:25-92, 110-240).
//
// The reference copies every predicted mask to the host (<= 100 x H x W booleans per image) and evaluates Dice,
// E-measure and S-measure with numpy.  For BINARY predictions all three are functions of pixel COUNTS:
//   Dice      : |P & G|, |P|, |G|                                                        (dice_metric.py:57-58)
//   E-measure : the alignment matrix takes one value per (pred, gt) combination           (:110-143)
//   S-measure : object part = means / stds over the gt and non-gt regions; region part = SSIM of the four quadrants
//               around the ground truth's centre of mass                                   (:146-240)
// so the device only has to produce, per (prediction, ground truth) pair, the 2 x 2 contingency table of each quadrant.
// Integer work: bit-exact.  The closed forms are evaluated on the host in float64 (adapteacher/evaluation/dice_metric.py).
//
//   gt_stats_kernel   : per ground-truth mask  n = |G|, sum of row indices, sum of column indices and the quadrant split
//                       (y, x) = (int(round(cy)) + 1, int(round(cx)) + 1), round = half-to-even like Python's (:227-229)
//   pair_counts_kernel: per pair, 4 quadrants x {n11, n10, n01, n00}
// HBM-bound: every pair reads H x W bytes of the prediction and of the ground truth once (uchar4-vectorised, coalesced);
// algorithmic bytes per pair = 2 H W.
#include "common.cuh"

namespace ttdg {

__device__ __forceinline__ long long block_sum_ll(long long v, long long *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TTDG_FULL, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    long long t = 0;
    for (int w = 0; w < nw; ++w) t += sh[w];
    return t;
}

// out[g] = {n, sum_y, sum_x, split_y, split_x}
__global__ void __launch_bounds__(256)
gt_stats_kernel(const unsigned char *__restrict__ gt, int H, int W, long long *__restrict__ out) {
    __shared__ long long sh[8];
    const unsigned char *m = gt + (size_t)blockIdx.x * H * W;
    long long n = 0, sy = 0, sx = 0;
    for (int p = threadIdx.x; p < H * W; p += blockDim.x)
        if (m[p]) { ++n; sy += p / W; sx += p % W; }
    n = block_sum_ll(n, sh); sy = block_sum_ll(sy, sh); sx = block_sum_ll(sx, sh);
    if (threadIdx.x == 0) {
        long long *o = out + (size_t)blockIdx.x * 5;
        o[0] = n; o[1] = sy; o[2] = sx;
        // ndimage.center_of_mass of an empty mask is nan (the reference never reaches region() then: y == 0 returns early)
        o[3] = n > 0 ? (long long)rint((double)sy / (double)n) + 1 : 0;
        o[4] = n > 0 ? (long long)rint((double)sx / (double)n) + 1 : 0;
    }
}

// out[pair][q][k]: q = quadrant (top-left, top-right, bottom-left, bottom-right), k = (n11, n10, n01, n00) with the
// prediction as the first index
__global__ void __launch_bounds__(256)
pair_counts_kernel(const unsigned char *__restrict__ pred, const unsigned char *__restrict__ gt, const int *__restrict__ pairs,
                   const long long *__restrict__ gstats, int H, int W, long long *__restrict__ out) {
    __shared__ long long sh[8];
    const int pi = pairs[2 * blockIdx.x], gi = pairs[2 * blockIdx.x + 1];
    const unsigned char *P = pred + (size_t)pi * H * W, *G = gt + (size_t)gi * H * W;
    const int ys = (int)gstats[(size_t)gi * 5 + 3], xs = (int)gstats[(size_t)gi * 5 + 4];
    int c[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) c[k] = 0;
    const int n = H * W;
    if ((W & 3) == 0 && (((uintptr_t)P | (uintptr_t)G) & 3) == 0) {
        const uchar4 *P4 = reinterpret_cast<const uchar4 *>(P), *G4 = reinterpret_cast<const uchar4 *>(G);
        for (int q4 = threadIdx.x; q4 < n / 4; q4 += blockDim.x) {
            const uchar4 a = P4[q4], b = G4[q4];
            const int p0 = q4 * 4, y = p0 / W, x0 = p0 - y * W;
            const int qy = (y >= ys) ? 8 : 0;
            const unsigned char av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int idx = qy + ((x0 + e >= xs) ? 4 : 0) + (av[e] ? 0 : 2) + (bv[e] ? 0 : 1);
#pragma unroll
                for (int k = 0; k < 16; ++k) c[k] += (idx == k);
            }
        }
    } else {
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            const int y = p / W, x = p - y * W;
            const int idx = ((y >= ys) ? 8 : 0) + ((x >= xs) ? 4 : 0) + (P[p] ? 0 : 2) + (G[p] ? 0 : 1);
#pragma unroll
            for (int k = 0; k < 16; ++k) c[k] += (idx == k);
        }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const long long t = block_sum_ll((long long)c[k], sh);
        if (threadIdx.x == 0) out[(size_t)blockIdx.x * 16 + k] = t;
    }
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_mask_gt_stats(const unsigned char *gt, int G, int H, int W, int64_t *stats, void *stream) {
    TTDG_CHECK_ARG(gt && stats && G >= 0 && H > 0 && W > 0);
    if (G == 0) return 0;
    count_launches(1);
    gt_stats_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(gt, H, W, reinterpret_cast<long long *>(stats));
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_mask_pair_counts(const unsigned char *pred, const unsigned char *gt, const int32_t *pairs, int n_pairs,
                                     const int64_t *gt_stats, int H, int W, int64_t *counts, void *stream) {
    TTDG_CHECK_ARG(pred && gt && pairs && gt_stats && counts && n_pairs >= 0 && H > 0 && W > 0);
    if (n_pairs == 0) return 0;
    count_launches(1);
    pair_counts_kernel<<<n_pairs, 256, 0, (cudaStream_t)stream>>>(pred, gt, pairs, reinterpret_cast<const long long *>(gt_stats), H, W,
                                                                  reinterpret_cast<long long *>(counts));
    TTDG_LAUNCH_RET();
}
