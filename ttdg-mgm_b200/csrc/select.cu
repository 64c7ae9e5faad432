// select.cu - device-side candidate selection of the detector: no torch sort / top-k / gather and no host round trip
// between the RPN head and the box head, or between the box predictor and its NMS.
//
// Restates Detectron2 0.5 `find_top_rpn_proposals` (reached from adapteacher/modeling/proposal_generator/rpn.py:52-54;
// SURVEY Appendix A, K4) and the candidate ordering of `fast_rcnn_inference_single_image` (roi_heads/roi_heads.py:173-205):
//   per (image, level): top-k of the H * W * A objectness logits, sorted descending (ties: lower anchor index first) -> decode
//   the selected anchors -> per image: drop non-finite / empty boxes, order all levels' candidates by score -> per-level NMS
//   (ttdg_nms) -> first POST_NMS_TOPK.
// Outputs are PADDED to a fixed capacity with a per-image count on the device; the one host read of a pass happens where
// the reference's API needs variable-length Python lists (the detections handed to the node sampler / the evaluator).
#include "common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace ttdg {
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t f32_ord(float x) {            // order-preserving float -> uint32 (ascending)
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_unord(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// descending bitonic sort of n (power of two) 64-bit keys in shared memory by the whole CTA
__device__ void bitonic_desc(unsigned long long *a, int n) {
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long x = a[i], y = a[p];
                    const bool desc = (i & k) == 0;
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[p] = x; }
                }
            }
        }
    __syncthreads();
}

constexpr int SEL_LEVELS = 5;
constexpr int SEL_MAX_IMG = 64;
constexpr int SEL_TOPK_CAP = 2048;

struct RpnSelParams {
    const float *logits[SEL_LEVELS], *deltas[SEL_LEVELS];      // NHWC head outputs of every level (N x H x W x ld)
    int H[SEL_LEVELS], W[SEL_LEVELS], stride[SEL_LEVELS], k[SEL_LEVELS], off[SEL_LEVELS];
    int ld_logits, ld_deltas, A, n_img, k_total;
    float img_h[SEL_MAX_IMG], img_w[SEL_MAX_IMG];
    float cell[16][4];
    float clampv;
    float *boxes;            // [n_img][k_total][4]
    float *scores;           // [n_img][k_total]
    unsigned char *valid;    // [n_img][k_total]
};

// Box2BoxTransform.apply_deltas with unit weights, the reference's operation order (detect.cu has the general form)
__device__ __forceinline__ void rpn_apply(const float a[4], const float *d, float clampv, float o[4]) {
    const float w = __fsub_rn(a[2], a[0]), h = __fsub_rn(a[3], a[1]);
    const float cx = __fadd_rn(a[0], __fmul_rn(0.5f, w)), cy = __fadd_rn(a[1], __fmul_rn(0.5f, h));
    const float dx = __fdiv_rn(d[0], 1.f), dy = __fdiv_rn(d[1], 1.f), dw = fminf(__fdiv_rn(d[2], 1.f), clampv), dh = fminf(__fdiv_rn(d[3], 1.f), clampv);
    const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    o[0] = __fsub_rn(pcx, __fmul_rn(0.5f, pw)); o[1] = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    o[2] = __fadd_rn(pcx, __fmul_rn(0.5f, pw)); o[3] = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
}

// grid (levels * CL, images) in thread-block clusters of CL CTAs per (level, image), 1024 threads.  Radix select of the k-th
// largest logit (3 digits of 11 / 11 / 10 bits of the ordered key, run-length-aggregated shared histogram), collection of the
// selected (key, anchor index) pairs, bitonic sort, anchor decoding.  The stride-4 level holds 3/4 of all anchors (245 760 per
// image at 512 x 512): with one CTA per (level, image) eight SMs did 3/4 of the work while 140 idled (0.57 ms per launch).
// With CL > 1 the CTAs of a cluster scan 1 / CL of the map each; after every digit pass they add up each other's histograms
// through distributed shared memory (every CTA redundantly, so all agree on the digit without a broadcast) and they append
// their selected candidates to CTA 0's list with remote shared-memory atomics; CTA 0 sorts and decodes.
template <int CL>
__global__ void __launch_bounds__(1024)
rpn_topk_decode_kernel(const __grid_constant__ RpnSelParams p) {
    __shared__ unsigned long long sel[SEL_TOPK_CAP];
    __shared__ unsigned int hist[2048], htot[CL > 1 ? 2048 : 1];
    __shared__ unsigned int s_prefix, s_remaining, s_cnt, s_tie_base;
    __shared__ unsigned int warp_off[32];
    const int l = blockIdx.x / CL, n = blockIdx.y, tid = threadIdx.x;
    const int rank = CL > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const int HW = p.H[l] * p.W[l], A = p.A, T = HW * A, k = p.k[l];
    const float *lg = p.logits[l] + (size_t)n * HW * p.ld_logits;
    if (k <= 0) return;                                  // uniform over the cluster
    const int per_cta = (HW + CL - 1) / CL;
    const int pix_lo = rank * per_cta, pix_hi = min(HW, pix_lo + per_cta);
    unsigned int kth = 0u, need_eq = 0u, n_eq = 0u;   // k-th largest key; how many elements EQUAL to it are selected / exist
    bool all = T <= k;
    if (!all) {
        if (tid == 0) { s_prefix = 0u; s_remaining = (unsigned)k; }
        // three digits of 11 / 11 / 10 bits (sign + exponent + 2 mantissa bits first: logits of one map share their top BYTE, so
        // an 8-bit first digit would resolve nothing and every pass would rescan everything at full contention)
        for (int pass = 2; pass >= 0; --pass) {
            const int shift = pass == 2 ? 21 : (pass == 1 ? 10 : 0), bits = pass == 0 ? 10 : 11;
            for (int b = tid; b < 2048; b += 1024) hist[b] = 0u;
            __syncthreads();
            const unsigned int prefix = s_prefix, hmask = pass == 2 ? 0u : (0xFFFFFFFFu << (shift + bits));
            // run-length aggregation in registers: the 15 anchors of a pixel and neighbouring pixels of a thread mostly fall into the
            // same coarse bin, so a thread issues one shared-memory atomic per RUN instead of one per element
            unsigned int run_bin = 0xFFFFFFFFu, run_cnt = 0u;
            for (int pix = pix_lo + tid; pix < pix_hi; pix += 1024) {
                const float *row = lg + (size_t)pix * p.ld_logits;
                for (int a = 0; a < A; ++a) {
                    const unsigned int key = f32_ord(row[a]);
                    if ((key & hmask) == prefix) {
                        const unsigned int bin = (key >> shift) & ((1u << bits) - 1u);
                        if (bin == run_bin) ++run_cnt;
                        else { if (run_cnt) atomicAdd(&hist[run_bin], run_cnt); run_bin = bin; run_cnt = 1u; }
                    }
                }
            }
            if (run_cnt) atomicAdd(&hist[run_bin], run_cnt);
            __syncthreads();
            const unsigned int *H = hist;
            if (CL > 1) {
                cg::cluster_group cluster = cg::this_cluster();
                cluster.sync();                                   // every CTA's histogram of this pass is complete
                for (int b = tid; b < 2048; b += 1024) {
                    unsigned int t = 0u;
#pragma unroll
                    for (int r = 0; r < CL; ++r) t += cluster.map_shared_rank(hist, r)[b];
                    htot[b] = t;
                }
                __syncthreads();
                H = htot;
            }
            if (tid < 32) {                           // warp 0: the bin where the running count from the top reaches `remaining`
                unsigned int rem = s_remaining;
                const int nb = 1 << bits, per = nb / 32;
                unsigned int mine = 0u;               // lane l owns bins [nb - (l + 1) * per, nb - l * per): lane 0 = the top bins
                for (int q = 0; q < per; ++q) mine += H[nb - 1 - (tid * per + q)];
                unsigned int incl = mine;
                for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(TTDG_FULL, incl, o); if (tid >= o) incl += t; }
                const unsigned int excl = incl - mine;
                const unsigned int hit = __ballot_sync(TTDG_FULL, incl >= rem);
                const int owner = hit ? __ffs(hit) - 1 : 31;
                if (tid == owner) {
                    unsigned int r = rem - excl;
                    int b = nb - 1 - tid * per;
                    for (int q = 0; q < per - 1; ++q, --b) { if (H[b] >= r) break; r -= H[b]; }
                    s_prefix = prefix | ((unsigned)b << shift);
                    s_remaining = r;                  // elements still to take inside bin b
                    if (pass == 0) s_cnt = H[b];      // (scratch) how many elements carry exactly the k-th key
                }
            }
            __syncthreads();
            if (CL > 1) cg::this_cluster().sync();                // all remote reads of hist done before the next pass clears it
        }
        kth = s_prefix; need_eq = s_remaining; n_eq = s_cnt;
        __syncthreads();
    }
    // ---- collect: everything above the threshold (any order: the sort below fixes it); the ties too when all of them are taken
    const bool ties_all = !all && n_eq == need_eq;
    if (tid == 0) { s_cnt = 0u; s_tie_base = 0u; }
    __syncthreads();
    unsigned long long *sel0 = sel;
    unsigned int *cnt0 = &s_cnt;
    if (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();                                           // CTA 0's counter is reset
        sel0 = cluster.map_shared_rank(sel, 0);
        cnt0 = cluster.map_shared_rank(&s_cnt, 0);
    }
    for (int pix = pix_lo + tid; pix < pix_hi; pix += 1024) {
        const float *row = lg + (size_t)pix * p.ld_logits;
        for (int a = 0; a < A; ++a) {
            const unsigned int key = f32_ord(row[a]);
            if (all || key > kth || (ties_all && key == kth)) {
                const unsigned int pos = atomicAdd(cnt0, 1u);
                if (pos < SEL_TOPK_CAP) sel0[pos] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(pix * A + a));
            }
        }
    }
    __syncthreads();
    if (CL > 1) {
        cg::this_cluster().sync();                                // every CTA's candidates have landed in CTA 0
        if (rank != 0) return;
    }
    if (!all && !ties_all) {
        // more candidates tied at the k-th value than slots left: exactly need_eq of them by lowest index (deterministic) -
        // chunks of 1024 consecutive elements, block-wide exclusive scan of the tie flags
        const unsigned int n_gt = s_cnt;
        for (int base = 0; base < T && s_tie_base < need_eq; base += 1024) {
            const int e = base + tid;
            bool tie = false;
            if (e < T) { const int pix = e / A, a = e - pix * A; tie = f32_ord(lg[(size_t)pix * p.ld_logits + a]) == kth; }
            const unsigned int bal = __ballot_sync(TTDG_FULL, tie);
            if ((tid & 31) == 0) warp_off[tid >> 5] = __popc(bal);
            __syncthreads();
            if (tid < 32) {
                unsigned int v = warp_off[tid], incl = v;
                for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(TTDG_FULL, incl, o); if (tid >= o) incl += t; }
                warp_off[tid] = incl - v;
                if (tid == 31) hist[0] = incl;        // ties in this chunk
            }
            __syncthreads();
            const unsigned int rank = s_tie_base + warp_off[tid >> 5] + __popc(bal & ((1u << (tid & 31)) - 1u));
            if (tie && rank < need_eq) sel[n_gt + rank] = ((unsigned long long)kth << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)e);
            __syncthreads();
            if (tid == 0) s_tie_base += hist[0];
            __syncthreads();
        }
    }
    const int got = all ? T : k;
    int npad = 1;
    while (npad < got) npad <<= 1;
    for (int i = got + tid; i < npad; i += 1024) sel[i] = 0ull;
    bitonic_desc(sel, npad);
    // ---- decode
    const float ih = p.img_h[n], iw = p.img_w[n];
    for (int j = tid; j < got; j += 1024) {
        const unsigned long long e = sel[j];
        const float score = f32_unord((unsigned int)(e >> 32));
        const int idx = (int)(0xFFFFFFFFu - (unsigned int)e);
        const int pix = idx / A, a = idx - pix * A;
        const int y = pix / p.W[l], x = pix - y * p.W[l];
        const float sx = (float)(x * p.stride[l]), sy = (float)(y * p.stride[l]);
        const float anc[4] = {sx + p.cell[a][0], sy + p.cell[a][1], sx + p.cell[a][2], sy + p.cell[a][3]};
        float b[4];
        rpn_apply(anc, p.deltas[l] + ((size_t)n * HW + pix) * p.ld_deltas + a * 4, p.clampv, b);
        const bool fin = isfinite(b[0]) && isfinite(b[1]) && isfinite(b[2]) && isfinite(b[3]) && isfinite(score);
        b[0] = fminf(fmaxf(b[0], 0.f), iw); b[1] = fminf(fmaxf(b[1], 0.f), ih);
        b[2] = fminf(fmaxf(b[2], 0.f), iw); b[3] = fminf(fmaxf(b[3], 0.f), ih);
        const size_t o = (size_t)n * p.k_total + p.off[l] + j;
        *reinterpret_cast<float4 *>(p.boxes + o * 4) = make_float4(b[0], b[1], b[2], b[3]);
        p.scores[o] = score;
        p.valid[o] = fin && (b[2] - b[0] > 0.f) && (b[3] - b[1] > 0.f);
    }
}

// Per image: order n candidates by (valid first, score descending, position ascending) and write them out in that order with
// the category ttdg_nms needs: valid -> cats_in[position] (cat_mod == 0) or position % cat_mod; invalid -> a unique negative
// value (never suppresses, never suppressed).  valid == NULL: valid <=> score > 0 (the box predictor marks rejected candidates
// with score -1).  grid = images, 1024 threads, n_pad * 8 bytes of shared memory (n_pad = next power of two >= n).
__global__ void __launch_bounds__(1024)
sort_candidates_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, const unsigned char *__restrict__ valid,
                       const int32_t *__restrict__ cats_in, int cat_mod, int n, int n_pad, float invalid_score,
                       float *__restrict__ boxes_out, float *__restrict__ scores_out, int32_t *__restrict__ cats_out,
                       int32_t *__restrict__ n_valid, int n_out) {
    extern __shared__ unsigned long long keys[];
    __shared__ int s_nv;
    const int img = blockIdx.x, tid = threadIdx.x;
    boxes += (size_t)img * n * 4; scores += (size_t)img * n;
    if (valid) valid += (size_t)img * n;
    if (tid == 0) s_nv = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < n_pad; i += 1024) {
        unsigned long long key = 0ull;
        if (i < n) {
            const float s = scores[i];
            const bool ok = valid ? (valid[i] != 0 && isfinite(s)) : (s > 0.f);
            if (ok) { key = ((unsigned long long)f32_ord(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i); ++mine; }
        }
        keys[i] = key;
    }
    if (mine) atomicAdd(&s_nv, mine);
    bitonic_desc(keys, n_pad);
    const int nv = s_nv;
    // n_out rows are written (n_out == n: the whole ordering; n_out != n: the first n_out candidates, padded - then n_valid
    // is capped at n_out and is the row count of the padded output)
    if (tid == 0) n_valid[img] = nv < n_out ? nv : n_out;
    boxes_out += (size_t)img * n_out * 4; scores_out += (size_t)img * n_out;
    if (cats_out) cats_out += (size_t)img * n_out;
    for (int j = tid; j < n_out; j += 1024) {
        if (j < nv) {
            const int src = (int)(0xFFFFFFFFu - (unsigned int)keys[j]);
            *reinterpret_cast<float4 *>(boxes_out + (size_t)j * 4) = *reinterpret_cast<const float4 *>(boxes + (size_t)src * 4);
            scores_out[j] = scores[src];
            if (cats_out) cats_out[j] = cat_mod > 0 ? src % cat_mod : cats_in[src];
        } else {
            *reinterpret_cast<float4 *>(boxes_out + (size_t)j * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            scores_out[j] = invalid_score;
            if (cats_out) cats_out[j] = -1 - j;
        }
    }
}

// Per image: the kept candidates (keep[0 .. n_keep), ascending indices into the sorted arrays) that are valid, padded to
// max_keep rows.  counts[img] = how many rows are real.  Padding rows: zero box, pad_score, class -1.
__global__ void __launch_bounds__(256)
gather_kept_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, const int32_t *__restrict__ cats,
                   const int32_t *__restrict__ keep, const int32_t *__restrict__ n_keep, const int32_t *__restrict__ n_valid, int n,
                   int max_keep, float pad_score, float *__restrict__ boxes_out, float *__restrict__ scores_out,
                   int64_t *__restrict__ cats_out, int32_t *__restrict__ counts) {
    const int img = blockIdx.x;
    const int nk = min(n_keep[img], max_keep), nv = n_valid[img];
    keep += (size_t)img * max_keep;
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int j = threadIdx.x; j < nk; j += 256) mine += keep[j] < nv ? 1 : 0;     // keep is ascending: the valid ones come first
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    const int cnt = s_cnt;
    if (threadIdx.x == 0) counts[img] = cnt;
    for (int j = threadIdx.x; j < max_keep; j += 256) {
        const size_t o = (size_t)img * max_keep + j;
        if (j < cnt) {
            const size_t s = (size_t)img * n + keep[j];
            *reinterpret_cast<float4 *>(boxes_out + o * 4) = *reinterpret_cast<const float4 *>(boxes + s * 4);
            scores_out[o] = scores[s];
            if (cats_out) cats_out[o] = (int64_t)cats[s];
        } else {
            *reinterpret_cast<float4 *>(boxes_out + o * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            scores_out[o] = pad_score;
            if (cats_out) cats_out[o] = -1;
        }
    }
}

// rois [n_img * P][5] = {image, box} for RoIAlign from padded proposals; padding rows become an empty box of image 0
__global__ void __launch_bounds__(256)
rois_from_padded_kernel(const float *__restrict__ boxes, const int32_t *__restrict__ counts, int n_img, int P, float *__restrict__ rois) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n_img * P) return;
    const int img = t / P, j = t - img * P;
    float *r = rois + (size_t)t * 5;
    if (j < counts[img]) {
        const float4 b = *reinterpret_cast<const float4 *>(boxes + (size_t)t * 4);
        r[0] = (float)img; r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
    } else {
        r[0] = 0.f; r[1] = 0.f; r[2] = 0.f; r[3] = 0.f; r[4] = 0.f;
    }
}

// RPN NMS per (image, level).  find_top_rpn_proposals applies batched_nms with the LEVEL as the category, i.e. the levels
// never interact: instead of one greedy sweep over the 8960 score-ordered candidates of an image (80 MB of suppression mask,
// 140 sequential rounds on 8 CTAs: 0.96 + 0.61 ms per step) every (image, level) pair - already sorted by rpn_topk_decode_kernel -
// is its own problem of <= 2048 boxes that lives in shared memory: grid (levels, images), rounds of 64 boxes, (1) the
// 64 x 64 diagonal block by warp ballots, (2) thread 0 resolves the round greedily (torchvision's order), (3) the kept
// boxes of the round against the later, still-alive boxes.  Same IoU expression as ttdg_nms (detect.cu).  Output: kept[img][pos]
// (uint8) - the final "first POST_NMS_TOPK by score" is one ordering pass over these flags (sort_candidates_kernel).  A level
// stops after max_keep kept boxes: later ones cannot be among the image's first max_keep.
__device__ __forceinline__ bool sel_nms_hit(float4 a, float area_a, float4 b, float thresh) {
    const float xx0 = fmaxf(a.x, b.x), yy0 = fmaxf(a.y, b.y), xx1 = fminf(a.z, b.z), yy1 = fminf(a.w, b.w);
    const float w = fmaxf(xx1 - xx0, 0.f), h = fmaxf(yy1 - yy0, 0.f);
    const float inter = w * h;
    const float areab = (b.z - b.x) * (b.w - b.y);
    return inter / (area_a + areab - inter) > thresh;
}

struct RpnNmsParams {
    const float *boxes;              // [n_img][k_total][4], every level's segment sorted by score (descending)
    const unsigned char *valid;      // [n_img][k_total]
    unsigned char *kept;             // [n_img][k_total]
    int off[SEL_LEVELS], k[SEL_LEVELS];
    int k_total, max_keep;
    float thresh;
};

__global__ void __launch_bounds__(1024)
rpn_nms_levels_kernel(const __grid_constant__ RpnNmsParams p) {
    __shared__ float4 sb[SEL_TOPK_CAP];
    __shared__ unsigned long long removed[SEL_TOPK_CAP / 64], keptw[SEL_TOPK_CAP / 64];
    __shared__ unsigned long long diag[64];
    __shared__ unsigned long long kept_bits;
    __shared__ int cnt;
    __shared__ float4 kb4[64];
    __shared__ float karea[64];
    const int l = blockIdx.x, img = blockIdx.y, n = p.k[l];
    if (n <= 0) return;
    const size_t base0 = (size_t)img * p.k_total + p.off[l];
    const float *boxes = p.boxes + base0 * 4;
    const unsigned char *valid = p.valid + base0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int words = (n + 63) / 64;
    for (int w = tid; w < SEL_TOPK_CAP / 64; w += 1024) { removed[w] = 0ull; keptw[w] = 0ull; }
    if (tid == 0) cnt = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        sb[i] = reinterpret_cast<const float4 *>(boxes)[i];
        if (!valid[i]) atomicOr(&removed[i >> 6], 1ull << (i & 63));             // dropped before the NMS: never kept, never suppresses
    }
    __syncthreads();
    for (int wb = 0; wb < words; ++wb) {
        const int base = wb * 64, nb = min(64, n - base);
        // (1) diagonal block: warp w owns rows 2 w and 2 w + 1, lane = column (and column + 32)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * warp + rr;
            bool lo = false, hi = false;
            if (r < nb) {
                const float4 a = sb[base + r];
                const float area = (a.z - a.x) * (a.w - a.y);
                if (lane > r && lane < nb) lo = sel_nms_hit(a, area, sb[base + lane], p.thresh);
                if (lane + 32 > r && lane + 32 < nb) hi = sel_nms_hit(a, area, sb[base + lane + 32], p.thresh);
            }
            const unsigned blo = __ballot_sync(0xffffffffu, lo), bhi = __ballot_sync(0xffffffffu, hi);
            if (lane == 0 && r < 64) diag[r] = (unsigned long long)blo | ((unsigned long long)bhi << 32);
        }
        __syncthreads();
        // (2) greedy resolution of the round
        if (tid == 0) {
            unsigned long long dead = removed[wb], kbits = 0ull;
            int c = cnt;
            for (int j = 0; j < nb && c < p.max_keep; ++j)
                if (!((dead >> j) & 1ull)) { kbits |= 1ull << j; ++c; dead |= diag[j]; }
            kept_bits = kbits; keptw[wb] = kbits;
            cnt = c;
        }
        __syncthreads();
        if (cnt >= p.max_keep) break;
        // (3) kept boxes of the round against every later box that is still alive
        const unsigned long long kbits = kept_bits;
        const int nk = __popcll(kbits);
        if (tid < 64 && ((kbits >> tid) & 1ull)) {
            const int pos = __popcll(kbits & ((1ull << tid) - 1ull));
            const float4 a = sb[base + tid];
            kb4[pos] = a; karea[pos] = (a.z - a.x) * (a.w - a.y);
        }
        __syncthreads();
        if (nk > 0)
            for (int j = base + 64 + tid; j < n; j += 1024) {
                if ((removed[j >> 6] >> (j & 63)) & 1ull) continue;
                const float4 b = sb[j];
                bool hit = false;
                for (int k = 0; k < nk && !hit; ++k) hit = sel_nms_hit(kb4[k], karea[k], b, p.thresh);
                if (hit) atomicOr(&removed[j >> 6], 1ull << (j & 63));
            }
        __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) p.kept[base0 + i] = (unsigned char)((keptw[i >> 6] >> (i & 63)) & 1ull);
}

// candidates of padding proposals never survive: score -1
__global__ void __launch_bounds__(256)
mask_padded_candidates_kernel(float *__restrict__ cand_scores, const int32_t *__restrict__ counts, int n_img, int P, int K) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n_img * P * K) return;
    const int img = t / (P * K), j = (t - img * P * K) / K;
    if (j >= counts[img]) cand_scores[t] = -1.f;
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_rpn_select(const void *const *logits_h, const void *const *deltas_h, const int32_t *lvl_hw_h, const int32_t *strides_h,
                               const int32_t *k_h, int n_levels, int ld_logits, int ld_deltas, int A, const float *cell_anchors_h,
                               int n_img, const float *img_hw_h, float *boxes, float *scores, unsigned char *valid, void *stream) {
    TTDG_CHECK_ARG(logits_h && deltas_h && lvl_hw_h && strides_h && k_h && cell_anchors_h && img_hw_h && boxes && scores && valid);
    TTDG_CHECK_ARG(n_levels >= 1 && n_levels <= SEL_LEVELS && A >= 1 && A <= 16 && n_img >= 0);
    if (n_img > SEL_MAX_IMG) return TTDG_E_LIMIT;
    if (n_img == 0) return 0;
    RpnSelParams p = {};
    int off = 0;
    for (int l = 0; l < SEL_LEVELS; ++l) {
        if (l < n_levels) {
            p.logits[l] = reinterpret_cast<const float *>(logits_h[l]); p.deltas[l] = reinterpret_cast<const float *>(deltas_h[l]);
            p.H[l] = lvl_hw_h[2 * l]; p.W[l] = lvl_hw_h[2 * l + 1]; p.stride[l] = strides_h[l]; p.k[l] = k_h[l];
            if (p.k[l] < 0 || p.k[l] > SEL_TOPK_CAP || p.k[l] > p.H[l] * p.W[l] * A) return TTDG_E_LIMIT;
            p.off[l] = off; off += p.k[l];
        }
    }
    p.ld_logits = ld_logits; p.ld_deltas = ld_deltas; p.A = A; p.n_img = n_img; p.k_total = off;
    for (int i = 0; i < n_img; ++i) { p.img_h[i] = img_hw_h[2 * i]; p.img_w[i] = img_hw_h[2 * i + 1]; }
    for (int a = 0; a < A; ++a) for (int c = 0; c < 4; ++c) p.cell[a][c] = cell_anchors_h[a * 4 + c];
    p.clampv = logf(1000.f / 16.f);
    p.boxes = boxes; p.scores = scores; p.valid = valid;
    count_launches(1);
    static int cl = -1;                                   // TTDG_SEL_CLUSTER = 1 | 8 (default): CTAs per (level, image)
    if (cl < 0) { const char *e = getenv("TTDG_SEL_CLUSTER"); cl = (e && atoi(e) == 1) ? 1 : 8; }
    if (cl == 1) {
        rpn_topk_decode_kernel<1><<<dim3(n_levels, n_img), 1024, 0, (cudaStream_t)stream>>>(p);
        TTDG_LAUNCH_RET();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_levels * 8, n_img);
    cfg.blockDim = dim3(1024);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, rpn_topk_decode_kernel<8>, p);
}

extern "C" int ttdg_sort_candidates(const float *boxes, const float *scores, const unsigned char *valid, const int32_t *cats_in,
                                    int cat_mod, int n_img, int n, float invalid_score, float *boxes_out, float *scores_out,
                                    int32_t *cats_out, int32_t *n_valid, void *stream) {
    TTDG_CHECK_ARG(boxes && scores && boxes_out && scores_out && cats_out && n_valid && n_img >= 0 && n >= 0 && (cat_mod > 0 || cats_in));
    if (n_img == 0) return 0;
    if (n == 0) return (int)cudaMemsetAsync(n_valid, 0, sizeof(int32_t) * n_img, (cudaStream_t)stream);
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * 8;
    if (smem > 200 * 1024) return TTDG_E_LIMIT;                  // n <= 16384 (RPN: 5 levels x 2000 = 8960 at 512 x 512)
    cudaError_t e = cudaFuncSetAttribute(sort_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    count_launches(1);
    sort_candidates_kernel<<<n_img, 1024, smem, (cudaStream_t)stream>>>(boxes, scores, valid, cats_in, cat_mod, n, n_pad, invalid_score,
                                                                        boxes_out, scores_out, cats_out, n_valid, n);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_rpn_nms_levels(const float *boxes, const unsigned char *valid, const int32_t *k_h, int n_levels, int n_img,
                                   float iou_thresh, int max_keep, unsigned char *kept, void *stream) {
    TTDG_CHECK_ARG(boxes && valid && k_h && kept && n_levels >= 1 && n_levels <= SEL_LEVELS && n_img >= 0 && max_keep >= 1);
    if (n_img == 0) return 0;
    if (n_img > 65535) return TTDG_E_LIMIT;
    RpnNmsParams p = {};
    int off = 0;
    for (int l = 0; l < n_levels; ++l) {
        if (k_h[l] < 0 || k_h[l] > SEL_TOPK_CAP) return TTDG_E_LIMIT;
        p.off[l] = off; p.k[l] = k_h[l]; off += k_h[l];
    }
    if (off == 0) return 0;
    p.boxes = boxes; p.valid = valid; p.kept = kept; p.k_total = off; p.max_keep = max_keep; p.thresh = iou_thresh;
    count_launches(1);
    rpn_nms_levels_kernel<<<dim3(n_levels, n_img), 1024, 0, (cudaStream_t)stream>>>(p);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_top_candidates(const float *boxes, const float *scores, const unsigned char *valid, int n_img, int n, int n_out,
                                   float pad_score, float *boxes_out, float *scores_out, int32_t *counts, void *stream) {
    TTDG_CHECK_ARG(boxes && scores && valid && boxes_out && scores_out && counts && n_img >= 0 && n >= 0 && n_out >= 1);
    if (n_img == 0) return 0;
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * 8;
    if (smem > 200 * 1024) return TTDG_E_LIMIT;
    cudaError_t e = cudaFuncSetAttribute(sort_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    count_launches(1);
    sort_candidates_kernel<<<n_img, 1024, smem, (cudaStream_t)stream>>>(boxes, scores, valid, nullptr, 1, n, n_pad, pad_score, boxes_out,
                                                                        scores_out, nullptr, counts, n_out);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_gather_kept(const float *boxes, const float *scores, const int32_t *cats, const int32_t *keep, const int32_t *n_keep,
                                const int32_t *n_valid, int n_img, int n, int max_keep, float pad_score, float *boxes_out,
                                float *scores_out, int64_t *cats_out, int32_t *counts, void *stream) {
    TTDG_CHECK_ARG(boxes && scores && cats && keep && n_keep && n_valid && boxes_out && scores_out && counts && n_img >= 0 && max_keep >= 1);
    if (n_img == 0) return 0;
    count_launches(1);
    gather_kept_kernel<<<n_img, 256, 0, (cudaStream_t)stream>>>(boxes, scores, cats, keep, n_keep, n_valid, n, max_keep, pad_score,
                                                               boxes_out, scores_out, cats_out, counts);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_rois_from_padded(const float *boxes, const int32_t *counts, int n_img, int P, float *rois, void *stream) {
    TTDG_CHECK_ARG(boxes && counts && rois && n_img >= 0 && P >= 0);
    if (n_img * P == 0) return 0;
    count_launches(1);
    rois_from_padded_kernel<<<ceil_div(n_img * P, 256), 256, 0, (cudaStream_t)stream>>>(boxes, counts, n_img, P, rois);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_mask_padded_candidates(float *cand_scores, const int32_t *counts, int n_img, int P, int K, void *stream) {
    TTDG_CHECK_ARG(cand_scores && counts && n_img >= 0 && P >= 0 && K >= 1);
    if (n_img * P == 0) return 0;
    count_launches(1);
    mask_padded_candidates_kernel<<<ceil_div(n_img * P * K, 256), 256, 0, (cudaStream_t)stream>>>(cand_scores, counts, n_img, P, K);
    TTDG_LAUNCH_RET();
}
