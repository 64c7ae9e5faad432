// detect.cu - proposal / detection / mask post-processing kernels of the detector (NHWC pyramid).
// Restates Detectron2 0.5 operators reached from adapteacher/modeling/proposal_generator/rpn.py:52-54
// (predict_proposals), roi_heads/roi_heads.py:173-205 (_forward_box -> ROIPooler / FastRCNNOutputLayers.inference),
// :112 (forward_with_given_boxes -> mask head) and detector_postprocess.
#include "common.cuh"
#include <cstdlib>

namespace ttdg {

// ------------------------------------------------------------------------------------------------ box transform
struct F4 { float x0, y0, x1, y1; };

__device__ __forceinline__ F4 apply_deltas(const F4 &b, float dx, float dy, float dw, float dh, float wx, float wy, float ww,
                                           float wh, float clampv) {
    // Box2BoxTransform.apply_deltas, evaluated in the reference's operation order (no FMA contraction)
    const float w = __fsub_rn(b.x1, b.x0), h = __fsub_rn(b.y1, b.y0);
    const float cx = __fadd_rn(b.x0, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y0, __fmul_rn(0.5f, h));
    dx = __fdiv_rn(dx, wx); dy = __fdiv_rn(dy, wy); dw = fminf(__fdiv_rn(dw, ww), clampv); dh = fminf(__fdiv_rn(dh, wh), clampv);
    const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    F4 o;
    o.x0 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)); o.y0 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    o.x1 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)); o.y1 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
    return o;
}

struct CellAnchors { float a[16][4]; };

// RPN: decode the selected anchors of one level.  idx[n][k] indexes (pixel * A + a); deltas are the NHWC conv output
// with channel (a * 4 + c) and row pitch ldd.  Writes boxes (clipped) and a validity flag (finite, non-empty).
__global__ void __launch_bounds__(256)
rpn_decode_kernel(const float *__restrict__ deltas, int ldd, const int64_t *__restrict__ idx, int Nimg, int k, int Wl, int HW,
                  int A, int stride, CellAnchors ca, float img_h, float img_w, float clampv, float *__restrict__ boxes,
                  unsigned char *__restrict__ valid) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= Nimg * k) return;
    const int n = t / k;
    const int64_t id = idx[t];
    const int pix = (int)(id / A), a = (int)(id - (int64_t)pix * A);
    const int y = pix / Wl, x = pix - y * Wl;
    const float sx = (float)(x * stride), sy = (float)(y * stride);
    F4 anc = {sx + ca.a[a][0], sy + ca.a[a][1], sx + ca.a[a][2], sy + ca.a[a][3]};
    const float *d = deltas + ((size_t)n * HW + pix) * ldd + a * 4;
    F4 b = apply_deltas(anc, d[0], d[1], d[2], d[3], 1.f, 1.f, 1.f, 1.f, clampv);
    const bool fin = isfinite(b.x0) && isfinite(b.y0) && isfinite(b.x1) && isfinite(b.y1);
    b.x0 = fminf(fmaxf(b.x0, 0.f), img_w); b.y0 = fminf(fmaxf(b.y0, 0.f), img_h);
    b.x1 = fminf(fmaxf(b.x1, 0.f), img_w); b.y1 = fminf(fmaxf(b.y1, 0.f), img_h);
    float *o = boxes + (size_t)t * 4;
    o[0] = b.x0; o[1] = b.y0; o[2] = b.x1; o[3] = b.y1;
    valid[t] = fin && (b.x1 - b.x0 > 0.f) && (b.y1 - b.y0 > 0.f);
}

// Box head: softmax over K+1 scores, class-specific deltas (weights 10, 10, 5, 5), clip.  One thread per (roi, class).
// cand_boxes [R*K][4], cand_scores [R*K] (-1 where score <= thresh or non-finite).
__global__ void __launch_bounds__(256)
box_predict_kernel(const float *__restrict__ cls, int ldc, const float *__restrict__ reg, int ldr, const float *__restrict__ props,
                   int R, int K, float img_h, float img_w, float thresh, float clampv, float *__restrict__ cand_boxes,
                   float *__restrict__ cand_scores) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= R * K) return;
    const int r = t / K, c = t - r * K;
    const float *s = cls + (size_t)r * ldc;
    float mx = s[0];
    for (int j = 1; j <= K; ++j) mx = fmaxf(mx, s[j]);
    float den = 0.f;
    for (int j = 0; j <= K; ++j) den += expf(s[j] - mx);
    const float prob = expf(s[c] - mx) / den;
    const float *pb = props + (size_t)r * 4;
    F4 anc = {pb[0], pb[1], pb[2], pb[3]};
    const float *d = reg + (size_t)r * ldr + c * 4;
    F4 b = apply_deltas(anc, d[0], d[1], d[2], d[3], 10.f, 10.f, 5.f, 5.f, clampv);
    bool fin = isfinite(b.x0) && isfinite(b.y0) && isfinite(b.x1) && isfinite(b.y1) && isfinite(prob);
    b.x0 = fminf(fmaxf(b.x0, 0.f), img_w); b.y0 = fminf(fmaxf(b.y0, 0.f), img_h);
    b.x1 = fminf(fmaxf(b.x1, 0.f), img_w); b.y1 = fminf(fmaxf(b.y1, 0.f), img_h);
    float *o = cand_boxes + (size_t)t * 4;
    o[0] = b.x0; o[1] = b.y0; o[2] = b.x1; o[3] = b.y1;
    cand_scores[t] = (fin && prob > thresh) ? prob : -1.f;
}

// ------------------------------------------------------------------------------------------------ NMS (per category)
// boxes sorted by descending score.  mask[i][w] bit j set <=> box (64 w + j) has the same category, comes later and
// IoU > thresh (torchvision nms: inter / (area_a + area_b - inter)).
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float *__restrict__ boxes, const int32_t *__restrict__ cat, int n, float thresh, unsigned long long *__restrict__ mask,
                int words) {
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;
    boxes += (size_t)blockIdx.z * n * 4; cat += (size_t)blockIdx.z * n; mask += (size_t)blockIdx.z * n * words;   // image
    __shared__ float sb[64][4];
    __shared__ int sc[64];
    const int cj = cb * 64 + threadIdx.x;
    if (cj < n) { sb[threadIdx.x][0] = boxes[cj * 4]; sb[threadIdx.x][1] = boxes[cj * 4 + 1]; sb[threadIdx.x][2] = boxes[cj * 4 + 2];
                  sb[threadIdx.x][3] = boxes[cj * 4 + 3]; sc[threadIdx.x] = cat[cj]; }
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i >= n) return;
    const float x0 = boxes[i * 4], y0 = boxes[i * 4 + 1], x1 = boxes[i * 4 + 2], y1 = boxes[i * 4 + 3];
    const float area = (x1 - x0) * (y1 - y0);
    const int ci = cat[i];
    unsigned long long bits = 0ull;
    const int cols = min(64, n - cb * 64);
    for (int j = (rb == cb ? threadIdx.x + 1 : 0); j < cols; ++j) {
        if (sc[j] != ci) continue;
        const float xx0 = fmaxf(x0, sb[j][0]), yy0 = fmaxf(y0, sb[j][1]), xx1 = fminf(x1, sb[j][2]), yy1 = fminf(y1, sb[j][3]);
        const float w = fmaxf(xx1 - xx0, 0.f), h = fmaxf(yy1 - yy0, 0.f);
        const float inter = w * h;
        const float areab = (sb[j][2] - sb[j][0]) * (sb[j][3] - sb[j][1]);
        if (inter / (area + areab - inter) > thresh) bits |= 1ull << j;
    }
    mask[(size_t)i * words + cb] = bits;
}

// Sweep of the suppression mask, one CTA, 64 boxes per round: thread 0 resolves the round's diagonal word serially in
// registers (64 bit-steps, no barriers), then all threads OR the mask rows of the survivors into `removed`.
// keep[] receives the kept indices in order, *n_keep their number (capped at max_keep).
__global__ void __launch_bounds__(256)
nms_sweep_kernel(const unsigned long long *__restrict__ mask, int n, int words, int max_keep, int32_t *__restrict__ keep,
                 int32_t *__restrict__ n_keep) {
    extern __shared__ unsigned long long removed[];
    mask += (size_t)blockIdx.x * n * words; keep += (size_t)blockIdx.x * max_keep; n_keep += blockIdx.x;              // image
    __shared__ unsigned long long diag[64];
    __shared__ unsigned long long kept_bits;
    __shared__ int cnt;
    for (int w = threadIdx.x; w < words; w += 256) removed[w] = 0ull;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    for (int wb = 0; wb < words; ++wb) {
        const int base = wb * 64, nb = min(64, n - base);
        if (threadIdx.x < nb) diag[threadIdx.x] = mask[(size_t)(base + threadIdx.x) * words + wb];
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long dead = removed[wb], kb = 0ull;
            int c = cnt;
            for (int j = 0; j < nb && c < max_keep; ++j) {
                if (!((dead >> j) & 1ull)) { kb |= 1ull << j; keep[c++] = base + j; dead |= diag[j]; }
            }
            kept_bits = kb;
            cnt = c;
        }
        __syncthreads();
        if (cnt >= max_keep) break;
        unsigned long long kb = kept_bits;
        while (kb) {
            const int j = __ffsll((long long)kb) - 1;
            kb &= kb - 1;
            const unsigned long long *row = mask + (size_t)(base + j) * words;
            for (int w = wb + 1 + threadIdx.x; w < words; w += 256) removed[w] |= row[w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_keep = cnt < max_keep ? cnt : max_keep;
}

// Fused NMS: ONE CTA per image, boxes + categories resident in shared memory, the suppression state a bitmask in shared
// memory.  Rounds of 64 boxes: (1) the 64 x 64 diagonal block of the suppression relation by warp ballots, (2) thread 0
// resolves the round serially in registers (greedy order = torchvision's), (3) only the KEPT boxes of the round (<= 64)
// are tested against the later, still-alive boxes - all threads, one candidate per thread and step.  The pair (mask,
// sweep) above evaluates the whole upper triangle (40 M IoUs per image for 8960 RPN candidates, 80 MB of mask) although
// the sweep stops at max_keep kept boxes; here only kept x later pairs up to the stopping round are evaluated and
// nothing goes through global memory.  Same IoU expression, same order: identical keep lists.
constexpr int NMSF_THREADS = 1024;

__device__ __forceinline__ bool nms_hit(float4 a, float area_a, float4 b, float thresh) {
    const float xx0 = fmaxf(a.x, b.x), yy0 = fmaxf(a.y, b.y), xx1 = fminf(a.z, b.z), yy1 = fminf(a.w, b.w);
    const float w = fmaxf(xx1 - xx0, 0.f), h = fmaxf(yy1 - yy0, 0.f);
    const float inter = w * h;
    const float areab = (b.z - b.x) * (b.w - b.y);
    return inter / (area_a + areab - inter) > thresh;
}

__global__ void __launch_bounds__(NMSF_THREADS)
nms_fused_kernel(const float *__restrict__ boxes, const int32_t *__restrict__ cat, int n, float thresh, int max_keep,
                 int32_t *__restrict__ keep, int32_t *__restrict__ n_keep) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    const int words = (n + 63) / 64;
    float4 *sb = reinterpret_cast<float4 *>(nms_smem);                                   // n boxes
    unsigned long long *removed = reinterpret_cast<unsigned long long *>(sb + n);        // words
    int *sc = reinterpret_cast<int *>(removed + words);                                  // n categories
    __shared__ unsigned long long diag[64];
    __shared__ unsigned long long kept_bits;
    __shared__ int cnt;
    __shared__ float4 kb4[64];
    __shared__ float karea[64];
    __shared__ int kcat[64];
    boxes += (size_t)blockIdx.x * n * 4; cat += (size_t)blockIdx.x * n; keep += (size_t)blockIdx.x * max_keep; n_keep += blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < n; i += NMSF_THREADS) { sb[i] = reinterpret_cast<const float4 *>(boxes)[i]; sc[i] = cat[i]; }
    for (int w = tid; w < words; w += NMSF_THREADS) removed[w] = 0ull;
    if (tid == 0) cnt = 0;
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
                const int ca = sc[base + r];
                if (lane > r && lane < nb && sc[base + lane] == ca) lo = nms_hit(a, area, sb[base + lane], thresh);
                if (lane + 32 > r && lane + 32 < nb && sc[base + lane + 32] == ca) hi = nms_hit(a, area, sb[base + lane + 32], thresh);
            }
            const unsigned blo = __ballot_sync(0xffffffffu, lo), bhi = __ballot_sync(0xffffffffu, hi);
            if (lane == 0 && r < 64) diag[r] = (unsigned long long)blo | ((unsigned long long)bhi << 32);
        }
        __syncthreads();
        // (2) greedy resolution of the round
        if (tid == 0) {
            unsigned long long dead = removed[wb], kbits = 0ull;
            int c = cnt;
            for (int j = 0; j < nb && c < max_keep; ++j)
                if (!((dead >> j) & 1ull)) { kbits |= 1ull << j; keep[c++] = base + j; dead |= diag[j]; }
            kept_bits = kbits;
            cnt = c;
        }
        __syncthreads();
        if (cnt >= max_keep) break;
        // (3) kept boxes of the round against every later box that is still alive
        const unsigned long long kbits = kept_bits;
        const int nk = __popcll(kbits);
        if (tid < 64 && ((kbits >> tid) & 1ull)) {
            const int pos = __popcll(kbits & ((1ull << tid) - 1ull));
            const float4 a = sb[base + tid];
            kb4[pos] = a; karea[pos] = (a.z - a.x) * (a.w - a.y); kcat[pos] = sc[base + tid];
        }
        __syncthreads();
        if (nk > 0)
            for (int j = base + 64 + tid; j < n; j += NMSF_THREADS) {
                if ((removed[j >> 6] >> (j & 63)) & 1ull) continue;
                const float4 b = sb[j];
                const int cj = sc[j];
                bool hit = false;
                for (int k = 0; k < nk && !hit; ++k)
                    if (kcat[k] == cj) hit = nms_hit(kb4[k], karea[k], b, thresh);
                if (hit) atomicOr(&removed[j >> 6], 1ull << (j & 63));
            }
        __syncthreads();
    }
    if (tid == 0) *n_keep = cnt < max_keep ? cnt : max_keep;
}

// ------------------------------------------------------------------------------------------------ ROIAlign (aligned, adaptive sampling)
struct Pyr4 { const float *p[4]; int h[4], w[4]; };

__device__ __forceinline__ int roi_level(float x0, float y0, float x1, float y1) {
    // ROIPooler.assign_boxes_to_levels: floor(4 + log2(sqrt(area) / 224 + 1e-8)) clamped to [2, 5], minus 2
    const float sz = sqrtf((x1 - x0) * (y1 - y0));
    float lv = floorf(4.f + log2f(sz / 224.f + 1e-8f));
    lv = fminf(fmaxf(lv, 2.f), 5.f);
    return (int)lv - 2;
}

// grid: n_rois CTAs of 256 threads = 4 bins in flight x 64 channel quads (float4 over channels); every thread walks the
// bins bin0, bin0 + 4, ...  (One CTA per (roi, bin) - 392 000 CTAs of 64 threads for 8000 rois at 7 x 7 - was bound by the
// CTA launch rate: 0.83 ms per call.)  rois: [n][5] = {image, x0, y0, x1, y1}
__global__ void __launch_bounds__(256)
roi_align_kernel(Pyr4 pyr, const float *__restrict__ rois, int C, int pooled, float *__restrict__ out) {
    const int ridx = blockIdx.x;
    const float *r = rois + (size_t)ridx * 5;
    const int img = (int)r[0];
    const int lvl = roi_level(r[1], r[2], r[3], r[4]);
    const float scale = 1.f / (float)(4 << lvl);
    const int H = pyr.h[lvl], W = pyr.w[lvl];
    const float *feat = pyr.p[lvl] + (size_t)img * H * W * C;
    const float rsw = r[1] * scale - 0.5f, rsh = r[2] * scale - 0.5f;
    const float rw = r[3] * scale - 0.5f - rsw, rh = r[4] * scale - 0.5f - rsh;
    const float bh = rh / (float)pooled, bw = rw / (float)pooled;
    const int gh = (int)ceilf(rh / (float)pooled), gw = (int)ceilf(rw / (float)pooled);
    const float count = fmaxf((float)(gh * gw), 1.f);
    const int c = (threadIdx.x & 63) * 4;
    // The sample grid is separable: pooled * gh row positions and pooled * gw column positions per roi.  They are
    // computed ONCE per CTA into shared memory (low index, high index, interpolation weight, validity) with exactly the
    // per-sample expressions of the loop below, instead of by each of the 64 channel threads for each of its samples.
    constexpr int RA_TAB = 256;
    __shared__ int t_lo[2][RA_TAB], t_hi[2][RA_TAB];
    __shared__ float t_l[2][RA_TAB];
    __shared__ unsigned char t_ok[2][RA_TAB];
    const bool tabled = pooled * gh <= RA_TAB && pooled * gw <= RA_TAB;      // block-uniform
    if (tabled) {
        for (int e = threadIdx.x; e < pooled * (gh + gw); e += 256) {
            const int ax = e >= pooled * gh;                                   // 0: rows, 1: columns
            const int k = ax ? e - pooled * gh : e, g = ax ? gw : gh, L = ax ? W : H;
            const int pb = k / g, i = k - pb * g;
            const float b = ax ? bw : bh, st = ax ? rsw : rsh;
            float v = st + pb * b + ((float)i + .5f) * b / (float)g;
            const bool ok = !(v < -1.f || v > (float)L);
            float vv = fmaxf(v, 0.f);
            int lo = (int)vv, hi;
            if (lo >= L - 1) { hi = lo = L - 1; vv = (float)lo; } else hi = lo + 1;
            t_lo[ax][k] = lo; t_hi[ax][k] = hi; t_l[ax][k] = vv - (float)lo; t_ok[ax][k] = ok ? 1 : 0;
        }
        __syncthreads();
        if (c < C)
        for (int bin = threadIdx.x >> 6; bin < pooled * pooled; bin += 4) {
            const int ph = bin / pooled, pw = bin - ph * pooled;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int iy = 0; iy < gh; ++iy) {
                const int ky = ph * gh + iy;
                if (!t_ok[0][ky]) continue;
                const float ly = t_l[0][ky], hy = 1.f - ly;
                const float *row_l = feat + (size_t)t_lo[0][ky] * W * C + c, *row_h = feat + (size_t)t_hi[0][ky] * W * C + c;
                for (int ix = 0; ix < gw; ++ix) {
                    const int kx = pw * gw + ix;
                    if (!t_ok[1][kx]) continue;
                    const float lx = t_l[1][kx], hx = 1.f - lx;
                    const int xl = t_lo[1][kx], xh = t_hi[1][kx];
                    const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                    const float4 v1 = *reinterpret_cast<const float4 *>(row_l + (size_t)xl * C);
                    const float4 v2 = *reinterpret_cast<const float4 *>(row_l + (size_t)xh * C);
                    const float4 v3 = *reinterpret_cast<const float4 *>(row_h + (size_t)xl * C);
                    const float4 v4 = *reinterpret_cast<const float4 *>(row_h + (size_t)xh * C);
                    acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
                    acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
                    acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
                    acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
                }
            }
            acc.x /= count; acc.y /= count; acc.z /= count; acc.w /= count;
            *reinterpret_cast<float4 *>(out + ((size_t)ridx * pooled * pooled + bin) * C + c) = acc;
        }
        return;
    }
    if (c < C)
    for (int bin = threadIdx.x >> 6; bin < pooled * pooled; bin += 4) {
        const int ph = bin / pooled, pw = bin - ph * pooled;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int iy = 0; iy < gh; ++iy) {
            float y = rsh + ph * bh + ((float)iy + .5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ++ix) {
                float x = rsw + pw * bw + ((float)ix + .5f) * bw / (float)gw;
                if (y < -1.f || y > (float)H || x < -1.f || x > (float)W) continue;
                float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
                int yl = (int)yy, xl = (int)xx, yh, xh;
                if (yl >= H - 1) { yh = yl = H - 1; yy = (float)yl; } else yh = yl + 1;
                if (xl >= W - 1) { xh = xl = W - 1; xx = (float)xl; } else xh = xl + 1;
                const float ly = yy - (float)yl, lx = xx - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
                const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                const float4 v1 = *reinterpret_cast<const float4 *>(feat + ((size_t)yl * W + xl) * C + c);
                const float4 v2 = *reinterpret_cast<const float4 *>(feat + ((size_t)yl * W + xh) * C + c);
                const float4 v3 = *reinterpret_cast<const float4 *>(feat + ((size_t)yh * W + xl) * C + c);
                const float4 v4 = *reinterpret_cast<const float4 *>(feat + ((size_t)yh * W + xh) * C + c);
                acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
                acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
                acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
                acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
            }
        }
        acc.x /= count; acc.y /= count; acc.z /= count; acc.w /= count;
        *reinterpret_cast<float4 *>(out + ((size_t)ridx * pooled * pooled + bin) * C + c) = acc;
    }
}

// ------------------------------------------------------------------------------------------------ mask head helpers
// deconv 2x2 stride 2 as a 1x1 conv with 4 C outputs: x [R, H, W, (a, b, c)] -> y [R, 2H, 2W, c]
__global__ void __launch_bounds__(256)
pixel_shuffle2_kernel(const float *__restrict__ x, int R, int H, int W, int C, float *__restrict__ y) {
    const int64_t total = (int64_t)R * H * W * 4 * (C / 4);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int c4 = (int)(i % (C / 4));
        int64_t t = i / (C / 4);
        const int ab = (int)(t % 4); t /= 4;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int r = (int)(t / H);
        const int a = ab >> 1, b = ab & 1;
        const float4 v = *reinterpret_cast<const float4 *>(x + i * 4);
        *reinterpret_cast<float4 *>(y + ((size_t)(r * 2 * H + 2 * h + a) * (2 * W) + 2 * w + b) * C + c4 * 4) = v;
    }
}

// paste: logits [R, M, M, ldl] (channel = class), boxes [R][4] in output-image coordinates, classes [R];
// out bool [R, H, W] = bilinear(sigmoid(logit[class])) >= threshold   (grid_sample, align_corners = False, zero padding)
__global__ void __launch_bounds__(256)
mask_paste_kernel(const float *__restrict__ logits, int ldl, int M, const float *__restrict__ boxes, const int64_t *__restrict__ classes,
                  int R, int H, int W, float threshold, unsigned char *__restrict__ out) {
    const int r = blockIdx.y;
    const float x0 = boxes[r * 4], y0 = boxes[r * 4 + 1], x1 = boxes[r * 4 + 2], y1 = boxes[r * 4 + 3];
    const int cls = (int)classes[r];
    const float *lg = logits + (size_t)r * M * M * ldl + cls;
    auto value = [&](int py, int px) -> unsigned char {
        const float gy = ((float)py + 0.5f - y0) / (y1 - y0) * 2.f - 1.f;
        const float gx = ((float)px + 0.5f - x0) / (x1 - x0) * 2.f - 1.f;
        const float iy = ((gy + 1.f) * (float)M - 1.f) * 0.5f, ix = ((gx + 1.f) * (float)M - 1.f) * 0.5f;
        const float fy = floorf(iy), fx = floorf(ix);
        const int yl = (int)fy, xl = (int)fx;
        const float wy1 = iy - fy, wx1 = ix - fx, wy0 = 1.f - wy1, wx0 = 1.f - wx1;
        float v = 0.f;
        // outside (-1, M): every tap is padding
        if (iy > -1.f && iy < (float)M && ix > -1.f && ix < (float)M) {
            auto tap = [&](int yy, int xx) -> float {
                if (yy < 0 || yy >= M || xx < 0 || xx >= M) return 0.f;
                return 1.f / (1.f + expf(-lg[((size_t)yy * M + xx) * ldl]));
            };
            v = tap(yl, xl) * wy0 * wx0 + tap(yl, xl + 1) * wy0 * wx1 + tap(yl + 1, xl) * wy1 * wx0 + tap(yl + 1, xl + 1) * wy1 * wx1;
        }
        return v >= threshold ? 1 : 0;
    };
    unsigned char *o = out + (size_t)r * H * W;
    if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {          // 4 pixels of a row per thread, one 32-bit store
        for (int q = blockIdx.x * 256 + threadIdx.x; q < H * W / 4; q += gridDim.x * 256) {
            const int pix = q * 4, py = pix / W, px = pix - py * W;
            uchar4 v4;
            v4.x = value(py, px); v4.y = value(py, px + 1); v4.z = value(py, px + 2); v4.w = value(py, px + 3);
            reinterpret_cast<uchar4 *>(o)[q] = v4;
        }
    } else {
        for (int pix = blockIdx.x * 256 + threadIdx.x; pix < H * W; pix += gridDim.x * 256) {
            const int py = pix / W;
            o[pix] = value(py, pix - py * W);
        }
    }
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_rpn_decode(const float *deltas, int ld_deltas, const int64_t *idx, int n_img, int k, int Hl, int Wl, int A,
                               int stride, const float *cell_anchors_h, float img_h, float img_w, float *boxes,
                               unsigned char *valid, void *stream) {
    TTDG_CHECK_ARG(deltas && idx && cell_anchors_h && boxes && valid && A >= 1 && A <= 16 && k >= 0);
    if (n_img * k == 0) return 0;
    CellAnchors ca;
    for (int a = 0; a < A; ++a) for (int c = 0; c < 4; ++c) ca.a[a][c] = cell_anchors_h[a * 4 + c];
    count_launches(1);
    rpn_decode_kernel<<<ceil_div(n_img * k, 256), 256, 0, (cudaStream_t)stream>>>(deltas, ld_deltas, idx, n_img, k, Wl, Hl * Wl, A, stride, ca,
                                                                                 img_h, img_w, logf(1000.f / 16.f), boxes, valid);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_box_predict(const float *cls, int ld_cls, const float *reg, int ld_reg, const float *proposals, int R, int K,
                                float img_h, float img_w, float score_thresh, float *cand_boxes, float *cand_scores, void *stream) {
    TTDG_CHECK_ARG(cls && reg && proposals && cand_boxes && cand_scores && R >= 0 && K >= 1);
    if (R == 0) return 0;
    count_launches(1);
    box_predict_kernel<<<ceil_div(R * K, 256), 256, 0, (cudaStream_t)stream>>>(cls, ld_cls, reg, ld_reg, proposals, R, K, img_h, img_w,
                                                                              score_thresh, logf(1000.f / 16.f), cand_boxes, cand_scores);
    TTDG_LAUNCH_RET();
}

extern "C" int64_t ttdg_nms_scratch_bytes(int batch, int n) { return (int64_t)batch * n * ((n + 63) / 64) * 8; }

extern "C" int ttdg_nms(const float *boxes_sorted, const int32_t *category, int batch, int n, float iou_thresh, int max_keep,
                        int32_t *keep, int32_t *n_keep, void *scratch, void *stream) {
    TTDG_CHECK_ARG(boxes_sorted && category && keep && n_keep && scratch && n >= 0 && max_keep >= 1 && batch >= 1 && batch <= 65535);
    if (n == 0) return (int)cudaMemsetAsync(n_keep, 0, sizeof(int32_t) * batch, (cudaStream_t)stream);
    const int words = (n + 63) / 64;
    if ((size_t)words * 8 > 200 * 1024) return TTDG_E_LIMIT;
    {   // boxes + categories + suppression bitmask fit in shared memory (n <= ~11 000): the fused one-CTA-per-image kernel
        const size_t smem = (size_t)n * 20 + (size_t)words * 8;
        static int fused = -1;
        if (fused < 0) { const char *e = getenv("TTDG_NMS_FUSED"); fused = (e && e[0] == '0') ? 0 : 1; }
        // measured on B200: 84 vs 103 us at n = 2000 (box head), but 1.34 vs 0.96 ms at n = 8960 (RPN: ~all 140 rounds run,
        // and 8 CTAs cannot match the whole GPU evaluating the triangle) - fused only for the small problems
        if (fused && n <= 2560 && smem <= 220 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(nms_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            count_launches(1);
            nms_fused_kernel<<<batch, NMSF_THREADS, smem, (cudaStream_t)stream>>>(boxes_sorted, category, n, iou_thresh, max_keep, keep, n_keep);
            TTDG_LAUNCH_RET();
        }
    }
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(scratch);
    cudaError_t e = cudaMemsetAsync(mask, 0, (size_t)batch * n * words * 8, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    count_launches(2);
    nms_mask_kernel<<<dim3(words, words, batch), 64, 0, (cudaStream_t)stream>>>(boxes_sorted, category, n, iou_thresh, mask, words);
    e = cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 8);
    if (e != cudaSuccess) return (int)e;
    nms_sweep_kernel<<<batch, 256, words * 8, (cudaStream_t)stream>>>(mask, n, words, max_keep, keep, n_keep);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_roi_align(const float *const *feat_ptrs_h, const int32_t *lvl_hw_h, const float *rois, int n_rois, int C,
                              int pooled, float *out, void *stream) {
    TTDG_CHECK_ARG(feat_ptrs_h && lvl_hw_h && rois && out && n_rois >= 0 && C % 4 == 0 && C <= 256 && pooled >= 1);
    if (n_rois == 0) return 0;
    Pyr4 pyr;
    for (int l = 0; l < 4; ++l) { pyr.p[l] = feat_ptrs_h[l]; pyr.h[l] = lvl_hw_h[2 * l]; pyr.w[l] = lvl_hw_h[2 * l + 1]; }
    count_launches(1);
    roi_align_kernel<<<n_rois, 256, 0, (cudaStream_t)stream>>>(pyr, rois, C, pooled, out);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_pixel_shuffle2(const float *x, int R, int H, int W, int C, float *y, void *stream) {
    TTDG_CHECK_ARG(x && y && C % 4 == 0);
    const int64_t total = (int64_t)R * H * W * C;
    if (total == 0) return 0;
    int64_t nb = (total + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    count_launches(1);
    pixel_shuffle2_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(x, R, H, W, C, y);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_mask_paste(const float *logits, int ld_logits, int M, const float *boxes, const int64_t *classes, int R, int H,
                               int W, float threshold, unsigned char *out, void *stream) {
    TTDG_CHECK_ARG(logits && boxes && classes && out && R >= 0 && M >= 1);
    if (R == 0) return 0;
    count_launches(1);
    int bx = ceil_div(H * W, 256);
    if (bx > 64) bx = 64;
    mask_paste_kernel<<<dim3(bx, R), 256, 0, (cudaStream_t)stream>>>(logits, ld_logits, M, boxes, classes, R, H, W, threshold, out);
    TTDG_LAUNCH_RET();
}
