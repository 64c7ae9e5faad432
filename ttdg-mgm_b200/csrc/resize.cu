// resize.cu - the image resize of the test data path on the device: Detectron2 ResizeShortestEdge -> ResizeTransform.apply_image
// on uint8 images = PIL.Image.resize((w, h), BILINEAR) (reference adapteacher/data/build.py:122-154 -> d2 DatasetMapper(cfg, False);
// "GPU-side resize").  Bit-exact with Pillow (tests/test_resize.py): Pillow's 8-bit resampler is a separable
// triangle filter whose support grows with the down-scaling factor, with coefficients in 22-bit fixed point and an int32
// accumulator per pass - integer work that maps 1 : 1 onto a kernel.  The coefficient tables (a few KB) are built on the HOST in
// double precision with Pillow's exact expressions (ttdg_resize_coeffs_u8, no device involved) and handed to the two passes.
#include "common.cuh"
#include <cmath>

namespace ttdg {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ unsigned char rs_clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src H x W x C (interleaved) -> dst H x nw x C; one thread per output pixel, C <= 4 accumulators
template <int C>
__global__ void __launch_bounds__(256)
resize_h_kernel(const unsigned char *__restrict__ src, int H, int W, const int32_t *__restrict__ bounds, const int32_t *__restrict__ kk,
                int ksize, int nw, unsigned char *__restrict__ dst) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)H * nw) return;
    const int y = (int)(t / nw), xx = (int)(t - (long long)y * nw);
    const int lo = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const int32_t *k = kk + (size_t)xx * ksize;
    const unsigned char *row = src + ((size_t)y * W + lo) * C;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (RS_PRECISION_BITS - 1);
    for (int x = 0; x < cnt; ++x) {
        const int w = k[x];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += (int)row[x * C + c] * w;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dst[((size_t)y * nw + xx) * C + c] = rs_clip8(acc[c]);
}

// vertical pass: src H x W x C -> dst nh x W, written either interleaved (planar == 0) or as C planes nh x W (planar != 0: the
// uint8 C x H x W tensor the detector's preprocess reads), optionally with the channel order reversed (INPUT.FORMAT BGR)
template <int C>
__global__ void __launch_bounds__(256)
resize_v_kernel(const unsigned char *__restrict__ src, int H, int W, const int32_t *__restrict__ bounds, const int32_t *__restrict__ kk,
                int ksize, int nh, int planar, int flip, unsigned char *__restrict__ dst) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)nh * W) return;
    const int yy = (int)(t / W), x = (int)(t - (long long)yy * W);
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (RS_PRECISION_BITS - 1);
    if (bounds) {
        const int lo = bounds[2 * yy], cnt = bounds[2 * yy + 1];
        const int32_t *k = kk + (size_t)yy * ksize;
        for (int y = 0; y < cnt; ++y) {
            const int w = k[y];
            const unsigned char *px = src + ((size_t)(lo + y) * W + x) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] += (int)px[c] * w;
        }
    } else {                                             // no vertical resize: a pure layout pass (copy)
        const unsigned char *px = src + ((size_t)yy * W + x) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = (int)px[c] << RS_PRECISION_BITS;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int co = flip ? C - 1 - c : c;
        const unsigned char v = rs_clip8(acc[c]);
        if (planar) dst[((size_t)co * nh + yy) * W + x] = v;
        else dst[((size_t)yy * W + x) * C + co] = v;
    }
}

static double rs_triangle(double x) {
    if (x < 0.0) x = -x;
    return x < 1.0 ? 1.0 - x : 0.0;
}

}  // namespace ttdg

using namespace ttdg;

// number of coefficients per output sample: Pillow's ksize = ceil(support) * 2 + 1 with support = max(in / out, 1)
extern "C" int ttdg_resize_ksize(int in_size, int out_size) {
    if (in_size < 1 || out_size < 1) return TTDG_E_ARG;
    double filterscale = (double)in_size / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    return (int)ceil(1.0 * filterscale) * 2 + 1;
}

// HOST function (no device work): Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter over the whole axis.
// bounds_h: 2 * out_size int32 (first input index, count); kk_h: out_size * ksize int32 fixed-point weights (ksize from
// ttdg_resize_ksize).  Returns ksize.
extern "C" int ttdg_resize_coeffs_u8(int in_size, int out_size, int32_t *bounds_h, int32_t *kk_h) {
    TTDG_CHECK_ARG(bounds_h && kk_h && in_size >= 1 && out_size >= 1);
    const double scale = (double)in_size / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double kd[64];
        double *k = ksize <= 64 ? kd : new double[ksize];
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            const double w = rs_triangle((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        int32_t *out = kk_h + (size_t)xx * ksize;
        for (int x = 0; x < ksize; ++x) {
            double v = 0.0;
            if (x < xmax) v = ww != 0.0 ? k[x] / ww : k[x];
            out[x] = v < 0 ? (int32_t)(-0.5 + v * (double)(1 << RS_PRECISION_BITS)) : (int32_t)(0.5 + v * (double)(1 << RS_PRECISION_BITS));
        }
        if (k != kd) delete[] k;
        bounds_h[2 * xx] = xmin;
        bounds_h[2 * xx + 1] = xmax;
    }
    return ksize;
}

// src: H x W x C uint8 (interleaved, C = 3 or 4 or 1) on the device -> dst: nh x nw, interleaved (planar = 0) or C planes
// (planar = 1), channel order reversed when flip.  bounds_x / kk_x (device copies of ttdg_resize_coeffs_u8(W, nw)) are ignored
// when nw == W, bounds_y / kk_y when nh == H (Pillow skips the pass).  tmp: H x nw x C bytes of scratch (unused when nw == W).
extern "C" int ttdg_resize_bilinear_u8(const unsigned char *src, int H, int W, int C, const int32_t *bounds_x, const int32_t *kk_x, int ksize_x,
                                       const int32_t *bounds_y, const int32_t *kk_y, int ksize_y, int nh, int nw, unsigned char *tmp,
                                       unsigned char *dst, int planar, int flip, void *stream) {
    TTDG_CHECK_ARG(src && dst && H >= 1 && W >= 1 && nh >= 1 && nw >= 1 && (C == 1 || C == 3 || C == 4));
    TTDG_CHECK_ARG(nw == W || (bounds_x && kk_x && tmp && ksize_x >= 1));
    TTDG_CHECK_ARG(nh == H || (bounds_y && kk_y && ksize_y >= 1));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned char *cur = src;
    if (nw != W) {
        const long long n = (long long)H * nw;
        const unsigned grid = (unsigned)((n + 255) / 256);
        count_launches(1);
        if (C == 3) resize_h_kernel<3><<<grid, 256, 0, st>>>(src, H, W, bounds_x, kk_x, ksize_x, nw, tmp);
        else if (C == 4) resize_h_kernel<4><<<grid, 256, 0, st>>>(src, H, W, bounds_x, kk_x, ksize_x, nw, tmp);
        else resize_h_kernel<1><<<grid, 256, 0, st>>>(src, H, W, bounds_x, kk_x, ksize_x, nw, tmp);
        cur = tmp;
    }
    const long long n = (long long)nh * nw;
    const unsigned grid = (unsigned)((n + 255) / 256);
    const int32_t *by = nh != H ? bounds_y : nullptr;
    count_launches(1);
    if (C == 3) resize_v_kernel<3><<<grid, 256, 0, st>>>(cur, H, nw, by, kk_y, ksize_y, nh, planar, flip, dst);
    else if (C == 4) resize_v_kernel<4><<<grid, 256, 0, st>>>(cur, H, nw, by, kk_y, ksize_y, nh, planar, flip, dst);
    else resize_v_kernel<1><<<grid, 256, 0, st>>>(cur, H, nw, by, kk_y, ksize_y, nh, planar, flip, dst);
    TTDG_LAUNCH_RET();
}
