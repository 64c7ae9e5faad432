// conv.cu - NHWC fp32 implicit-GEMM convolution family of the detector (forward, data gradient, weight gradient).
//
// Replaces the cuDNN / cuBLAS calls behind Detectron2's ResNet-50-FPN, RPN head, box FCs and mask head as driven by
// adapteacher/modeling/meta_arch/rcnn.py:219-226, proposal_generator/rpn.py:27, roi_heads/roi_heads.py:182-184,112
// (SURVEY 2.3 K1-K3, K6, K7, K17).  Round-1 version: CUDA-core fp32 FMA (exact fp32 products, the precision the
// 1e-4 mIoU parity gate of BASELINE.json configs[1] needs); the tcgen05 / TMA version of the same tiling is the next
// step.
//
// One GEMM core, three operand gathers:
//   FWD    C[m = (img, ho, wo)][n = cout] = sum_{k = (r, s, cin)} X[img, ho*st + r - pad, wo*st + s - pad, cin] * W[r][s][cin][cout]
//   DGRAD  the same gather on dY with B[k = (r, s, cout)][n = cin] = W[R-1-r][S-1-s][cin][cout]      (stride-1 convs;
//          1x1 stride-2 convs use it with an output scatter stride of 2 into a zero-filled dX)
//   WGRAD  C[m = (r, s, cin)][n = cout] += sum_{k = pixel} X[pixel shifted by (r, s)][cin] * dY[pixel][cout]  (split over pixels, atomics)
// Epilogue (FWD/DGRAD): v = acc * scale[n] + bias[n] (+ residual[same pixel] | + residual[(ho/2, wo/2)], the FPN
// top-down nearest upsample) -> optional ReLU -> Y[(img, ho*os, wo*os)][n].
// Weights are stored [R][S][Cin][Cout] (the host wrapper permutes torch's [Cout][Cin][R][S]).
#include "common.cuh"
#include <cuda_bf16.h>

namespace ttdg {

constexpr int CBM = 128, CBN = 128, CBK = 16, CTHREADS = 256;

struct ConvParams {
    const float *X;      // FWD: input NHWC; DGRAD: dY; WGRAD: input X
    const float *W;      // FWD/DGRAD: weights [R][S][Cin][Cout]; WGRAD: dY (pixels x Cout)
    float *Y;            // FWD/DGRAD: output; WGRAD: dW [R][S][Cin][Cout]
    const float *scale, *bias, *residual;
    int N, H, W_, Cin;   // gathered tensor: N x H x W_ x Cin (channels of the gathered operand)
    int Ho, Wo, Cout;    // GEMM pixel grid (Ho x Wo per image) and GEMM n extent
    int R, S, stride, pad;
    int os, Hy, Wy;      // output scatter stride: pixel (ho, wo) is stored at (ho, wo) * os of the N x Hy x Wy x Cout map Y
    int res_mode;        // 0 none, 1 same pixel, 2 nearest-upsampled from (Ho/2 x Wo/2)
    int relu;
    int wR, wS, wCin, wCout;   // real weight dims for DGRAD's transposed read
    int splits;          // WGRAD: pixel-range splits (gridDim.z)
};

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2, MODE_FWD_GENERIC = 3 };

__device__ __forceinline__ void mma_tile(const float (*As)[CBM], const float (*Bs)[CBN], float (&acc)[8][8], int tm, int tn) {
#pragma unroll
    for (int k = 0; k < CBK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][tm]);
        const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][tm + 64]);
        const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][tn]);
        const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[k][tn + 64]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

template <int MODE>
__global__ void __launch_bounds__(CTHREADS, 2)
conv_gemm_kernel(const ConvParams p) {
    __shared__ __align__(16) float As[CBK][CBM];
    __shared__ __align__(16) float Bs[CBK][CBN];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * CBM, n0 = blockIdx.y * CBN;
    const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    if (MODE == MODE_WGRAD) {
        // ---------------- C[(r,s,ci)][co] += sum_pixels ; m-tile lies inside one (r, s) because Cin % 128 == 0
        const int M = p.R * p.S * p.Cin;
        const int rs = m0 / p.Cin, ci0 = m0 - rs * p.Cin;
        const int r = rs / p.S, s = rs - r * p.S;
        const int P = p.N * p.Ho * p.Wo;
        const int per = (P + p.splits - 1) / p.splits;
        const int pbeg = blockIdx.z * per, pend = min(P, pbeg + per);
        const int lk = tid >> 5, lq = (tid & 31) * 4;           // A: pixel lk (+8), float4 along m; B: same along n
        for (int p0 = pbeg; p0 < pend; p0 += CBK) {
            float4 av[2], bv[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int pix = p0 + lk + 8 * h;
                av[h] = make_float4(0.f, 0.f, 0.f, 0.f);
                bv[h] = av[h];
                if (pix < pend) {
                    const int img = pix / (p.Ho * p.Wo), rem = pix - img * p.Ho * p.Wo;
                    const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
                    const int hi = ho * p.stride + r - p.pad, wi = wo * p.stride + s - p.pad;
                    if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W_)
                        av[h] = *reinterpret_cast<const float4 *>(p.X + ((size_t)(img * p.H + hi) * p.W_ + wi) * p.Cin + ci0 + lq);
                    if (n0 + lq < p.Cout) bv[h] = *reinterpret_cast<const float4 *>(p.W + (size_t)pix * p.Cout + n0 + lq);
                }
            }
            __syncthreads();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                *reinterpret_cast<float4 *>(&As[lk + 8 * h][lq]) = av[h];
                *reinterpret_cast<float4 *>(&Bs[lk + 8 * h][lq]) = bv[h];
            }
            __syncthreads();
            mma_tile(As, Bs, acc, tm, tn);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + tm + (i & 3) + (i >> 2) * 64;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + tn + (j & 3) + (j >> 2) * 64;
                if (n < p.Cout) atomicAdd(p.Y + (size_t)m * p.Cout + n, acc[i][j]);
            }
        }
        return;
    }

    // ---------------- FWD / DGRAD / FWD_GENERIC
    const int Mtot = p.N * p.Ho * p.Wo;
    const int K = p.R * p.S * p.Cin;
    // A loader: one GEMM row per thread pair, 8 consecutive k each
    const int arow = tid >> 1, akq = (tid & 1) * 8;
    const int am = m0 + arow;
    int a_img = 0, a_ho = 0, a_wo = 0;
    const bool a_ok = am < Mtot;
    if (a_ok) { a_img = am / (p.Ho * p.Wo); const int rem = am - a_img * p.Ho * p.Wo; a_ho = rem / p.Wo; a_wo = rem - a_ho * p.Wo; }
    // B loader
    const int bk = tid >> 5, bq = (tid & 31) * 4;                // FWD: k = bk (+8), float4 along n
    const int brow = tid >> 1, bkq = (tid & 1) * 8;              // DGRAD: n = brow, 8 consecutive k

    for (int k0 = 0; k0 < K; k0 += CBK) {
        float a_reg[8], b_reg[8];
        if (MODE == MODE_FWD_GENERIC) {
            // arbitrary Cin (the 7x7 stem, Cin = 3): decode (r, s, c) per element
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = k0 + akq + e;
                float v = 0.f;
                if (a_ok && k < K) {
                    const int rs = k / p.Cin, c = k - rs * p.Cin;
                    const int r = rs / p.S, s = rs - r * p.S;
                    const int hi = a_ho * p.stride + r - p.pad, wi = a_wo * p.stride + s - p.pad;
                    if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W_) v = p.X[((size_t)(a_img * p.H + hi) * p.W_ + wi) * p.Cin + c];
                }
                a_reg[e] = v;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = k0 + bk + 8 * h;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < K && n0 + bq < p.Cout) v = *reinterpret_cast<const float4 *>(p.W + (size_t)k * p.Cout + n0 + bq);
                b_reg[4 * h] = v.x; b_reg[4 * h + 1] = v.y; b_reg[4 * h + 2] = v.z; b_reg[4 * h + 3] = v.w;
            }
        } else {
            // Cin % 16 == 0: the whole k-tile shares one (r, s)
            const int rs = k0 / p.Cin, c0 = k0 - rs * p.Cin;
            const int r = rs / p.S, s = rs - r * p.S;
            const int hi = a_ho * p.stride + r - p.pad, wi = a_wo * p.stride + s - p.pad;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (a_ok && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W_) {
                const float *src = p.X + ((size_t)(a_img * p.H + hi) * p.W_ + wi) * p.Cin + c0 + akq;
                v0 = *reinterpret_cast<const float4 *>(src);
                v1 = *reinterpret_cast<const float4 *>(src + 4);
            }
            a_reg[0] = v0.x; a_reg[1] = v0.y; a_reg[2] = v0.z; a_reg[3] = v0.w;
            a_reg[4] = v1.x; a_reg[5] = v1.y; a_reg[6] = v1.z; a_reg[7] = v1.w;
            if (MODE == MODE_FWD) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n0 + bq < p.Cout) v = *reinterpret_cast<const float4 *>(p.W + (size_t)(k0 + bk + 8 * h) * p.Cout + n0 + bq);
                    b_reg[4 * h] = v.x; b_reg[4 * h + 1] = v.y; b_reg[4 * h + 2] = v.z; b_reg[4 * h + 3] = v.w;
                }
            } else {   // DGRAD: B[k = (r, s, co)][n = ci] = W[wR-1-r][wS-1-s][ci][co], contiguous along co
                float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
                const int n = n0 + brow;
                if (n < p.Cout) {
                    const float *src = p.W + ((size_t)((p.wR - 1 - r) * p.wS + (p.wS - 1 - s)) * p.wCin + n) * p.wCout + c0 + bkq;
                    w0 = *reinterpret_cast<const float4 *>(src);
                    w1 = *reinterpret_cast<const float4 *>(src + 4);
                }
                b_reg[0] = w0.x; b_reg[1] = w0.y; b_reg[2] = w0.z; b_reg[3] = w0.w;
                b_reg[4] = w1.x; b_reg[5] = w1.y; b_reg[6] = w1.z; b_reg[7] = w1.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; ++e) As[akq + e][arow] = a_reg[e];
        if (MODE == MODE_DGRAD) {
#pragma unroll
            for (int e = 0; e < 8; ++e) Bs[bkq + e][brow] = b_reg[e];
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                *reinterpret_cast<float4 *>(&Bs[bk + 8 * h][bq]) = make_float4(b_reg[4 * h], b_reg[4 * h + 1], b_reg[4 * h + 2], b_reg[4 * h + 3]);
        }
        __syncthreads();
        mma_tile(As, Bs, acc, tm, tn);
    }

    // ---------------- epilogue
    const int Hy = p.Hy, Wy = p.Wy;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + tm + (i & 3) + (i >> 2) * 64;
        if (m >= Mtot) continue;
        const int img = m / (p.Ho * p.Wo), rem = m - img * p.Ho * p.Wo;
        const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
        float *yrow = p.Y + ((size_t)(img * Hy + ho * p.os) * Wy + wo * p.os) * p.Cout;
        const float *rrow = nullptr;
        if (p.res_mode == 1) rrow = p.residual + (size_t)m * p.Cout;
        else if (p.res_mode == 2) rrow = p.residual + ((size_t)(img * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1)) * p.Cout;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + tn + jh * 64;
            if (n >= p.Cout) continue;
            float4 v = make_float4(acc[i][4 * jh], acc[i][4 * jh + 1], acc[i][4 * jh + 2], acc[i][4 * jh + 3]);
            if (p.scale) {
                const float4 sc = *reinterpret_cast<const float4 *>(p.scale + n);
                v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
            }
            if (p.bias) {
                const float4 b = *reinterpret_cast<const float4 *>(p.bias + n);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (rrow) {
                const float4 rr = *reinterpret_cast<const float4 *>(rrow + n);
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4 *>(yrow + n) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ small NHWC helpers
// g_out = (y > 0 ? g : 0) * scale[c]   (ReLU mask from the stored output, FrozenBN scale); g_out may alias g
__global__ void __launch_bounds__(256)
relu_bn_bwd_kernel(const float *__restrict__ g, const float *__restrict__ y, const float *__restrict__ scale, int C,
                   int64_t n4, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 gv = reinterpret_cast<const float4 *>(g)[i];
        if (y) {
            const float4 yv = reinterpret_cast<const float4 *>(y)[i];
            gv.x = yv.x > 0.f ? gv.x : 0.f; gv.y = yv.y > 0.f ? gv.y : 0.f; gv.z = yv.z > 0.f ? gv.z : 0.f; gv.w = yv.w > 0.f ? gv.w : 0.f;
        }
        if (scale) {
            const float4 sc = *reinterpret_cast<const float4 *>(scale + (int)((i * 4) % C));
            gv.x *= sc.x; gv.y *= sc.y; gv.z *= sc.z; gv.w *= sc.w;
        }
        reinterpret_cast<float4 *>(out)[i] = gv;
    }
}

// out[c] += sum over pixels of g[pixel][c]   (bias gradient); grid (C / 32, pixel chunks), one atomic per (CTA, channel)
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float *__restrict__ g, int64_t P, int C, float *__restrict__ out) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), pr = threadIdx.x >> 5;
    const int64_t per = (P + gridDim.y - 1) / gridDim.y;
    const int64_t p0 = (int64_t)blockIdx.y * per, p1 = min(P, p0 + per);
    float s = 0.f;
    if (c < C) for (int64_t p = p0 + pr; p < p1; p += 8) s += g[p * C + c];
    red[pr][threadIdx.x & 31] = s;
    __syncthreads();
    if (pr == 0 && c < C) {
        float t = 0.f;
        for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

// 3x3 stride-2 pad-1 max pooling, NHWC (ResNet stem)
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const float *__restrict__ x, int N, int H, int W, int C, int Ho, int Wo, float *__restrict__ y) {
    const int64_t total = (int64_t)N * Ho * Wo * (C / 4);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int c4 = (int)(i % (C / 4));
        int64_t t = i / (C / 4);
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int r = 0; r < 3; ++r) {
            const int hi = ho * 2 + r - 1;
            if (hi < 0 || hi >= H) continue;
            for (int s = 0; s < 3; ++s) {
                const int wi = wo * 2 + s - 1;
                if (wi < 0 || wi >= W) continue;
                const float4 v = *reinterpret_cast<const float4 *>(x + ((size_t)(n * H + hi) * W + wi) * C + c4 * 4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        *reinterpret_cast<float4 *>(y + i * 4) = m;
    }
}

// mode 0: y[n, ho, wo] = x[n, 2ho, 2wo]            (p6 = max_pool2d(p5, kernel 1, stride 2))
// mode 1: y[n, 2ho, 2wo] += x[n, ho, wo]           (its backward, into the p5 gradient)
// mode 2: y[n, ho, wo] += sum of the 2x2 children x[n, 2ho + a, 2wo + b]   (backward of the FPN nearest upsample)
__global__ void __launch_bounds__(256)
resample2_kernel(const float *__restrict__ x, float *__restrict__ y, int N, int Hs, int Ws, int C, int mode) {
    // Hs x Ws = the SMALL grid; the large grid is 2Hs x 2Ws
    const int64_t total = (int64_t)N * Hs * Ws * (C / 4);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int c4 = (int)(i % (C / 4));
        int64_t t = i / (C / 4);
        const int ws = (int)(t % Ws); t /= Ws;
        const int hs = (int)(t % Hs);
        const int n = (int)(t / Hs);
        const size_t small = ((size_t)(n * Hs + hs) * Ws + ws) * C + c4 * 4;
        const size_t big = ((size_t)(n * 2 * Hs + 2 * hs) * (2 * Ws) + 2 * ws) * C + c4 * 4;
        if (mode == 0) {
            *reinterpret_cast<float4 *>(y + small) = *reinterpret_cast<const float4 *>(x + big);
        } else if (mode == 1) {
            float4 a = *reinterpret_cast<float4 *>(y + big);
            const float4 b = *reinterpret_cast<const float4 *>(x + small);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            *reinterpret_cast<float4 *>(y + big) = a;
        } else {
            float4 a = *reinterpret_cast<float4 *>(y + small);
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    const float4 b = *reinterpret_cast<const float4 *>(x + big + ((size_t)dy * 2 * Ws + dx) * C);
                    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                }
            *reinterpret_cast<float4 *>(y + small) = a;
        }
    }
}

// uint8 N x 3 x H x W (planar, as the dataset mapper delivers) -> fp32 NHWC padded to 4 channels, minus the pixel mean.
// The output rows hold Wp >= left + W pixels: `left` zero pixels, the image, zeros up to Wp (the zero padding of the
// stem convolution materialised, so that the tensor-core stem can read 8-pixel windows as plain TMA boxes).
__global__ void __launch_bounds__(256)
preprocess_kernel(const unsigned char *__restrict__ img, int N, int H, int W, int Wp, int left, float m0, float m1, float m2,
                  float *__restrict__ out) {
    const int64_t total = (int64_t)N * H * Wp;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t row = i / Wp;                       // n * H + h
        const int col = (int)(i - row * Wp) - left;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col >= 0 && col < W) {
            const int64_t n = row / H, h = row - n * H;
            const unsigned char *b = img + n * 3 * H * W + h * W + col;
            v.x = (float)b[0] - m0; v.y = (float)b[(int64_t)H * W] - m1; v.z = (float)b[2 * (int64_t)H * W] - m2;
        }
        reinterpret_cast<float4 *>(out)[i] = v;
    }
}

static int launch_conv(int mode, const ConvParams &p, cudaStream_t st) {
    if (mode == MODE_WGRAD) {
        const int M = p.R * p.S * p.Cin;
        dim3 grid(ceil_div(M, CBM), ceil_div(p.Cout, CBN), p.splits);
        count_launches(1);
        conv_gemm_kernel<MODE_WGRAD><<<grid, CTHREADS, 0, st>>>(p);
    } else {
        const int Mtot = p.N * p.Ho * p.Wo;
        dim3 grid(ceil_div(Mtot, CBM), ceil_div(p.Cout, CBN));
        count_launches(1);
        if (mode == MODE_FWD) conv_gemm_kernel<MODE_FWD><<<grid, CTHREADS, 0, st>>>(p);
        else if (mode == MODE_DGRAD) conv_gemm_kernel<MODE_DGRAD><<<grid, CTHREADS, 0, st>>>(p);
        else conv_gemm_kernel<MODE_FWD_GENERIC><<<grid, CTHREADS, 0, st>>>(p);
    }
    return (int)cudaGetLastError();
}

}  // namespace ttdg

using namespace ttdg;

extern "C" int ttdg_conv_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual,
                             int res_mode, int relu, int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad,
                             float *y, void *stream) {
    TTDG_CHECK_ARG(x && w && y && N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0);
    TTDG_CHECK_ARG(Cout % 4 == 0 && (res_mode == 0 || residual));
    if (N == 0) return 0;
    ConvParams p = {};
    p.X = x; p.W = w; p.Y = y; p.scale = scale; p.bias = bias; p.residual = residual;
    p.N = N; p.H = H; p.W_ = W; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.stride = stride; p.pad = pad;
    p.Ho = (H + 2 * pad - R) / stride + 1; p.Wo = (W + 2 * pad - S) / stride + 1;
    p.os = 1; p.Hy = p.Ho; p.Wy = p.Wo; p.res_mode = res_mode; p.relu = relu;
    if (res_mode == 2 && ((p.Ho | p.Wo) & 1)) return TTDG_E_ARG;
    return launch_conv(Cin % 16 == 0 ? MODE_FWD : MODE_FWD_GENERIC, p, (cudaStream_t)stream);
}

extern "C" int ttdg_conv_dgrad(const float *dy, const float *w, int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                               int pad, float *dx, void *stream) {
    // (N, H, W, Cin) = the forward INPUT geometry; dy is N x Ho x Wo x Cout; dx must be zero-filled when stride == 2
    TTDG_CHECK_ARG(dy && w && dx && N >= 0 && Cin % 4 == 0 && Cout % 16 == 0);
    if (!((stride == 1) || (stride == 2 && R == 1 && S == 1 && pad == 0))) return TTDG_E_LIMIT;
    if (N == 0) return 0;
    ConvParams p = {};
    const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
    p.X = dy; p.W = w; p.Y = dx;
    p.N = N; p.H = Ho; p.W_ = Wo; p.Cin = Cout;          // gathered operand = dY with Cout channels
    p.Cout = Cin;                                        // GEMM n = cin
    p.R = R; p.S = S; p.stride = 1; p.pad = R - 1 - pad;
    p.Ho = stride == 1 ? H : Ho; p.Wo = stride == 1 ? W : Wo;
    p.os = stride; p.Hy = H; p.Wy = W;                   // odd H / W: the last input row / column gets no gradient
    p.wR = R; p.wS = S; p.wCin = Cin; p.wCout = Cout;
    return launch_conv(MODE_DGRAD, p, (cudaStream_t)stream);
}

extern "C" int ttdg_conv_wgrad(const float *x, const float *dy, int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                               int pad, float *dw, void *stream) {
    // dw [R][S][Cin][Cout] is ACCUMULATED into (atomics over pixel splits): the caller's gradient bucket
    TTDG_CHECK_ARG(x && dy && dw && N >= 0 && Cin % 128 == 0 && Cout % 4 == 0);
    if (N == 0) return 0;
    ConvParams p = {};
    p.X = x; p.W = dy; p.Y = dw;
    p.N = N; p.H = H; p.W_ = W; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.stride = stride; p.pad = pad;
    p.Ho = (H + 2 * pad - R) / stride + 1; p.Wo = (W + 2 * pad - S) / stride + 1;
    const int tiles = ceil_div(R * S * Cin, CBM) * ceil_div(Cout, CBN);
    const int P = N * p.Ho * p.Wo;
    int splits = (148 * 4 + tiles - 1) / tiles;          // fill the machine ~4 CTAs deep
    const int max_splits = (P + 255) / 256;              // at least 256 pixels per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.splits = splits;
    return launch_conv(MODE_WGRAD, p, (cudaStream_t)stream);
}

extern "C" int ttdg_relu_bn_bwd(const float *g, const float *y, const float *scale, int C, int64_t numel, float *out, void *stream) {
    TTDG_CHECK_ARG(g && out && C % 4 == 0 && numel % 4 == 0 && numel >= 0);
    if (numel == 0) return 0;
    int64_t nb = (numel / 4 + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    count_launches(1);
    relu_bn_bwd_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(g, y, scale, C, numel / 4, out);
    TTDG_LAUNCH_RET();
}

namespace ttdg {
// one pass, two results: out_pre = (y > 0 ? g : 0) (the gradient of the residual branch), out_conv = out_pre * scale[c]
__global__ void __launch_bounds__(256)
relu_bn_bwd2_kernel(const float *__restrict__ g, const float *__restrict__ y, const float *__restrict__ scale, int C, int64_t n4,
                    float *__restrict__ out_pre, float *__restrict__ out_conv) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 gv = reinterpret_cast<const float4 *>(g)[i];
        const float4 yv = reinterpret_cast<const float4 *>(y)[i];
        gv.x = yv.x > 0.f ? gv.x : 0.f; gv.y = yv.y > 0.f ? gv.y : 0.f; gv.z = yv.z > 0.f ? gv.z : 0.f; gv.w = yv.w > 0.f ? gv.w : 0.f;
        reinterpret_cast<float4 *>(out_pre)[i] = gv;
        const float4 sc = *reinterpret_cast<const float4 *>(scale + (int)((i * 4) % C));
        gv.x *= sc.x; gv.y *= sc.y; gv.z *= sc.z; gv.w *= sc.w;
        reinterpret_cast<float4 *>(out_conv)[i] = gv;
    }
}
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 v) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&v.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
// bf16 backbone: the ReLU mask comes from the stored bf16 output, gradients stay fp32
__global__ void __launch_bounds__(256)
relu_bn_bwd_bf16y_kernel(const float *__restrict__ g, const __nv_bfloat16 *__restrict__ y, const float *__restrict__ scale, int C,
                         int64_t n4, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 gv = reinterpret_cast<const float4 *>(g)[i];
        const float4 yv = bf16x4_to_float4(reinterpret_cast<const uint2 *>(y)[i]);
        gv.x = yv.x > 0.f ? gv.x : 0.f; gv.y = yv.y > 0.f ? gv.y : 0.f; gv.z = yv.z > 0.f ? gv.z : 0.f; gv.w = yv.w > 0.f ? gv.w : 0.f;
        if (scale) {
            const float4 sc = *reinterpret_cast<const float4 *>(scale + (int)((i * 4) % C));
            gv.x *= sc.x; gv.y *= sc.y; gv.z *= sc.z; gv.w *= sc.w;
        }
        reinterpret_cast<float4 *>(out)[i] = gv;
    }
}
// 3x3 stride-2 pad-1 max pooling on bf16 NHWC maps (the bf16 backbone's stem)
__global__ void __launch_bounds__(256)
maxpool3x3s2_bf16_kernel(const __nv_bfloat16 *__restrict__ x, int N, int H, int W, int C, int Ho, int Wo, __nv_bfloat16 *__restrict__ y) {
    const int64_t total = (int64_t)N * Ho * Wo * (C / 4);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int c4 = (int)(i % (C / 4));
        int64_t t = i / (C / 4);
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int r = 0; r < 3; ++r) {
            const int hi = ho * 2 + r - 1;
            if (hi < 0 || hi >= H) continue;
            for (int s = 0; s < 3; ++s) {
                const int wi = wo * 2 + s - 1;
                if (wi < 0 || wi >= W) continue;
                const float4 v = bf16x4_to_float4(*reinterpret_cast<const uint2 *>(x + ((size_t)(n * H + hi) * W + wi) * C + c4 * 4));
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        const __nv_bfloat162 lo2 = __floats2bfloat162_rn(m.x, m.y), hi2 = __floats2bfloat162_rn(m.z, m.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t *>(&lo2); pk.y = *reinterpret_cast<const uint32_t *>(&hi2);
        *reinterpret_cast<uint2 *>(y + i * 4) = pk;
    }
}
}  // namespace ttdg

extern "C" int ttdg_relu_bn_bwd2(const float *g, const float *y, const float *scale, int C, int64_t numel, float *out_pre, float *out_conv,
                                 void *stream) {
    TTDG_CHECK_ARG(g && y && scale && out_pre && out_conv && C % 4 == 0 && numel % 4 == 0 && numel >= 0);
    if (numel == 0) return 0;
    int64_t nb = (numel / 4 + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    count_launches(1);
    relu_bn_bwd2_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(g, y, scale, C, numel / 4, out_pre, out_conv);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_relu_bn_bwd_bf16y(const float *g, const void *y_bf16, const float *scale, int C, int64_t numel, float *out, void *stream) {
    TTDG_CHECK_ARG(g && y_bf16 && out && C % 4 == 0 && numel % 4 == 0 && numel >= 0);
    if (numel == 0) return 0;
    int64_t nb = (numel / 4 + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    count_launches(1);
    relu_bn_bwd_bf16y_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(g, reinterpret_cast<const __nv_bfloat16 *>(y_bf16), scale, C, numel / 4, out);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_maxpool3x3s2_bf16(const void *x, int N, int H, int W, int C, void *y, void *stream) {
    TTDG_CHECK_ARG(x && y && C % 4 == 0);
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int64_t total = (int64_t)N * Ho * Wo * (C / 4);
    if (total == 0) return 0;
    int64_t nb = (total + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    count_launches(1);
    maxpool3x3s2_bf16_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16 *>(x), N, H, W, C, Ho, Wo,
                                                                            reinterpret_cast<__nv_bfloat16 *>(y));
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_bias_grad(const float *g, int64_t pixels, int C, float *out, void *stream) {
    TTDG_CHECK_ARG(g && out && pixels >= 0 && C > 0);
    if (pixels == 0) return 0;
    count_launches(1);
    int chunks = (int)((pixels + 1023) / 1024);
    if (chunks > 148) chunks = 148;
    bias_grad_kernel<<<dim3(ceil_div(C, 32), chunks), 256, 0, (cudaStream_t)stream>>>(g, pixels, C, out);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_maxpool3x3s2(const float *x, int N, int H, int W, int C, float *y, void *stream) {
    TTDG_CHECK_ARG(x && y && C % 4 == 0);
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int64_t total = (int64_t)N * Ho * Wo * (C / 4);
    if (total == 0) return 0;
    int64_t nb = (total + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    count_launches(1);
    maxpool3x3s2_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, Ho, Wo, y);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_resample2(const float *x, float *y, int N, int Hs, int Ws, int C, int mode, void *stream) {
    TTDG_CHECK_ARG(x && y && C % 4 == 0 && mode >= 0 && mode <= 2);
    const int64_t total = (int64_t)N * Hs * Ws * (C / 4);
    if (total == 0) return 0;
    int64_t nb = (total + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    count_launches(1);
    resample2_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(x, y, N, Hs, Ws, C, mode);
    TTDG_LAUNCH_RET();
}

extern "C" int ttdg_preprocess(const unsigned char *img_u8, int N, int H, int W, int Wp, int left, float mean0, float mean1, float mean2,
                               float *out, void *stream) {
    TTDG_CHECK_ARG(img_u8 && out && N >= 0 && H > 0 && W > 0 && left >= 0 && Wp >= left + W);
    if (N == 0) return 0;
    int64_t nb = ((int64_t)N * H * Wp + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    count_launches(1);
    preprocess_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(img_u8, N, H, W, Wp, left, mean0, mean1, mean2, out);
    TTDG_LAUNCH_RET();
}
