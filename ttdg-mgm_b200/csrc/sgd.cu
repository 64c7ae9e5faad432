// sgd.cu - fused SGD(momentum, weight decay) step over one flat fp32 bucket.
// Reference: the optimizer the caller owns at engine/trainer.py:480-482 (torch.optim.SGD built by Detectron2's
// build_optimizer, train_net.py:65): g += wd * p; m = first ? g : mu * m + g; p -= lr * m  (dampening 0, no
// nesterov).  grad_scale folds the 1/world_size of the gradient all-reduce into the same pass.
// HBM-bound: 5 x 4 bytes per parameter (read p, g, m; write p, m), float4 accesses, grid = 148 x 4 CTAs.
#include "common.cuh"

namespace ttdg {

__global__ void __launch_bounds__(256)
sgd_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, int64_t n, float lr, float mu,
           float wd, float gs, int first) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
        float4 pv = reinterpret_cast<float4 *>(p)[i];
        const float4 gv = reinterpret_cast<const float4 *>(g)[i];
        float4 mv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4 *>(m)[i];
        float d;
        d = fmaf(wd, pv.x, gv.x * gs); mv.x = first ? d : fmaf(mu, mv.x, d); pv.x = fmaf(-lr, mv.x, pv.x);
        d = fmaf(wd, pv.y, gv.y * gs); mv.y = first ? d : fmaf(mu, mv.y, d); pv.y = fmaf(-lr, mv.y, pv.y);
        d = fmaf(wd, pv.z, gv.z * gs); mv.z = first ? d : fmaf(mu, mv.z, d); pv.z = fmaf(-lr, mv.z, pv.z);
        d = fmaf(wd, pv.w, gv.w * gs); mv.w = first ? d : fmaf(mu, mv.w, d); pv.w = fmaf(-lr, mv.w, pv.w);
        reinterpret_cast<float4 *>(p)[i] = pv;
        reinterpret_cast<float4 *>(m)[i] = mv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        const float d = fmaf(wd, p[i], g[i] * gs);
        const float mv = first ? d : fmaf(mu, m[i], d);
        m[i] = mv;
        p[i] = fmaf(-lr, mv, p[i]);
    }
}

}  // namespace ttdg

extern "C" int ttdg_sgd_step(float *p, const float *g, float *m, int64_t n, float lr, float momentum, float weight_decay,
                             float grad_scale, int first_step, void *stream) {
    TTDG_CHECK_ARG(p && g && m && n >= 0);
    TTDG_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m) & 15) == 0);
    if (n == 0) return 0;
    int64_t blocks = ((n >> 2) + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    ttdg::count_launches(1);
    ttdg::sgd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, n, lr, momentum, weight_decay, grad_scale,
                                                                        first_step);
    TTDG_LAUNCH_RET();
}
