// sinkhorn_small.cuh - in-shared-memory log-space Sinkhorn on one small matrix, fp64.
//
// Restates pygmtools.sinkhorn(backend='pytorch') 0.3.8 (reference call site utils/sinkhorn.py:85-87;
// SURVEY.md Appendix B) for ONE item in its working orientation (rows <= cols):
//     z = s / tau;   dummy rows = (cols - rows) extra rows filled with -100;
//     step k even: z[r,:] -= logsumexp_q z[r,:]   (every row, dummy rows too)
//     step k odd : z[:,q] -= logsumexp_r z[:,q]   (every column, over real + dummy rows)
//     out = exp(z) on the real rows.
// All dummy rows are identical at every step, so they are carried as ONE vector padv[q] with multiplicity
// `mult` (their contribution to a column sum is mult * exp(padv[q])).
//
// The working matrix is addressed through two strides so the same code serves the row-major tile of the
// stand-alone kernels and both orientations inside the GA-GM solver:
//     element(r, q) = z[r * ldr + q * ldq].
// Lines (rows or columns) are distributed over the warps of a thread group; lanes stride over a line.
#pragma once
#include "common.cuh"

namespace ttdg {

// logsumexp of one line held by a warp: elements v(e), e in [0, n), plus an optional extra element
// `extra` with integer multiplicity `mult` (mult == 0 -> ignored).  Every lane returns the result.
template <class F>
__device__ __forceinline__ double line_lse(int n, F v, double extra, int mult, int lane) {
    double m = -INFINITY;
    for (int e = lane; e < n; e += 32) m = fmax(m, v(e));
    m = warp_max(m);
    if (mult > 0) m = fmax(m, extra);
    if (m == -INFINITY) return -INFINITY;
    double s = 0.0;
    for (int e = lane; e < n; e += 32) s += exp(v(e) - m);
    s = warp_sum(s);
    if (mult > 0) s += (double)mult * exp(extra - m);
    return m + log(s);
}

// One normalisation step (step index `it`) in place.  If L != nullptr the subtracted log-sum-exp of every
// line is recorded (L[line], dummy row at L[nr]) for the backward pass.
// gw = warp index inside the group, gnw = warps in the group.  Caller synchronises the group afterwards.
__device__ __forceinline__ void sinkhorn_step(double *z, int ldr, int ldq, int nr, int nq, double *padv, int mult,
                                              int it, double *L, int gw, int gnw, int lane) {
    if ((it & 1) == 0) {
        for (int r = gw; r < nr + (mult > 0 ? 1 : 0); r += gnw) {
            if (r < nr) {
                double *zr = z + (size_t)r * ldr;
                const double l = line_lse(nq, [=](int q) { return zr[q * ldq]; }, 0.0, 0, lane);
                for (int q = lane; q < nq; q += 32) zr[q * ldq] -= l;
                if (L && lane == 0) L[r] = l;
            } else {
                const double l = line_lse(nq, [=](int q) { return padv[q]; }, 0.0, 0, lane);
                __syncwarp();
                for (int q = lane; q < nq; q += 32) padv[q] -= l;
                if (L && lane == 0) L[nr] = l;
            }
        }
    } else {
        for (int q = gw; q < nq; q += gnw) {
            double *zq = z + (size_t)q * ldq;
            const double pe = mult > 0 ? padv[q] : 0.0;
            const double l = line_lse(nr, [=](int r) { return zq[(size_t)r * ldr]; }, pe, mult, lane);
            for (int r = lane; r < nr; r += 32) zq[(size_t)r * ldr] -= l;
            __syncwarp();
            if (mult > 0 && lane == 0) padv[q] = pe - l;
            if (L && lane == 0) L[q] = l;
        }
    }
}

// Backward of step `it`: on entry z, padv hold the values AFTER the step (z_k), g/gp the gradient w.r.t.
// z_k; on exit z, padv hold z_{k-1} and g/gp the gradient w.r.t. z_{k-1}.  L = recorded lse of this step.
//   row step:  g[r,q] -= exp(z_k[r,q]) * sum_q' g[r,q']
//   col step:  g[r,q] -= exp(z_k[r,q]) * (sum_r' g[r',q] + mult * gp[q])      (same for the dummy row)
__device__ __forceinline__ void sinkhorn_step_bwd(double *z, double *g, int ldr, int ldq, int nr, int nq,
                                                  double *padv, double *gp, int mult, int it, const double *L,
                                                  int gw, int gnw, int lane) {
    if ((it & 1) == 0) {
        for (int r = gw; r < nr + (mult > 0 ? 1 : 0); r += gnw) {
            if (r < nr) {
                double *zr = z + (size_t)r * ldr, *gr = g + (size_t)r * ldr;
                double t = 0.0;
                for (int q = lane; q < nq; q += 32) t += gr[q * ldq];
                t = warp_sum(t);
                const double l = L[r];
                for (int q = lane; q < nq; q += 32) {
                    const double zk = zr[q * ldq];
                    gr[q * ldq] -= exp(zk) * t;
                    zr[q * ldq] = zk + l;
                }
            } else {
                double t = 0.0;
                for (int q = lane; q < nq; q += 32) t += gp[q];
                t = warp_sum(t);
                const double l = L[nr];
                __syncwarp();
                for (int q = lane; q < nq; q += 32) {
                    const double pk = padv[q];
                    gp[q] -= exp(pk) * t;
                    padv[q] = pk + l;
                }
            }
        }
    } else {
        for (int q = gw; q < nq; q += gnw) {
            double *zq = z + (size_t)q * ldq, *gq = g + (size_t)q * ldq;
            double t = 0.0;
            for (int r = lane; r < nr; r += 32) t += gq[(size_t)r * ldr];
            t = warp_sum(t);
            if (mult > 0) t += (double)mult * gp[q];
            const double l = L[q];
            for (int r = lane; r < nr; r += 32) {
                const double zk = zq[(size_t)r * ldr];
                gq[(size_t)r * ldr] -= exp(zk) * t;
                zq[(size_t)r * ldr] = zk + l;
            }
            __syncwarp();
            if (mult > 0 && lane == 0) {
                const double pk = padv[q];
                gp[q] -= exp(pk) * t;
                padv[q] = pk + l;
            }
        }
    }
}

}  // namespace ttdg
