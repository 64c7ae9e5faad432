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

// max over the warp of an fp64 value in two 32-bit redux.max (order-preserving integer key) instead of five rounds of 64-bit
// shuffles (2 x 36 cycles each on B200, profiles/r01_microlat_b200.txt); exact
__device__ __forceinline__ double warp_max_redux(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long k = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mhi = __reduce_max_sync(TTDG_FULL, hi);
    const unsigned mlo = __reduce_max_sync(TTDG_FULL, hi == mhi ? lo : 0u);
    const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
    return __longlong_as_double((long long)((mk >> 63) ? (mk & 0x7FFFFFFFFFFFFFFFull) : ~mk));
}

// logsumexp of B lines at once (v(b, e) = element e of line b, all lines n long; `extra[b]` with multiplicity mult as in
// line_lse).  The B reductions are independent dependency chains, so their shuffle / exp / log latencies overlap - a warp that
// owns three lines of a step used to walk them one after the other (3 x ~1.3 k cycles per step, 20 steps per projection).  Same
// per-lane summation order and the same xor tree as line_lse: identical results.
template <int B, class F>
__device__ __forceinline__ void lines_lse(int n, F v, const double (&extra)[B], int mult, int lane, double (&l)[B]) {
    double m[B], s[B];
#pragma unroll
    for (int b = 0; b < B; ++b) m[b] = -INFINITY;
    for (int e = lane; e < n; e += 32) {
#pragma unroll
        for (int b = 0; b < B; ++b) m[b] = fmax(m[b], v(b, e));
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
        m[b] = warp_max_redux(m[b]);
        if (mult > 0) m[b] = fmax(m[b], extra[b]);
        s[b] = 0.0;
    }
    for (int e = lane; e < n; e += 32) {
#pragma unroll
        for (int b = 0; b < B; ++b) s[b] += (m[b] == -INFINITY) ? 0.0 : exp(v(b, e) - m[b]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int b = 0; b < B; ++b) s[b] += __shfl_xor_sync(TTDG_FULL, s[b], o);
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
        if (m[b] == -INFINITY) { l[b] = -INFINITY; continue; }
        if (mult > 0) s[b] += (double)mult * exp(extra[b] - m[b]);
        l[b] = m[b] + log(s[b]);
    }
}

// One normalisation step (step index `it`) in place.  If L != nullptr the subtracted log-sum-exp of every
// line is recorded (L[line], dummy row at L[nr]) for the backward pass.
// gw = warp index inside the group, gnw = warps in the group.  Caller synchronises the group afterwards.
// A warp's lines (gw, gw + gnw, ...) are processed three at a time (lines_lse).
__device__ __forceinline__ void sinkhorn_step(double *z, int ldr, int ldq, int nr, int nq, double *padv, int mult,
                                              int it, double *L, int gw, int gnw, int lane) {
    constexpr int B = 3;
    if ((it & 1) == 0) {
        const int nlines = nr + (mult > 0 ? 1 : 0);                   // the dummy rows are one more line (padv)
        for (int r0 = gw; r0 < nlines; r0 += B * gnw) {
            double *base[B];
            int step[B];
            bool on[B];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int r = r0 + b * gnw;
                on[b] = r < nlines;
                const int rr = on[b] ? r : r0;                        // lines past the end repeat the first one (results unused)
                base[b] = rr < nr ? z + (size_t)rr * ldr : padv;
                step[b] = rr < nr ? ldq : 1;
            }
            const double none[B] = {0.0, 0.0, 0.0};
            double l[B];
            lines_lse<B>(nq, [&](int b, int q) { return base[b][q * step[b]]; }, none, 0, lane, l);
            __syncwarp();
#pragma unroll
            for (int b = 0; b < B; ++b) {
                if (!on[b]) continue;
                const int r = r0 + b * gnw;
                for (int q = lane; q < nq; q += 32) base[b][q * step[b]] -= l[b];
                if (L && lane == 0) L[r < nr ? r : nr] = l[b];
            }
        }
    } else {
        for (int q0 = gw; q0 < nq; q0 += B * gnw) {
            double *base[B];
            double pe[B];
            bool on[B];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int q = q0 + b * gnw;
                on[b] = q < nq;
                const int qq = on[b] ? q : q0;
                base[b] = z + (size_t)qq * ldq;
                pe[b] = mult > 0 ? padv[qq] : 0.0;
            }
            double l[B];
            lines_lse<B>(nr, [&](int b, int r) { return base[b][(size_t)r * ldr]; }, pe, mult, lane, l);
#pragma unroll
            for (int b = 0; b < B; ++b) {
                if (!on[b]) continue;
                const int q = q0 + b * gnw;
                for (int r = lane; r < nr; r += 32) base[b][(size_t)r * ldr] -= l[b];
                if (lane == 0) {
                    if (mult > 0) padv[q] = pe[b] - l[b];
                    if (L) L[q] = l[b];
                }
            }
            __syncwarp();
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
