// lap.cuh - warp-cooperative rectangular linear sum assignment, fp64.
//
// Restates scipy.optimize.linear_sum_assignment (rectangular_lsap.cpp: Crouse's shortest augmenting
// path) including its tie-breaking, as used by the reference's hungarian()
// (adapteacher/modeling/GModule/utils/hungarian.py:34,58-65):
//   * the column scan runs over the `remaining` list, filled in reverse, compacted by swap-with-last;
//   * a column replaces the current best when strictly lower, or equal AND unassigned, so the winner is
//     the LAST unassigned position among the minima if one exists, else the FIRST minimum;
//   * r = ((minVal + cost) - u[i]) - v[j] in exactly that order (no fused ops are possible: no products).
// The scan is distributed over the 32 lanes (lane handles positions it = lane, lane+32, ...) and the
// (min, first, last-unassigned) triple is reduced with shuffles, which reproduces the sequential result.
#pragma once
#include "common.cuh"

namespace ttdg {

constexpr int LAP_MAX_DIM = 128;     // rows, cols <= 128 (graphs have <= ~95 nodes, universe is 32)
constexpr int LAP_SLOTS = LAP_MAX_DIM / 32;

struct LapWork {                     // per-warp shared-memory workspace (state that crosses lanes / augmentations)
    double u[LAP_MAX_DIM];
    double v[LAP_MAX_DIM];
    double shortest[LAP_MAX_DIM];    // final shortest[] of the columns scanned in the current augmentation
    int path[LAP_MAX_DIM];
    int col4row[LAP_MAX_DIM];
    int row4col[LAP_MAX_DIM];
    int visited[LAP_MAX_DIM];        // rows put into SR in the current augmentation, in order
    int stat_steps, stat_hops;       // running totals (lane 0): Dijkstra steps and augmenting-path hops, for profiling
};

// order-preserving map double -> uint64 (no NaNs here)
__device__ __forceinline__ unsigned long long lap_ord(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double lap_unord(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// Cost: functor double operator()(int i, int j) for the WORKING matrix (nr <= nc).
// On return w.col4row[i] (i < nr) holds the column assigned to row i.  Must be called by a full warp.
//
// Per-column state (shortest, v, path, position in SciPy's `remaining` list, row4col) lives in REGISTERS of the lane
// that owns the column (j = lane + 32 t, t < SLOTS = ceil(nc / 32)); one inner iteration is a few FP64 ops per owned
// column plus three redux.sync reductions (value high word, value low word, winner key) instead of shuffle butterflies
// over shared memory.
// SciPy's selection rule "strictly lower, or equal and unassigned" over the scan order it = 0 .. num_remaining-1 picks,
// among the minima, the LAST unassigned position if any, else the FIRST position; with the tie key
//     unassigned: 256 + it,   assigned: 255 - it      (maximised)
// the winner of (value ascending, key descending) is exactly that element.  The tie key is unique among the remaining
// columns, so the winner's column index and its row4col ride along in the low bits of the same 32-bit key
// (tie << 16 | column << 8 | row4col + 1) and one redux.max delivers all three.  `remaining` is filled in reverse
// (position it holds column nc-1-it) and compacted by moving the last element into the freed position - tracked here
// as a per-column position register.
template <int SLOTS, class Cost>
__device__ void lap_solve_warp_t(int nr, int nc, Cost cost, LapWork &w) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < nr; k += 32) { w.u[k] = 0.0; w.col4row[k] = -1; }
    for (int k = lane; k < nc; k += 32) { w.v[k] = 0.0; w.row4col[k] = -1; }
    __syncwarp();
    for (int cur = 0; cur < nr; ++cur) {
        double sh[SLOTS], vj[SLOTS];
        unsigned long long ks[SLOTS];                           // lap_ord(sh[t])
        int pos[SLOTS], r4c[SLOTS], pth[SLOTS];
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            const int j = lane + 32 * t;
            sh[t] = INFINITY; ks[t] = lap_ord(INFINITY); pth[t] = -1;
            if (j < nc) { vj[t] = w.v[j]; pos[t] = nc - 1 - j; r4c[t] = w.row4col[j]; }
            else { vj[t] = 0.0; pos[t] = -1; r4c[t] = 0; }
        }
        double minVal = 0.0;
        int num_remaining = nc, sink = -1, i = cur, nvis = 0;
        while (sink == -1) {
            if (lane == 0) w.visited[nvis] = i;
            ++nvis;
            const double ui = w.u[i];
            unsigned long long bk = 0xFFFFFFFFFFFFFFFFull;      // lane-local best (ordered value)
            unsigned bkey = 0;                                  // its winner key
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) {
                if (pos[t] >= 0) {
                    const double r = ((minVal + cost(i, lane + 32 * t)) - ui) - vj[t];
                    if (r < sh[t]) { sh[t] = r; pth[t] = i; ks[t] = lap_ord(r); }
                    const unsigned tie = (r4c[t] == -1) ? 256u + (unsigned)pos[t] : 255u - (unsigned)pos[t];
                    const unsigned key = (tie << 16) | ((unsigned)(lane + 32 * t) << 8) | (unsigned)(r4c[t] + 1);
                    if (ks[t] < bk || (ks[t] == bk && key > bkey)) { bk = ks[t]; bkey = key; }
                }
            }
            const unsigned hi = (unsigned)(bk >> 32), lo = (unsigned)bk;
            const unsigned mhi = __reduce_min_sync(TTDG_FULL, hi);
            const unsigned mlo = __reduce_min_sync(TTDG_FULL, hi == mhi ? lo : 0xFFFFFFFFu);
            const bool cand = (hi == mhi) && (lo == mlo) && (bkey != 0u);
            const unsigned mkey = __reduce_max_sync(TTDG_FULL, cand ? bkey : 0u);
            minVal = lap_unord(((unsigned long long)mhi << 32) | mlo);
            const unsigned mtie = mkey >> 16;
            const int psel = mtie >= 256u ? (int)(mtie - 256u) : (int)(255u - mtie);
            const int jsel = (int)((mkey >> 8) & 0xFFu), rsel = (int)(mkey & 0xFFu) - 1;
            if (rsel == -1) sink = jsel; else i = rsel;
            const int last = num_remaining - 1;
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) {
                const int j = lane + 32 * t;
                if (j == jsel) {                                  // scanned: SC[j] = true, leaves `remaining`
                    w.shortest[j] = sh[t]; w.path[j] = pth[t]; pos[t] = -2;
                } else if (pos[t] == last) pos[t] = psel;         // remaining[index] = remaining[--num_remaining]
            }
            num_remaining = last;
        }
        __syncwarp();
        // dual update
        for (int k = lane; k < nvis; k += 32) {
            const int r = w.visited[k];
            if (r == cur) w.u[r] += minVal;
            else w.u[r] += minVal - w.shortest[w.col4row[r]];
        }
#pragma unroll
        for (int t = 0; t < SLOTS; ++t)
            if (pos[t] == -2) w.v[lane + 32 * t] = vj[t] - (minVal - sh[t]);
        __syncwarp();
        // augment along the path (sequential, short)
        if (lane == 0) {
            w.stat_steps += nvis;
            int j = sink;
            while (true) {
                ++w.stat_hops;
                const int r = w.path[j];
                w.row4col[j] = r;
                const int t = w.col4row[r];
                w.col4row[r] = j;
                j = t;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }
}

template <class Cost>
__device__ __forceinline__ void lap_solve_warp(int nr, int nc, Cost cost, LapWork &w) {
    if (nc <= 32) lap_solve_warp_t<1>(nr, nc, cost, w);
    else if (nc <= 64) lap_solve_warp_t<2>(nr, nc, cost, w);
    else lap_solve_warp_t<LAP_SLOTS>(nr, nc, cost, w);
}

// hungarian(s) for one stored n1 x n2 fp32 score matrix (leading dimension ld): perm = 0/1 matrix of the
// max-weight assignment.  cost = (double)(-s); the working matrix is the transpose when n2 < n1.
__device__ inline void hungarian_warp(const float *s, float *perm, int n1, int n2, int ld_s, int ld_p, LapWork &w) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < n1 * n2; k += 32) perm[(k / n2) * ld_p + (k % n2)] = 0.0f;
    if (n1 == 0 || n2 == 0) return;
    __syncwarp();
    if (n2 >= n1) {
        lap_solve_warp(n1, n2, [=](int i, int j) { return (double)(s[i * ld_s + j] * -1.0f); }, w);
        for (int i = lane; i < n1; i += 32) perm[i * ld_p + w.col4row[i]] = 1.0f;
    } else {
        lap_solve_warp(n2, n1, [=](int i, int j) { return (double)(s[j * ld_s + i] * -1.0f); }, w);
        for (int i = lane; i < n2; i += 32) perm[w.col4row[i] * ld_p + i] = 1.0f;
    }
    __syncwarp();
}

}  // namespace ttdg
