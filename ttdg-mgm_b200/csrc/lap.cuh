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

struct LapWork {                     // per-warp shared-memory workspace
    double u[LAP_MAX_DIM];
    double v[LAP_MAX_DIM];
    double shortest[LAP_MAX_DIM];
    int path[LAP_MAX_DIM];
    int col4row[LAP_MAX_DIM];
    int row4col[LAP_MAX_DIM];
    int remaining[LAP_MAX_DIM];
    unsigned char SR[LAP_MAX_DIM];
    unsigned char SC[LAP_MAX_DIM];
};

// Cost: functor double operator()(int i, int j) for the WORKING matrix (nr <= nc).
// On return w.col4row[i] (i < nr) holds the column assigned to row i.  Must be called by a full warp.
template <class Cost>
__device__ void lap_solve_warp(int nr, int nc, Cost cost, LapWork &w) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < nr; k += 32) { w.u[k] = 0.0; w.col4row[k] = -1; }
    for (int k = lane; k < nc; k += 32) { w.v[k] = 0.0; w.row4col[k] = -1; }
    __syncwarp();
    for (int cur = 0; cur < nr; ++cur) {
        for (int k = lane; k < nc; k += 32) {
            w.remaining[k] = nc - k - 1;
            w.shortest[k] = INFINITY;
            w.path[k] = -1;
            w.SC[k] = 0;
        }
        for (int k = lane; k < nr; k += 32) w.SR[k] = 0;
        __syncwarp();
        double minVal = 0.0;
        int num_remaining = nc;
        int sink = -1;
        int i = cur;
        while (sink == -1) {
            if (lane == 0) w.SR[i] = 1;
            const double ui = w.u[i];
            double m = INFINITY;     // lane-local minimum of shortest[]
            int first = 0x7fffffff;  // first position holding m
            int lastfree = -1;       // last unassigned position holding m
            for (int it = lane; it < num_remaining; it += 32) {
                const int j = w.remaining[it];
                const double r = ((minVal + cost(i, j)) - ui) - w.v[j];
                double s = w.shortest[j];
                if (r < s) { w.path[j] = i; w.shortest[j] = r; s = r; }
                const bool fr = (w.row4col[j] == -1);
                if (s < m) { m = s; first = it; lastfree = fr ? it : -1; }
                else if (s == m && fr) { lastfree = it; }
            }
            double gm = m;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) gm = fmin(gm, __shfl_xor_sync(TTDG_FULL, gm, o));
            const bool mine = (m == gm) && (first != 0x7fffffff);
            const int gfirst = warp_min_i(mine ? first : 0x7fffffff);
            const int glast = warp_max_i(mine ? lastfree : -1);
            const int index = (glast != -1) ? glast : gfirst;
            minVal = gm;
            __syncwarp();
            const int j = w.remaining[index];
            const int r4c = w.row4col[j];
            if (r4c == -1) sink = j; else i = r4c;
            --num_remaining;
            __syncwarp();
            if (lane == 0) { w.SC[j] = 1; w.remaining[index] = w.remaining[num_remaining]; }
            __syncwarp();
        }
        // dual update
        for (int r = lane; r < nr; r += 32) {
            if (r == cur) w.u[r] += minVal;
            else if (w.SR[r]) w.u[r] += minVal - w.shortest[w.col4row[r]];
        }
        for (int j = lane; j < nc; j += 32)
            if (w.SC[j]) w.v[j] -= minVal - w.shortest[j];
        __syncwarp();
        // augment along the path (sequential, short)
        if (lane == 0) {
            int j = sink;
            while (true) {
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
