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
    int stat_fast_ok, stat_fast_fallback;      // certified fast solves / fall-backs to the SciPy-order solve
    int fast_mode;                   // 1: row-reduction start, 2: Jacobi auction start (set by the caller)
    int stat_free[8];                // lean solve: free rows at the start of each auction round and after the last (running totals)
    long long stat_ck[6];            // lean solve (lane 0): cycles in the auction scans / bid resolution / augmentations / certificate
                                     // (all / reachability / Kahn)
    unsigned long long bidkey[LAP_MAX_DIM];    // auction: best bid per column of the current round
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

__device__ __forceinline__ uint32_t lap_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-space fp64 load from a 32-bit shared address (no generic->shared window arithmetic on the critical path)
__device__ __forceinline__ double lap_lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
// cost(i, j) = -Z[i * si + j * sj] for a double matrix in shared memory (the GA-GM projector's V tile)
struct LapSmemNegCost {
    uint32_t base; int si, sj;
    __device__ __forceinline__ double operator()(int i, int j) const { return -lap_lds_f64(base + (uint32_t)((i * si + j * sj) << 3)); }
};

// Cost: functor double operator()(int i, int j) for the WORKING matrix (nr <= nc).
// On return w.col4row[i] (i < nr) holds the column assigned to row i.  Must be called by a full warp.
//
// Per-column state (shortest, v, path, position in SciPy's `remaining` list, row4col) lives in REGISTERS of the lane
// that owns the column (j = lane + 32 t, t < SLOTS = ceil(nc / 32)); v stays there across augmentations.  One Dijkstra
// step is a branch-free relaxation of the owned columns (so the slots of a lane overlap in the FP64 pipe) plus ONE
// redux.sync on the high word of the ordered value: if a single lane holds that minimum (the usual case) its low word and
// winner key are fetched with two independent shuffles; only on a high-word tie are the low word and the key reduced too.
// SciPy's selection rule "strictly lower, or equal and unassigned" over the scan order it = 0 .. num_remaining-1 picks,
// among the minima, the LAST unassigned position if any, else the FIRST position; with the tie key
//     unassigned: 256 + it,   assigned: 255 - it      (maximised)
// the winner of (value ascending, key descending) is exactly that element.  The tie key is unique among the remaining
// columns, so the winner's column index and its row4col ride along in the low bits of the same 32-bit key
// (tie << 16 | column << 8 | row4col + 1).  `remaining` is filled in reverse (position it holds column nc-1-it) and
// compacted by moving the last element into the freed position - tracked here as a per-column position register.
// Dual update: the rows visited in an augmentation are `cur` and row4col[j] of every scanned, assigned column j, so
// u[row4col[j]] += minVal - shortest[j] is applied by the lane that owns column j (SciPy: u[i] += minVal -
// shortest[col4row[i]], the same operands).
//
// FAST = true (nr <= 32, opt-in): SciPy's order of augmentations and its tie rule only matter when the optimum is not
// unique, so the solve is (1) started from a row reduction - lane = row: u[i] = min_j c[i][j], the row takes its arg-min
// column if no lower row wants it (match.any) - which is dual feasible with v = 0 and leaves only the losers of a column
// conflict to the Dijkstra augmentations below, and (2) CERTIFIED: with c̄ = (c - u) - v the optimum is unique by the
// margin delta iff no chain of delta-tight edges leads from a row whose column has a zero price to a free column, and the
// digraph {i -> i' : c̄[i][col4row[i']] <= delta} is acyclic (Kahn's elimination on 32-bit adjacency masks).  Returns false when the certificate fails (structural ties: universe
// slots no graph uses) - the caller then runs the SciPy-order solve.
template <int SLOTS, bool FAST, class Cost>
__device__ bool lap_solve_warp_t(int nr, int nc, Cost cost, LapWork &w) {
    const int lane = threadIdx.x & 31;
    const uint32_t u_sa = lap_smem_u32(w.u);
    for (int k = lane; k < nr; k += 32) { w.u[k] = 0.0; w.col4row[k] = -1; }
    for (int k = lane; k < nc; k += 32) { w.row4col[k] = -1; }
    double vj[SLOTS];
    int jc[SLOTS];
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) { vj[t] = 0.0; jc[t] = min(lane + 32 * t, nc - 1); }
    int steps = 0, hops = 0;
    __syncwarp();
    if (FAST && w.fast_mode == 2) {
        // Jacobi auction with epsilon = 0 (the augmenting row reduction of Jonker-Volgenant, all free rows bidding at once):
        // lane = row.  A free row finds its best and second-best reduced cost c[i][j] - v[j]; it bids for the best column
        // with the price cut (second - best); per column the largest cut wins, the previous owner is set free.  Prices
        // only fall, so the duals stay feasible (u_i = second-best value, tight on the won column) and free columns keep
        // v = 0.  Ties (cut 0) can ping-pong, hence the round cap: whoever is still free goes to the Dijkstra loop below.
        for (int k = lane; k < nc; k += 32) { w.v[k] = 0.0; w.bidkey[k] = 0ull; }
        __syncwarp();
        int myc = -1, prev_free = 33, stalls = 0;
        for (int round = 0; round < 16; ++round) {
            const bool isfree = lane < nr && myc == -1;
            const int nfree = __popc(__ballot_sync(TTDG_FULL, isfree));
            if (nfree == 0) break;
            if (nfree >= prev_free && ++stalls >= 2) break;     // ping-pong on ties: leave the rest to the Dijkstra loop
            prev_free = nfree;
            double m1 = INFINITY, m2 = INFINITY;
            int j1 = 0;
            unsigned long long mykey = 0ull;
            if (isfree) {
                for (int j = 0; j < nc; ++j) {
                    const double val = cost(lane, j) - w.v[j];
                    if (val < m1) { m2 = m1; m1 = val; j1 = j; } else if (val < m2) m2 = val;
                }
                const double cut = (m2 < INFINITY) ? m2 - m1 : 0.0;
                mykey = ((lap_ord(cut) & ~31ull) | (unsigned long long)(31 - lane)) | (1ull << 63);
                atomicMax(&w.bidkey[j1], mykey);
            }
            __syncwarp();
            if (isfree) {
                if (w.bidkey[j1] == mykey) {                      // winner of column j1
                    const int old = w.row4col[j1];
                    if (old >= 0) w.col4row[old] = -1;
                    w.row4col[j1] = lane; w.col4row[lane] = j1;
                    w.u[lane] = (m2 < INFINITY) ? m2 : m1;
                    if (m2 < INFINITY) w.v[j1] -= (m2 - m1);
                } else {
                    w.u[lane] = m1;                               // still free: a feasible lower bound for the Dijkstra start
                }
            }
            __syncwarp();
            if (isfree) w.bidkey[j1] = 0ull;
            __syncwarp();
            myc = lane < nr ? w.col4row[lane] : 0;
        }
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) if (lane + 32 * t < nc) vj[t] = w.v[lane + 32 * t];
        __syncwarp();
    } else if (FAST) {
        double umin = INFINITY;
        int arg = 0;
        if (lane < nr)
            for (int j = 0; j < nc; ++j) { const double cij = cost(lane, j); if (cij < umin) { umin = cij; arg = j; } }
        const unsigned same = __match_any_sync(TTDG_FULL, lane < nr ? arg : -1 - lane);
        const bool win = lane < nr && (__ffs(same) - 1 == lane);
        if (lane < nr) { w.u[lane] = umin; if (win) { w.col4row[lane] = arg; w.row4col[arg] = lane; } }
        __syncwarp();
    }
    for (int cur = 0; cur < nr; ++cur) {
        if (FAST && w.col4row[cur] != -1) continue;             // assigned by the row reduction (warp-uniform)
        double sh[SLOTS];
        int pos[SLOTS], r4c[SLOTS], pth[SLOTS];
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            const int j = lane + 32 * t;
            sh[t] = INFINITY; pth[t] = -1;
            if (j < nc) { pos[t] = nc - 1 - j; r4c[t] = w.row4col[j]; }
            else { pos[t] = -1; r4c[t] = 0; }
        }
        double minVal = 0.0;
        int num_remaining = nc, sink = -1, i = cur;
        while (sink == -1) {
            ++steps;
            const double ui = lap_lds_f64(u_sa + (uint32_t)(i << 3));
            double c[SLOTS];
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) c[t] = cost(i, jc[t]);
            double best = INFINITY;                             // lane-local best value
            unsigned bkey = 0u;                                 // its winner key (0: no live column in this lane)
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) {
                const bool live = pos[t] >= 0;
                const double r = ((minVal + c[t]) - ui) - vj[t];
                const bool upd = live && (r < sh[t]);
                sh[t] = upd ? r : sh[t];
                pth[t] = upd ? i : pth[t];
                const unsigned tie = (r4c[t] == -1) ? 256u + (unsigned)pos[t] : 255u - (unsigned)pos[t];
                const unsigned key = live ? ((tie << 16) | ((unsigned)(lane + 32 * t) << 8) | (unsigned)(r4c[t] + 1)) : 0u;
                const double cv = live ? sh[t] : INFINITY;
                const bool better = live && ((cv < best) || (cv == best && key > bkey));
                best = better ? cv : best;
                bkey = better ? key : bkey;
            }
            const unsigned long long bk = lap_ord(best);
            const unsigned hi = (unsigned)(bk >> 32), lo = (unsigned)bk;
            const unsigned mhi = __reduce_min_sync(TTDG_FULL, bkey != 0u ? hi : 0xFFFFFFFFu);
            const bool at_min = (hi == mhi) && (bkey != 0u);
            const unsigned tied = __ballot_sync(TTDG_FULL, at_min);
            unsigned mlo, mkey;
            if (__popc(tied) == 1) {
                const int src = __ffs(tied) - 1;
                mlo = __shfl_sync(TTDG_FULL, lo, src);
                mkey = __shfl_sync(TTDG_FULL, bkey, src);
            } else {
                mlo = __reduce_min_sync(TTDG_FULL, at_min ? lo : 0xFFFFFFFFu);
                mkey = __reduce_max_sync(TTDG_FULL, (at_min && lo == mlo) ? bkey : 0u);
            }
            minVal = lap_unord(((unsigned long long)mhi << 32) | mlo);
            const unsigned mtie = mkey >> 16;
            const int psel = mtie >= 256u ? (int)(mtie - 256u) : (int)(255u - mtie);
            const int jsel = (int)((mkey >> 8) & 0xFFu), rsel = (int)(mkey & 0xFFu) - 1;
            if (rsel == -1) sink = jsel; else i = rsel;
            const int last = num_remaining - 1;
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) {
                const int j = lane + 32 * t;
                if (j == jsel) pos[t] = -2;                       // scanned: SC[j] = true, leaves `remaining`
                else if (pos[t] == last) pos[t] = psel;           // remaining[index] = remaining[--num_remaining]
            }
            num_remaining = last;
        }
        // dual update (u in shared memory by the owner of the scanned column, v in registers) + path of the scanned columns
        if (lane == 0) w.u[cur] += minVal;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t)
            if (pos[t] == -2) {
                const double d = minVal - sh[t];
                if (r4c[t] >= 0) w.u[r4c[t]] += d;
                vj[t] -= d;
                w.path[lane + 32 * t] = pth[t];
            }
        __syncwarp();
        // augment along the path (sequential, short)
        if (lane == 0) {
            int j = sink;
            while (true) {
                ++hops;
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
    if (lane == 0) { w.stat_steps += steps; w.stat_hops += hops; }
    __syncwarp();
    if (!FAST) return true;
    // ---- uniqueness certificate
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) if (lane + 32 * t < nc) w.v[lane + 32 * t] = vj[t];
    __syncwarp();
    double scale = 0.0;
    unsigned adj = 0u;
    bool free_hit = false;
    if (lane < nr) scale = fabs(w.u[lane]);
    for (int o = 16; o > 0; o >>= 1) scale = fmax(scale, __shfl_xor_sync(TTDG_FULL, scale, o));
    const double delta = 1e-9 * (1.0 + scale);
    if (lane < nr) {
        const double ui = w.u[lane];
        const int mine = w.col4row[lane];
        for (int j = 0; j < nc; ++j) {
            if (j == mine) continue;
            const double cb = (cost(lane, j) - ui) - w.v[j];
            if (cb <= delta) { const int r = w.row4col[j]; if (r < 0) free_hit = true; else adj |= 1u << r; }
        }
    }
    // A switch to a free column costs  sum of c̄ on the new edges + |v| of the ABANDONED column  (free columns have v = 0,
    // prices are <= 0): it ties only if the chain of tight edges that ends in the free column STARTS at a row whose own column
    // has a zero price.  reach = rows from which a free column is reachable over tight edges.
    unsigned reach = __ballot_sync(TTDG_FULL, free_hit);
    while (true) {
        const unsigned nxt = __ballot_sync(TTDG_FULL, lane < nr && (free_hit || (adj & reach) != 0u));
        if (nxt == reach) break;
        reach = nxt;
    }
    bool zero_start = false;
    if (lane < nr && ((reach >> lane) & 1u)) zero_start = fabs(w.v[w.col4row[lane]]) <= delta;
    bool ok = !__any_sync(TTDG_FULL, zero_start);
    if (ok) {
        unsigned alive = nr >= 32 ? 0xFFFFFFFFu : ((1u << nr) - 1u);
        while (alive) {                                         // Kahn: drop the rows no alive row points to
            const unsigned pointed = __reduce_or_sync(TTDG_FULL, ((alive >> lane) & 1u) ? (adj & alive) : 0u);
            const unsigned next = alive & pointed;
            if (next == alive) break;
            alive = next;
        }
        ok = alive == 0u;
    }
    if (lane == 0) { if (ok) ++w.stat_fast_ok; else ++w.stat_fast_fallback; }
    __syncwarp();
    return ok;
}

// ------------------------------------------------------------------------------------------------------------------------
// Lean certified solve (fast mode 3, nr <= 32): the GA-GM solver's default Hungarian projection.
//
// The optimum of the projector's LAP is almost always unique by a wide margin (generic fp64 costs); then ANY exact method
// returns SciPy's assignment and SciPy's scan order / tie rule are irrelevant.  So: (A) a few Jacobi auction rounds with
// epsilon = 0 (lane = row: best and second-best reduced cost, bid = price cut, largest cut wins the column, feasible duals
// throughout) settle the uncontested rows; (B) the rows still free are augmented by a Dijkstra whose step carries none of
// SciPy's bookkeeping (no `remaining` list positions, no tie keys): lane = column, per-column state in registers, one
// 32-bit redux.min on the high word of the non-negative fp64 distance + one ballot + two shuffles per step (~1/3 of the
// instructions of the SciPy-order step); (C) lap_certificate() proves from the duals that the assignment is optimal AND the
// only optimum within a margin delta - otherwise the caller runs the SciPy-order solve.  Measured on the bench workload's
// cost matrices (oracle trajectory, 1600 LAPs): 3 rounds leave 13 of 32 rows free and 163 instead of 246 Dijkstra steps on
// the slowest graph of an iteration; a full auction instead of (B) does NOT work here (rows want the same few columns:
// median 220 rounds of price war).
struct LapSmemNegCostP {             // cost(i, j) = -z[i * si + j * sj], plain shared-memory loads (the tile is constant during a solve)
    const double *z; int si, sj;
    __device__ __forceinline__ double operator()(int i, int j) const { return -z[i * si + j * sj]; }
};

// (first, second) smallest of a stream with the index of the first: v joins.  Strict '<' keeps the earliest index on ties.
__device__ __forceinline__ void lap_min2(double v, int j, double &m1, double &m2, int &j1) {
    const bool l1 = v < m1;
    const double mx = l1 ? m1 : v;          // max(m1, v)
    m1 = l1 ? v : m1;
    j1 = l1 ? j : j1;
    m2 = mx < m2 ? mx : m2;
}

constexpr int LAP_ARR_ROUNDS = 4;

// Proves optimality and uniqueness of w.col4row from the duals w.u / w.v (fp64).  c̄ = (c - u) - v.
//   feasibility:  c̄ >= -eps everywhere, |c̄| <= eps on the assignment, v <= eps, v = 0 on free columns  => optimal (weak duality);
//   uniqueness:   any other assignment costs  sum c̄(new edges) + sum |v|(abandoned columns)  more, so the optimum is unique by
//   the margin delta iff (a) the digraph {i -> i' : c̄[i][col4row[i']] <= delta} on the rows is acyclic (Kahn's elimination on
//   32-bit adjacency masks) and (b) no chain of delta-tight edges leads from a row whose own column has |v| <= delta to a free
//   column (backward reachability by ballots).  lane = row, nr <= 32.
template <class Cost>
__device__ bool lap_certificate(int nr, int nc, Cost cost, LapWork &w) {
    const int lane = threadIdx.x & 31;
    double scale = 0.0;
    unsigned adj = 0u;
    bool free_hit = false, bad = false;
    if (lane < nr) scale = fabs(w.u[lane]);
    // max |u| to ~2^-20 relative (it only sets the margins): one redux on the high words instead of ten shuffles
    scale = __hiloint2double((int)__reduce_max_sync(TTDG_FULL, (unsigned)__double2hiint(scale)), 0);
    const double delta = 1e-9 * (1.0 + scale), eps = 1e-11 * (1.0 + scale);
    for (int j = lane; j < nc; j += 32) {                        // column properties, lane = column
        const double vj = w.v[j];
        bad = bad || vj > eps || (w.row4col[j] < 0 && vj != 0.0);
    }
    if (lane < nr) {
        const double ui = w.u[lane];
        const int mine = w.col4row[lane];
        bad = bad || mine < 0 || mine >= nc || w.row4col[mine] != lane;
        // branch-free (the lanes would diverge on every column) and unrolled (independent columns: the loads overlap); what depends
        // on the column alone (v <= eps, v == 0 on free columns) was checked by the column's lane above
        if (mine >= 0 && mine < nc) bad = bad || fabs((cost(lane, mine) - ui) - w.v[mine]) > eps;
#pragma unroll 4
        for (int j = 0; j < nc; ++j) {
            const double cb = (cost(lane, j) - ui) - w.v[j];
            const int r = w.row4col[j];
            const bool other = j != mine;
            bad = bad || (other && cb < -eps);
            const bool tight = other && cb <= delta;
            free_hit = free_hit || (tight && r < 0);
            adj |= (tight && r >= 0) ? (1u << (r & 31)) : 0u;
        }
    }
    if (__any_sync(TTDG_FULL, bad)) return false;
    const long long ckc0 = clock64();
    unsigned reach = __ballot_sync(TTDG_FULL, free_hit);
    while (true) {
        const unsigned nxt = __ballot_sync(TTDG_FULL, lane < nr && (free_hit || (adj & reach) != 0u));
        if (nxt == reach) break;
        reach = nxt;
    }
    bool zero_start = false;
    if (lane < nr && ((reach >> lane) & 1u)) zero_start = fabs(w.v[w.col4row[lane]]) <= delta;
    if (__any_sync(TTDG_FULL, zero_start)) return false;
    const long long ckc1 = clock64();
    unsigned alive = nr >= 32 ? 0xFFFFFFFFu : ((1u << nr) - 1u);
    while (alive) {                                             // Kahn: drop the rows no alive row points to
        const unsigned pointed = __reduce_or_sync(TTDG_FULL, ((alive >> lane) & 1u) ? (adj & alive) : 0u);
        const unsigned next = alive & pointed;
        if (next == alive) break;
        alive = next;
    }
    if (lane == 0) { w.stat_ck[4] += ckc1 - ckc0; w.stat_ck[5] += clock64() - ckc1; }
    return alive == 0u;
}

template <int SLOTS, bool BF, class Cost>
__device__ bool lap_lean_warp_t(int nr, int nc, Cost cost, LapWork &w) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < nr; k += 32) { w.u[k] = 0.0; w.col4row[k] = -1; }
    unsigned *bid32 = reinterpret_cast<unsigned *>(w.bidkey);        // auction: best bid per column of the current round
    for (int k = lane; k < nc; k += 32) { w.row4col[k] = -1; w.v[k] = 0.0; bid32[k] = 0u; }
    __syncwarp();
    const long long ck0 = clock64();
    long long ck_res = 0;
    // ---- (A) Jacobi auction rounds, lane = row
    int steps = 0, hops = 0;
    {
        int myc = -1;
        for (int round = 0; round < LAP_ARR_ROUNDS; ++round) {
            const bool isfree = lane < nr && myc == -1;
            const unsigned freemask = __ballot_sync(TTDG_FULL, isfree);
            if (lane == 0) w.stat_free[round] += __popc(freemask);
            if (freemask == 0u) break;
            double m1 = INFINITY, m2 = INFINITY;
            int j1 = 0;
            unsigned mykey = 0u;
            if (isfree) {
                // two independent half scans (even / odd columns) halve the compare-select dependency chain
                double a1 = INFINITY, a2 = INFINITY, b1 = INFINITY, b2 = INFINITY;
                int ja = 0, jb = 0;
                int j = 0;
                for (; j + 1 < nc; j += 2) {
                    const double va = cost(lane, j) - w.v[j], vb = cost(lane, j + 1) - w.v[j + 1];
                    // branch-free with plain compare + select (fmin / fmax carry NaN handling: 30 instructions per column instead of 13)
                    lap_min2(va, j, a1, a2, ja);
                    lap_min2(vb, j + 1, b1, b2, jb);
                }
                if (j < nc) lap_min2(cost(lane, j) - w.v[j], j, a1, a2, ja);
                if (b1 < a1) { m1 = b1; j1 = jb; m2 = fmin(a1, b2); } else { m1 = a1; j1 = ja; m2 = fmin(b1, a2); }
                const double cut = (m2 < INFINITY) ? m2 - m1 : 0.0;
                mykey = (((unsigned)(lap_ord(cut) >> 32) & ~31u) | (unsigned)(31 - lane)) | 0x80000000u;
            }
            // the winner of a column = the largest key among its bidders: ONE native 32-bit shared-memory atomicMax per bidder (a 64-bit
            // key needs a CAS loop; match.any + redux per group was measured slower still).  Any bidder may win - the price falls by the
            // winner's own cut, the duals stay feasible - so the 27 leading bits of the cut + the lane are key enough; deterministic.
            const long long cka = clock64();
            if (isfree) atomicMax(&bid32[j1], mykey);
            __syncwarp();
            if (isfree) {
                if (bid32[j1] == mykey) {                         // winner of column j1
                    const int old = w.row4col[j1];
                    if (old >= 0) w.col4row[old] = -1;
                    w.row4col[j1] = lane; w.col4row[lane] = j1;
                    w.u[lane] = (m2 < INFINITY) ? m2 : m1;
                    if (m2 < INFINITY) w.v[j1] -= (m2 - m1);
                } else {
                    w.u[lane] = m1;                               // still free: a feasible lower bound (prices only fall)
                }
            }
            __syncwarp();
            if (isfree) bid32[j1] = 0u;
            __syncwarp();
            myc = lane < nr ? w.col4row[lane] : 0;
            ck_res += clock64() - cka;
        }
    }
    const long long ck1 = clock64();
    { const unsigned fm = __ballot_sync(TTDG_FULL, lane < nr && w.col4row[lane < nr ? lane : 0] == -1); if (lane == 0) w.stat_free[LAP_ARR_ROUNDS] += __popc(fm); }
    if (BF) {
    // ---- (B) augmentations for the rows still free, lane = column: LABEL-CORRECTING shortest paths instead of Dijkstra.
    // Dijkstra settles one column per step and every step is a ~360-cycle chain of dependent warp-wide operations (load the row,
    // relax, 64-bit arg-min over the lanes, fetch the next row); near-square problems with near-tied costs - the bench workload
    // after a few adaptation steps - need 300-500 such steps per solve.  Here a ROUND relaxes all rows whose distance improved
    // (independent, pipelined: ~15 cycles each) and only then reduces once: the free column's best distance bounds the search
    // (rows at or beyond it are never expanded), distances below it come out exact, and the number of rounds is the hop depth
    // of the shortest-path tree, not the number of columns.  Duals: v_j -= max(0, D - d_j), u_{row(j)} += the same, u_cur += D.
    double vj[SLOTS];
    int jc[SLOTS];
    bool colok[SLOTS];
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) { colok[t] = lane + 32 * t < nc; jc[t] = min(lane + 32 * t, nc - 1); vj[t] = w.v[jc[t]]; }
    double *Drow = w.shortest;                                  // distance of every row (through its matched column); [cur] = 0
    for (int cur = 0; cur < nr; ++cur) {
        if (w.col4row[cur] != -1) continue;                     // warp-uniform
        double d[SLOTS];
        int pth[SLOTS], r4c[SLOTS];
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) { d[t] = INFINITY; pth[t] = -1; r4c[t] = w.row4col[jc[t]]; }
        if (lane < nr) Drow[lane] = lane == cur ? 0.0 : INFINITY;
        __syncwarp();
        unsigned active = 1u << cur;
        double best_free = INFINITY;
        int rounds = 0;
        while (active) {
            if (++rounds > 2 * LAP_MAX_DIM) return false;       // cannot happen (distances only decrease); never loop forever
            bool chg[SLOTS];
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) chg[t] = false;
            unsigned a = active;
            while (a) {
                const int i = __ffs(a) - 1;
                a &= a - 1;
                ++steps;
                // reduced costs are >= 0 in exact arithmetic; clamping the rounding noise (-1e-17) rules out negative cycles, on
                // which a label-correcting search would improve forever by one ulp per round
                const double Di = Drow[i], ui = w.u[i];
#pragma unroll
                for (int t = 0; t < SLOTS; ++t) {
                    const double r = fmax((cost(i, jc[t]) - vj[t]) - ui, 0.0) + Di;
                    const bool upd = colok[t] && (r < d[t]);
                    d[t] = upd ? r : d[t];
                    pth[t] = upd ? i : pth[t];
                    chg[t] = chg[t] || upd;
                }
            }
            // best distance of a free column (64-bit min in two 32-bit reductions; distances are >= 0 up to rounding)
            double bf = INFINITY;
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) if (colok[t] && r4c[t] < 0) bf = fmin(bf, d[t]);
            const int hi = __double2hiint(bf);
            const int mhi = __reduce_min_sync(TTDG_FULL, hi);
            const unsigned mlo = __reduce_min_sync(TTDG_FULL, hi == mhi ? (unsigned)__double2loint(bf) : 0xFFFFFFFFu);
            best_free = __hiloint2double(mhi, (int)mlo);
            // rows to expand next: matched rows of the columns that improved and are still closer than the best free column
            unsigned bits = 0u;
#pragma unroll
            for (int t = 0; t < SLOTS; ++t)
                if (chg[t] && r4c[t] >= 0 && d[t] < best_free) { bits |= 1u << r4c[t]; Drow[r4c[t]] = d[t]; }
            __syncwarp();
            active = __reduce_or_sync(TTDG_FULL, bits);
        }
        if (!(best_free < INFINITY)) return false;              // no free column reachable: impossible for nr <= nc
        // sink = a free column at the best distance
        int mine = -1;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) if (mine < 0 && colok[t] && r4c[t] < 0 && d[t] == best_free) mine = lane + 32 * t;
        const unsigned who = __ballot_sync(TTDG_FULL, mine >= 0);
        if (who == 0u) return false;
        const int sink = __shfl_sync(TTDG_FULL, mine, __ffs(who) - 1);
        const double minVal = best_free;
        if (lane == 0) w.u[cur] += minVal;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t)
            if (colok[t] && d[t] < INFINITY) {
                w.path[lane + 32 * t] = pth[t];
                if (d[t] < minVal && r4c[t] >= 0) {
                    const double dd = minVal - d[t];
                    w.u[r4c[t]] += dd;
                    vj[t] -= dd;
                }
            }
        __syncwarp();
        if (lane == 0) {                                        // augment along the path (sequential, short)
            int j = sink, guard = 0;
            while (true) {
                ++hops;
                const int r = w.path[j];
                w.row4col[j] = r;
                const int t = w.col4row[r];
                w.col4row[r] = j;
                j = t;
                if (r == cur || ++guard > LAP_MAX_DIM) break;
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) if (lane + 32 * t < nc) w.v[lane + 32 * t] = vj[t];
    }
    else {
    // ---- (B) lean Dijkstra augmentations for the rows still free, lane = column
    double vj[SLOTS];
    int jc[SLOTS];
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) { jc[t] = min(lane + 32 * t, nc - 1); vj[t] = w.v[jc[t]]; }
    for (int cur = 0; cur < nr; ++cur) {
        if (w.col4row[cur] != -1) continue;                     // warp-uniform
        double sh[SLOTS];
        int pth[SLOTS], r4c[SLOTS];
        bool live[SLOTS], scanned[SLOTS];
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            const int j = lane + 32 * t;
            sh[t] = INFINITY; pth[t] = -1; scanned[t] = false;
            live[t] = j < nc;
            r4c[t] = w.row4col[jc[t]];
        }
        double minVal = 0.0;
        int sink = -1, i = cur;
        while (true) {
            ++steps;
            const double tv = minVal - w.u[i];
            double best = INFINITY;
            int bpk = 0;                                        // (row4col + 1) << 2 | slot of the lane-local best column
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) {
                const double r = (cost(i, jc[t]) - vj[t]) + tv;
                const bool upd = live[t] && (r < sh[t]);
                sh[t] = upd ? r : sh[t];
                pth[t] = upd ? i : pth[t];
                const double cand = live[t] ? sh[t] : INFINITY;
                const bool better = cand < best;
                best = better ? cand : best;
                bpk = better ? (((r4c[t] + 1) << 2) | t) : bpk;
            }
            // arg-min over the warp in two 32-bit reductions, no ballot / shuffle: distances are >= 0 up to rounding, so the high
            // word orders them as a signed integer; among the lanes that hold the minimal high word the second reduction takes
            // the low word with its 7 lowest bits replaced by (slot, lane) - a 2^-45 relative perturbation of the comparison,
            // far below the certificate's margin - so its result names the winning lane and slot AND carries minVal's low word
            const int hi = __double2hiint(best);
            const int mhi = __reduce_min_sync(TTDG_FULL, hi);
            const unsigned k2 = (hi == mhi) ? (((unsigned)__double2loint(best) & ~127u) | (unsigned)((bpk & 3) << 5) | (unsigned)lane) : 0xFFFFFFFFu;
            const unsigned m2 = __reduce_min_sync(TTDG_FULL, k2);
            const int src = (int)(m2 & 31u), ssel = (int)((m2 >> 5) & 3u);
            minVal = __hiloint2double(mhi, (int)(m2 & ~127u));
            const int jsel = src + 32 * ssel;
            const int rsel = (mhi == 0x7FF00000) ? -2 : w.row4col[jsel];         // smem read by all lanes (broadcast)
            if (rsel == -2) return false;                       // cannot happen with finite costs; never loop forever
#pragma unroll
            for (int t = 0; t < SLOTS; ++t)
                if (lane == src && t == ssel) { live[t] = false; scanned[t] = true; sh[t] = minVal; }
            if (rsel < 0) { sink = jsel; break; }
            i = rsel;
        }
        // dual update: u of the visited rows by the owner of their (scanned) column, v in registers; path of the scanned columns
        if (lane == 0) w.u[cur] += minVal;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t)
            if (scanned[t]) {
                const double d = minVal - sh[t];
                if (r4c[t] >= 0) w.u[r4c[t]] += d;
                vj[t] -= d;
                w.path[lane + 32 * t] = pth[t];
            }
        __syncwarp();
        if (lane == 0) {                                        // augment along the path (sequential, short)
            int j = sink;
            while (true) {
                ++hops;
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
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) if (lane + 32 * t < nc) w.v[lane + 32 * t] = vj[t];
    }
    if (lane == 0) { w.stat_steps += steps; w.stat_hops += hops; }
    __syncwarp();
    const long long ck2 = clock64();
    const bool ok = lap_certificate(nr, nc, cost, w);
    if (lane == 0) {
        if (ok) ++w.stat_fast_ok; else ++w.stat_fast_fallback;
        w.stat_ck[0] += ck1 - ck0 - ck_res; w.stat_ck[1] += ck_res; w.stat_ck[2] += ck2 - ck1; w.stat_ck[3] += clock64() - ck2;
    }
    __syncwarp();
    return ok;
}

// lean certified solve; false = not certified (the caller runs the SciPy-order solve).  Full warp, nr <= 32.
// bf = false: Dijkstra augmentations (one column settled per step); true: label-correcting rounds (all improved rows relaxed per
// round, one reduction per round)
template <class Cost>
__device__ __forceinline__ bool lap_lean_warp(int nr, int nc, Cost cost, LapWork &w, bool bf) {
    if (nr > 32 || nc > LAP_MAX_DIM) return false;
    if (bf) {
        if (nc <= 32) return lap_lean_warp_t<1, true>(nr, nc, cost, w);
        if (nc <= 64) return lap_lean_warp_t<2, true>(nr, nc, cost, w);
        return lap_lean_warp_t<LAP_SLOTS, true>(nr, nc, cost, w);
    }
    if (nc <= 32) return lap_lean_warp_t<1, false>(nr, nc, cost, w);
    if (nc <= 64) return lap_lean_warp_t<2, false>(nr, nc, cost, w);
    return lap_lean_warp_t<LAP_SLOTS, false>(nr, nc, cost, w);
}

// fast != 0: try the certified row-reduction solve first (rows <= 32 only), fall back to the SciPy-order solve
template <class Cost>
__device__ __forceinline__ void lap_solve_warp(int nr, int nc, Cost cost, LapWork &w, int fast = 0) {
    if ((fast == 1 || fast == 2) && nr <= 32) {
        if ((threadIdx.x & 31) == 0) w.fast_mode = fast;
        __syncwarp();
        bool ok;
        if (nc <= 32) ok = lap_solve_warp_t<1, true>(nr, nc, cost, w);
        else if (nc <= 64) ok = lap_solve_warp_t<2, true>(nr, nc, cost, w);
        else ok = lap_solve_warp_t<LAP_SLOTS, true>(nr, nc, cost, w);
        if (ok) return;
    }
    if (nc <= 32) lap_solve_warp_t<1, false>(nr, nc, cost, w);
    else if (nc <= 64) lap_solve_warp_t<2, false>(nr, nc, cost, w);
    else lap_solve_warp_t<LAP_SLOTS, false>(nr, nc, cost, w);
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
