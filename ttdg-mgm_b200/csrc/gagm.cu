// gagm.cu - the graduated-assignment multi-graph-matching solver as ONE persistent thread-block cluster.
//
// Reference: GA_GM.forward + GA_GM.gagm (adapteacher/modeling/GModule/multi_graph_matching.py:223-244, 300-389)
// with num_clusters = 1 (cluster weight == 1), projector0 = 'sinkhorn', hung_iter = True - the only
// configuration MGM3_unsup uses (mgm:469-474, 533).  Per iteration the reference does
//     UUt = U U^T;  V = chain_matmul(A, UUt, A, U) * quad_weight * 2 + W U;  V /= G          (mgm:317-321)
//     U   = batched Sinkhorn(V / tau)   or   per-graph Hungarian(V)                             (mgm:324-353)
//     G == 2: U[:ms0] = eye (mgm:358-359);  stop on ||U - lastU|| < tol or ||U - lastU2|| == 0   (mgm:361)
// i.e. ~213 iterations of 4 small GEMMs + a projector + 2 norms, each norm a host sync, each Hungarian a
// device->host->device round trip (~800 per step).  Here: zero host involvement.
//
// Mapping: cluster of C = min(G, 8) CTAs; CTA c owns graphs c, c + C, ...  A is block diagonal, so with
// X_g = A_gg U_g the chain is  A (U (U^T (A U))) = A_gg (U_g T),  T = sum_g U_g^T X_g  (32 x 32):
//   phase 1: X_g, partial T  -> global scratch  | cluster barrier
//   phase 2: T, Q_g = U_g T, V_g = (2 qw A_gg Q_g + W[g,:] U) / G, projector, U_new_g, partial norms | cluster barrier
// All arithmetic is fp64 (A, W, U0 are fp32 inputs widened exactly): the iteration is a discrete dynamical
// system that amplifies rounding noise, fp64 makes it independent of summation order.
//
// Hungarian stage (200 of the ~213 iterations): U_t is a 0/1 partial permutation, described by node_of[g][u] (the node
// of graph g sitting in universe slot u, or -1).  Then X_g = A_gg U_g, T and W U are gathers:
//     T[u1][u2] = sum_h A_hh[node_h(u1)][node_h(u2)]         (every CTA builds the whole T itself: no phase-1 barrier)
//     Q_g[r][:] = T[slot_g(r)][:],   (W U)[r][u] = sum_h W[r][node_h(u)]
// and only V1 = A_gg Q_g stays dense: one cluster barrier per iteration, ~1/8 of the FMAs.  Zero terms are skipped in
// the same order the dense loops add them, so both paths produce the same bits (for G <= 8).
//
// Round 2 (cycle accounting through ttdg_gagm_read_profile):
//   * G <= 8 (one graph per CTA): node_of and the norm partials travel between the CTAs through DSMEM stores before the cluster
//     barrier (no L2 round trips, no __threadfence), the norms ||U' - U||^2 are counted from node_of (0 / 1 entries: exact), the
//     diagonal blocks of A and the CTA's rows of W are read from fp32 copies in shared memory when they fit, V1 from an fp64 copy;
//   * Sinkhorn-stage iterations (dense U): X = A_gg U_g, V1 = A_gg Q and V2 = W U run on the FP64 tensor cores (mma.sync m8n8k4
//     f64, one 8 x 8 tile per warp) instead of one warp-uniform load + F2F per DFMA;
//   * ops.SOLVER_WINDOW_HOOK (host side): the ~10 ms this kernel keeps <= 8 SMs busy are filled with the previous dataset's
//     evaluation pass on another stream (adapteacher/engine/trainer.py).
#include <cstdlib>
#include "lap.cuh"
#include "sinkhorn_small.cuh"
#include <cooperative_groups.h>

namespace ttdg {

constexpr int GAGM_THREADS = 512;
constexpr int GAGM_WARPS = GAGM_THREADS / 32;
constexpr int GAGM_MAX_N = 96;            // nodes per graph
constexpr int GAGM_MAX_G = 64;
constexpr int GAGM_MAX_C = 8;             // portable cluster size
constexpr int NU = 32;                    // universe size (rcnn.py:116)
constexpr int RPW = GAGM_MAX_N / GAGM_WARPS;   // rows per warp = 6
constexpr int ZP = NU + 1;                // pitch of the V / z tile

struct GagmParams {
    const float *A, *W, *U0;
    float *U_out;
    int32_t *info;
    double *Ubuf;        // 3 x M x NU
    double *Tpart;       // C x NU*NU
    double *normpart;    // C x 2
    double *trace;       // optional: (trace_cap + 1) x M x NU, U_t of every iteration (tests)
    double *trace_meta;  // optional: trace_cap x 2 = {projector, tau}
    int32_t *nodeof;     // 2 x GAGM_MAX_G x NU: node_of of U_t for Hungarian-stage iterations, double-buffered by iteration parity
    int trace_cap;
    int G, M, C;
    double init_tau, min_tau, sk_gamma, tol, quad_weight;
    int max_iter, sk_iter, mode, step_projector, sq_transposed;
    int uall;            // shared memory holds U of all graphs: 1 = pitch NU (vector product), 2 = pitch GAGM_UP (FP64 tensor-core product)
    int hfast;           // G <= cluster size (one graph per CTA): Hungarian-stage iterations exchange node_of / norms through DSMEM,
                         // count the norms, keep an fp64 copy of the CTA's own block of A in shared memory
    int hcache;          // hfast and the diagonal blocks of A + this CTA's rows of W fit in shared memory as fp32 copies (the region
                         // U-all used during the Sinkhorn stage): the gathers of T and W U read them instead of global memory
    int aoff[GAGM_MAX_C + 1];   // hfast: float offset of graph h's n_h x n_h block in the shared copy
    int lap_fast;        // TTDG_LAP_FAST=1: certified row-reduction LAP first, SciPy-order solve as the fall-back (lap.cuh)
    int node_off[GAGM_MAX_G + 1];
};

__device__ long long g_gagm_prof[24];      // CTA 0 / thread 0: cycles per segment of the Hungarian-stage iterations (ttdg_gagm_read_profile)

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// acc[r] += sum_j Mg[(warp + 16 r) * ld + j] * S[j * NU + lane],  r < R = ceil(nrows / 16); rows >= nrows are
// computed on row 0 and ignored by the caller
template <int R>
__device__ __forceinline__ void rows_times_tile_r(const float *__restrict__ Mg, int ld, int nrows, int ncols,
                                                  const double *__restrict__ S, double (&acc)[RPW], int warp, int lane) {
    const float *rp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = warp + GAGM_WARPS * r;
        rp[r] = Mg + (size_t)(i < nrows ? i : 0) * ld;
    }
    // 8 columns per batch for the narrow variant: the rows of W come from L2 (~700 cycles), the batch's loads are in flight together
#pragma unroll(R <= 3 ? 8 : 4)
    for (int j = 0; j < ncols; ++j) {
        const double sv = S[j * NU + lane];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fma((double)__ldg(rp[r] + j), sv, acc[r]);
    }
}
// the same product with an fp64 copy of the matrix in shared memory (pitch = nrows = ncols = n): no widening at all
template <int R>
__device__ __forceinline__ void rows_times_tile_d_r(const double *Md, int n, const double *S, double (&acc)[RPW], int warp, int lane) {
    const double *rp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { const int i = warp + GAGM_WARPS * r; rp[r] = Md + (i < n ? i : 0) * n; }
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
        const double sv = S[j * NU + lane];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fma(rp[r][j], sv, acc[r]);
    }
}
__device__ __forceinline__ void rows_times_tile_d(const double *Md, int n, const double *S, double (&acc)[RPW], int warp, int lane) {
    switch ((n + GAGM_WARPS - 1) / GAGM_WARPS) {
        case 1: rows_times_tile_d_r<1>(Md, n, S, acc, warp, lane); break;
        case 2: rows_times_tile_d_r<2>(Md, n, S, acc, warp, lane); break;
        case 3: rows_times_tile_d_r<3>(Md, n, S, acc, warp, lane); break;
        default: rows_times_tile_d_r<4>(Md, n, S, acc, warp, lane); break;      // n * n <= GAGM_MAX_N * NU: n <= 55
    }
}
// Two row-count variants only (<= 48 and <= 96 nodes; surplus rows repeat row 0 and are dropped): with six variants inlined at four
// call sites the kernel's code grew enough to slow the LAP's loops down (measured), and an out-of-line copy keeps the accumulators
// in local memory (3x slower).
__device__ __forceinline__ void rows_times_tile(const float *__restrict__ Mg, int ld, int nrows, int ncols,
                                                const double *__restrict__ S, double (&acc)[RPW], int warp, int lane) {
    if (nrows <= 3 * GAGM_WARPS) rows_times_tile_r<3>(Mg, ld, nrows, ncols, S, acc, warp, lane);
    else rows_times_tile_r<6>(Mg, ld, nrows, ncols, S, acc, warp, lane);
}

// V2 = W[rows of one graph][0 .. M) (n x M fp32, global) x U (M x 32 fp64, shared, pitch GAGM_UP) on the FP64 tensor cores
// (mma.sync m8n8k4 f64).  The vector form pays one warp-uniform global load and one F2F per row and column for every DFMA - LSU
// issue bound, 50-90 k cycles per Sinkhorn-stage iteration, and a temperature stage that does not converge runs 200 of them.
// Here a warp owns 8 x 8 output tiles: per k-step of 4 each lane loads ONE entry of W (the fragment's, widened once) and ONE of U
// and the warp issues one DMMA = 256 FMAs; GAGM_UP = 36 makes the U fragment loads conflict-free (row step 288 B: the 16 lanes of
// a half-warp hit 16 different 8-byte banks).  fp64 accumulation in another order than the vector form: ~1e-16 apart.
constexpr int GAGM_UP = 36;
// (c0, c1) += A[8 rows][0 .. K) x B[0 .. K)[8 columns]: wr = this lane's row of the fp32 matrix + its fragment column, ub = the
// fp64 operand at [fragment row][this lane's column], pitch UPITCH doubles
template <int UPITCH>
__device__ __forceinline__ void dmma_tile(const float *__restrict__ wr, int K, const double *ub, int ac, double &c0, double &c1) {
    constexpr int KB = 8;                                   // k-steps per batch: the batch's global loads are in flight together
    int k = 0;
    for (; k + 4 * KB <= K; k += 4 * KB) {
        float a[KB];
#pragma unroll
        for (int q = 0; q < KB; ++q) a[q] = __ldg(wr + k + 4 * q);
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            const double av = (double)a[q], bv = ub[(k + 4 * q) * UPITCH];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
        }
    }
    for (; k < K; k += 4) {
        const bool in = k + ac < K;
        const double av = in ? (double)__ldg(wr + k) : 0.0, bv = in ? ub[k * UPITCH] : 0.0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
    }
}
// out[n x 32, pitch OP] = Mg[n rows][0 .. K) (fp32, global, leading dimension ld) x B (K x 32 fp64, shared, pitch NU): the small
// products of the Sinkhorn-stage iterations (X = A_gg U_g) as 8 x 8 tiles, one warp per tile
__device__ __forceinline__ void rows_times_tile_tc(const float *__restrict__ Mg, int ld, int n, int K, const double *B, double *out,
                                                   int op, int warp, int lane) {
    const int ar = lane >> 2, ac = lane & 3;                // fragment coordinates: A[ar][ac], B[ac][ar], C[ar][2 ac + {0, 1}]
    const int ntiles = ((n + 7) >> 3) * (NU / 8);
    for (int t = warp; t < ntiles; t += GAGM_WARPS) {
        const int rb = t / (NU / 8), nt = t % (NU / 8);
        const int row = rb * 8 + ar;
        double c0 = 0.0, c1 = 0.0;
        dmma_tile<NU>(Mg + (size_t)(row < n ? row : 0) * ld + ac, K, B + ac * NU + nt * 8 + ar, ac, c0, c1);
        if (row < n) { out[row * op + nt * 8 + 2 * ac] = c0; out[row * op + nt * 8 + 2 * ac + 1] = c1; }
    }
}

// bytes of the fixed part of the dynamic shared memory (everything before the U-all / Hungarian-stage cache region), 16-aligned
constexpr size_t GAGM_FIXED_SMEM =
    (((size_t)(3 * GAGM_MAX_N * NU + GAGM_MAX_N * ZP + NU * NU + 2 * GAGM_MAX_N + 2 * GAGM_WARPS + NU + 8 + 4 * GAGM_MAX_C) * sizeof(double) +
      sizeof(LapWork) + (size_t)(GAGM_MAX_G * NU + GAGM_MAX_N) * sizeof(int)) + 15) & ~(size_t)15;

__global__ void __launch_bounds__(GAGM_THREADS, 1)
gagm_kernel(const __grid_constant__ GagmParams p) {
    extern __shared__ __align__(16) unsigned char gsm[];
    double *Ug = reinterpret_cast<double *>(gsm);          // n x NU   (U_g, later U_new_g)
    double *X = Ug + GAGM_MAX_N * NU;                      // n x NU   (A U, then Q = U T)
    double *Uo = X + GAGM_MAX_N * NU;                      // n x NU   (U_h tile for W U)
    double *Z = Uo + GAGM_MAX_N * NU;                      // n x ZP   (V, then Sinkhorn log-matrix)
    double *T = Z + GAGM_MAX_N * ZP;                       // NU x NU
    double *padv = T + NU * NU;                            // GAGM_MAX_N
    double *red = padv + GAGM_MAX_N;                       // 2 * GAGM_WARPS
    double *avec = red + 2 * GAGM_WARPS;                   // NU + 8: row scalings of the Sinkhorn projector (+ the dummy rows' one)
    double *bvec = avec + NU + 8;                          // GAGM_MAX_N: its column scalings
    double *normx = bvec + GAGM_MAX_N;                     // 2 x 2 GAGM_MAX_C (hfast): every CTA's norm partials, by iteration parity
    LapWork *lapw = reinterpret_cast<LapWork *>(normx + 4 * GAGM_MAX_C);
    int *nodeof_s = reinterpret_cast<int *>(lapw + 1);      // G x NU: node_of of every graph (Hungarian stage)
    int *slot_s = nodeof_s + GAGM_MAX_G * NU;               // GAGM_MAX_N: universe slot of each node of the current graph
    // optional: U_t of ALL graphs (M x NU fp64) for the dense W U product of the Sinkhorn-stage iterations, when it fits
    // (carved at a compile-time offset: pointer arithmetic through an integer would hide the shared address space - generic loads)
    double *Uall = reinterpret_cast<double *>(gsm + GAGM_FIXED_SMEM);
    // hfast, Hungarian stage (U-all is not needed any more): the diagonal blocks of A and this CTA's rows of W as fp32 copies
    float *Ad_s = reinterpret_cast<float *>(Uall);
    float *W_s = Ad_s + ((p.aoff[p.G <= GAGM_MAX_C ? p.G : 0] + 3) & ~3);

    const int c = blockIdx.x, C = p.C, G = p.G, M = p.M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t UB = (size_t)M * NU;

    // ---- init: U buffers
    for (int g = c; g < G; g += C) {
        const int o = p.node_off[g], n = p.node_off[g + 1] - o;
        for (int e = tid; e < n * NU; e += GAGM_THREADS) {
            p.Ubuf[(size_t)o * NU + e] = (double)p.U0[(size_t)o * NU + e];
            p.Ubuf[UB + (size_t)o * NU + e] = 0.0;           // lastU = zeros_like(U)  (mgm:305)
            p.Ubuf[2 * UB + (size_t)o * NU + e] = 0.0;
        }
    }
    __threadfence();
    cluster_sync_all();

    if (tid == 0) {
        lapw->stat_steps = 0; lapw->stat_hops = 0; lapw->stat_fast_ok = 0; lapw->stat_fast_fallback = 0;
        for (int i = 0; i < 6; ++i) lapw->stat_ck[i] = 0;
        for (int i = 0; i < 8; ++i) lapw->stat_free[i] = 0;
    }
    // cycle accounting of CTA 0 / thread 0 (info[8..12], units of 1024 cycles): whole kernel, Hungarian-stage iterations, LAP, cluster-barrier waits
    const long long ck_start = clock64();
    long long ck_hung = 0, ck_lap = 0, ck_bar = 0;
    long long ck_p1 = 0, ck_v = 0, ck_proj = 0;             // Sinkhorn-stage iterations: phase 1 + its barrier, V = chain + W U, projector
    long long hs[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // Hungarian-stage segments
    long long hk = 0;
#define HSEG(i) do { const long long now_ = clock64(); if (binU) hs[i] += now_ - hk; hk = now_; } while (0)
    int cur = 0, last = 1, last2 = 2;
    double tau = p.init_tau;
    int projector = (p.mode == 1) ? p.step_projector : 0;   // 0 sinkhorn, 1 hungarian
    int it_total = 0, it_sk = 0, it_hg = 0, n_lap = 0, n_stage = 0;
    bool binU = false;                                      // U_t came from a Hungarian projection: node_of[it_total & 1] is valid
    const bool hf = p.hfast != 0;                           // then g == c is this CTA's only graph
    const bool gp2 = (G & (G - 1)) == 0;
    const double invG = 1.0 / (double)G;
    bool cache_ready = false, last_lean = false;
    int nd_cur = -1, nd_prev = -1;                          // hfast, tid < NU: node of this CTA's graph in slot tid under U_t / U_{t-1}
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();

    while (true) {
        bool stop_all = false;
        for (int i = 0; i < p.max_iter; ++i) {
            { const int nxt = last2; last2 = last; last = cur; cur = nxt; }   // lastU2 = lastU; lastU = U (mgm:313-314)
            const long long ck_it = clock64();
            long long ck_ph2 = ck_it;
            const double *Ul = p.Ubuf + (size_t)last * UB;      // U_t
            const double *Ul2 = p.Ubuf + (size_t)last2 * UB;    // U_{t-1}
            double *Un = p.Ubuf + (size_t)cur * UB;             // U_{t+1}

            const int e0 = 2 * tid, u1 = e0 / NU, u2 = e0 % NU;
            const int32_t *nodeof_cur = p.nodeof + (size_t)(it_total & 1) * GAGM_MAX_G * NU;
            int32_t *nodeof_nxt = p.nodeof + (size_t)((it_total + 1) & 1) * GAGM_MAX_G * NU;
            // hfast: node_of of every graph arrives in shared memory (written by its owner through DSMEM before the last barrier)
            const int *nos = hf ? nodeof_s + (it_total & 1) * (GAGM_MAX_C * NU) : nodeof_s;
            if (binU && hf) {
                // ================= Hungarian stage without global memory
                const int o_c = p.node_off[c], n_c = p.node_off[c + 1] - o_c;
                const bool a64 = n_c * n_c <= GAGM_MAX_N * NU;     // fp64 copy of the own block of A in the (idle) U_h tile
                if (!cache_ready) {
                    if (p.hcache) {
                        for (int h = 0; h < G; ++h) {
                            const int oh = p.node_off[h], nh = p.node_off[h + 1] - oh;
                            for (int e = tid; e < nh * nh; e += GAGM_THREADS)
                                Ad_s[p.aoff[h] + e] = __ldg(p.A + (size_t)(oh + e / nh) * M + oh + e % nh);
                        }
                        for (int e = tid; e < n_c * M; e += GAGM_THREADS) W_s[e] = __ldg(p.W + (size_t)o_c * M + e);
                    }
                    if (a64)
                        for (int e = tid; e < n_c * n_c; e += GAGM_THREADS)
                            Uo[e] = (double)__ldg(p.A + (size_t)(o_c + e / n_c) * M + o_c + e % n_c);
                    cache_ready = true;
                    __syncthreads();
                }
                double t0 = 0.0, t1 = 0.0;
                float a2v[GAGM_MAX_C], a3v[GAGM_MAX_C];            // branch-free: the eight gathers overlap; absent terms add + 0.0 (exact)
#pragma unroll
                for (int h = 0; h < GAGM_MAX_C; ++h) {
                    a2v[h] = 0.f; a3v[h] = 0.f;
                    if (h < G) {
                        const int oh = p.node_off[h], nh = p.node_off[h + 1] - oh;
                        const int n1 = nos[h * NU + u1], n2 = nos[h * NU + u2], n3 = nos[h * NU + u2 + 1];
                        float a2, a3;
                        if (p.hcache) { const float *arow = Ad_s + p.aoff[h] + max(n1, 0) * nh; a2 = arow[max(n2, 0)]; a3 = arow[max(n3, 0)]; }
                        else { const float *arow = p.A + (size_t)(oh + max(n1, 0)) * M + oh; a2 = __ldg(arow + max(n2, 0)); a3 = __ldg(arow + max(n3, 0)); }
                        a2v[h] = (n1 >= 0 && n2 >= 0) ? a2 : 0.f;
                        a3v[h] = (n1 >= 0 && n3 >= 0) ? a3 : 0.f;
                    }
                }
#pragma unroll
                for (int h = 0; h < GAGM_MAX_C; ++h) { t0 += (double)a2v[h]; t1 += (double)a3v[h]; }
                T[e0] = t0; T[e0 + 1] = t1;
                hk = ck_it; HSEG(0);
            } else if (binU) {
                // ================= Hungarian stage: T from the permutation description, no barrier
                for (int e = tid; e < G * NU; e += GAGM_THREADS) nodeof_s[e] = __ldcg(nodeof_cur + e);
                __syncthreads();
                double t0 = 0.0, t1 = 0.0;
                for (int h = 0; h < G; ++h) {
                    const int n1 = nodeof_s[h * NU + u1];
                    if (n1 >= 0) {
                        const float *arow = p.A + (size_t)(p.node_off[h] + n1) * M + p.node_off[h];
                        const int n2 = nodeof_s[h * NU + u2], n3 = nodeof_s[h * NU + u2 + 1];
                        if (n2 >= 0) t0 += (double)__ldg(arow + n2);
                        if (n3 >= 0) t1 += (double)__ldg(arow + n3);
                    }
                }
                T[e0] = t0; T[e0 + 1] = t1;
                hk = ck_it; HSEG(0);
            } else {
            // ================= phase 1: X_g = A_gg U_g, partial T = sum_g U_g^T X_g
            double t0 = 0.0, t1 = 0.0;
            for (int g = c; g < G; g += C) {
                const int o = p.node_off[g], n = p.node_off[g + 1] - o;
                for (int e = tid; e < n * NU; e += GAGM_THREADS) Ug[e] = __ldcg(Ul + (size_t)o * NU + e);
                __syncthreads();
                if (p.uall == 2) rows_times_tile_tc(p.A + (size_t)o * M + o, M, n, n, Ug, X, NU, warp, lane);
                else {
                    double acc[RPW];
#pragma unroll
                    for (int r = 0; r < RPW; ++r) acc[r] = 0.0;
                    rows_times_tile(p.A + (size_t)o * M + o, M, n, n, Ug, acc, warp, lane);
#pragma unroll
                    for (int r = 0; r < RPW; ++r) { const int row = warp + GAGM_WARPS * r; if (row < n) X[row * NU + lane] = acc[r]; }
                }
                __syncthreads();
                for (int r = 0; r < n; ++r) {
                    const double uv = Ug[r * NU + u1];
                    t0 = fma(uv, X[r * NU + u2], t0);
                    t1 = fma(uv, X[r * NU + u2 + 1], t1);
                }
                __syncthreads();
            }
            p.Tpart[(size_t)c * NU * NU + e0] = t0;
            p.Tpart[(size_t)c * NU * NU + e0 + 1] = t1;
            __threadfence();
            cluster_sync_all();
            ck_p1 += clock64() - ck_it;
            ck_ph2 = clock64();

            // ================= phase 2
            {
                double s0 = 0.0, s1 = 0.0;
                double2 tp[GAGM_MAX_C];                               // all partial products in flight, summed in CTA order
#pragma unroll
                for (int cc = 0; cc < GAGM_MAX_C; ++cc)
                    tp[cc] = cc < C ? __ldcg(reinterpret_cast<const double2 *>(p.Tpart + (size_t)cc * NU * NU + e0)) : make_double2(0.0, 0.0);
#pragma unroll
                for (int cc = 0; cc < GAGM_MAX_C; ++cc) if (cc < C) { s0 += tp[cc].x; s1 += tp[cc].y; }
                T[e0] = s0; T[e0 + 1] = s1;
            }
            }
            double d1 = 0.0, d2 = 0.0;
            for (int g = c; g < G; g += C) {
                const int o = p.node_off[g], n = p.node_off[g + 1] - o;
                if (binU) {
                    for (int r = tid; r < n; r += GAGM_THREADS) slot_s[r] = -1;
                    __syncthreads();                               // also publishes T
                    if (tid < NU) { const int nd = nos[g * NU + tid]; if (nd >= 0) slot_s[nd] = tid; }
                    __syncthreads();
                    // Q = U_g T = rows of T picked by the node's slot
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const int row = warp + GAGM_WARPS * r;
                        if (row < n) { const int sl = slot_s[row]; X[row * NU + lane] = sl >= 0 ? T[sl * NU + lane] : 0.0; }
                    }
                } else {
                for (int e = tid; e < n * NU; e += GAGM_THREADS) Ug[e] = __ldcg(Ul + (size_t)o * NU + e);
                __syncthreads();                                   // also publishes T
                // Q = U_g T
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int row = warp + GAGM_WARPS * r;
                    if (row < n) {
                        double q = 0.0;
                        for (int v = 0; v < NU; ++v) q = fma(Ug[row * NU + v], T[v * NU + lane], q);
                        X[row * NU + lane] = q;
                    }
                }
                }
                __syncthreads();
                HSEG(1);
                // V1 = A_gg Q
                double v1[RPW], v2[RPW];
#pragma unroll
                for (int r = 0; r < RPW; ++r) { v1[r] = 0.0; v2[r] = 0.0; }
                const bool tc = !binU && p.uall == 2;              // dense U_t: both products of V on the FP64 tensor cores
                if (tc) {
                    for (int e0_ = tid; e0_ < M * NU; e0_ += 8 * GAGM_THREADS) {       // U_t of all graphs, pitch GAGM_UP
                        double uv[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) { const int e = e0_ + q * GAGM_THREADS; uv[q] = e < M * NU ? __ldcg(Ul + e) : 0.0; }
#pragma unroll
                        for (int q = 0; q < 8; ++q) { const int e = e0_ + q * GAGM_THREADS; if (e < M * NU) Uall[(e / NU) * GAGM_UP + e % NU] = uv[q]; }
                    }
                    __syncthreads();
                    const int ar = lane >> 2, ac = lane & 3;
                    const int ntiles = ((n + 7) >> 3) * (NU / 8);
                    for (int t = warp; t < ntiles; t += GAGM_WARPS) {
                        const int rb = t / (NU / 8), nt = t % (NU / 8);
                        const int row = rb * 8 + ar, rowc = row < n ? row : 0;
                        double a0 = 0.0, a1 = 0.0, w0 = 0.0, w1 = 0.0;
                        dmma_tile<NU>(p.A + (size_t)(o + rowc) * M + o + ac, n, X + ac * NU + nt * 8 + ar, ac, a0, a1);          // V1 = A_gg Q
                        dmma_tile<GAGM_UP>(p.W + (size_t)(o + rowc) * M + ac, M, Uall + ac * GAGM_UP + nt * 8 + ar, ac, w0, w1);  // V2 = W U
                        if (row < n) {
                            const double s0 = a0 * p.quad_weight * 2.0 + w0, s1 = a1 * p.quad_weight * 2.0 + w1;
                            const double q0 = gp2 ? s0 * invG : s0 / (double)G, q1 = gp2 ? s1 * invG : s1 / (double)G;
                            double *z = Z + row * ZP + nt * 8 + 2 * ac;
                            z[0] = projector == 0 ? q0 / tau : q0; z[1] = projector == 0 ? q1 / tau : q1;
                        }
                    }
                } else
                if (binU && hf && n * n <= GAGM_MAX_N * NU) rows_times_tile_d(Uo, n, X, v1, warp, lane);
                else rows_times_tile(p.A + (size_t)o * M + o, M, n, n, X, v1, warp, lane);
                HSEG(2);
                // V2 = W[g rows, :] U   (Hungarian stage: one entry of W per graph h; else tile by graph h)
                if (tc) {
                    // done above, V is in place
                } else if (binU && hf) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const int row = warp + GAGM_WARPS * r;
                        if (row < n) {
                            const float *wrow_s = W_s + row * M, *wrow_g = p.W + (size_t)(o + row) * M;
                            double acc = 0.0;
                            float wv[GAGM_MAX_C];
#pragma unroll
                            for (int h = 0; h < GAGM_MAX_C; ++h) {
                                wv[h] = 0.f;
                                if (h < G) {
                                    const int nd = nos[h * NU + lane];
                                    const float w = p.hcache ? wrow_s[p.node_off[h] + max(nd, 0)] : __ldg(wrow_g + p.node_off[h] + max(nd, 0));
                                    wv[h] = nd >= 0 ? w : 0.f;
                                }
                            }
#pragma unroll
                            for (int h = 0; h < GAGM_MAX_C; ++h) acc += (double)wv[h];
                            v2[r] = acc;
                        }
                    }
                } else if (binU) {
#pragma unroll
                    for (int r = 0; r < RPW; ++r) {
                        const int row = warp + GAGM_WARPS * r;
                        if (row < n) {
                            const float *wrow = p.W + (size_t)(o + row) * M;
                            double acc = 0.0;
                            for (int h = 0; h < G; ++h) {
                                const int nd = nodeof_s[h * NU + lane];
                                if (nd >= 0) acc += (double)__ldg(wrow + p.node_off[h] + nd);
                            }
                            v2[r] = acc;
                        }
                    }
                } else if (p.uall) {
                    // all graphs' U_t in shared memory: ONE pass over the M columns of W (without 2 G block barriers and G tile
                    // loads; 8 L2 loads in flight per thread - one round trip is ~700 cycles, a plain loop pays it per element)
                    for (int e0_ = tid; e0_ < M * NU; e0_ += 8 * GAGM_THREADS) {
                        double uv[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) { const int e = e0_ + q * GAGM_THREADS; uv[q] = e < M * NU ? __ldcg(Ul + e) : 0.0; }
#pragma unroll
                        for (int q = 0; q < 8; ++q) { const int e = e0_ + q * GAGM_THREADS; if (e < M * NU) Uall[e] = uv[q]; }
                    }
                    __syncthreads();
                    rows_times_tile(p.W + (size_t)o * M, M, n, M, Uall, v2, warp, lane);
                } else
                for (int h = 0; h < G; ++h) {
                    const int oh = p.node_off[h], nh = p.node_off[h + 1] - oh;
                    const double *Us;
                    if (h == g) Us = Ug;
                    else {
                        __syncthreads();                           // previous tile fully consumed
                        for (int e = tid; e < nh * NU; e += GAGM_THREADS) Uo[e] = __ldcg(Ul + (size_t)oh * NU + e);
                        __syncthreads();
                        Us = Uo;
                    }
                    rows_times_tile(p.W + (size_t)o * M + oh, M, n, nh, Us, v2, warp, lane);
                }
                HSEG(3);
                // V = (V1 * qw * 2 + V2) / G
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int row = warp + GAGM_WARPS * r;
                    if (row < n && !tc) {
                        const double vs = v1[r] * p.quad_weight * 2.0 + v2[r];
                        const double v = gp2 ? vs * invG : vs / (double)G;     // a power of two: the reciprocal is exact
                        Z[row * ZP + lane] = projector == 0 ? v / tau : v;
                    }
                }
                __syncthreads();
                HSEG(4);
                const long long ck_pj = clock64();
                if (projector == 0) ck_v += ck_pj - ck_ph2;
                // ---- projector -> U_new_g in Ug
                if (projector == 0) {
                    // working matrix = transpose (rows = universe) for n > 32, and for n == 32 inside a ragged batch
                    // whose padded shape is tall (pygmtools transposes the whole batch)
                    const bool tr = n > NU || (n == NU && p.sq_transposed);
                    const int nr = tr ? NU : n, nq = tr ? n : NU;
                    const int ldr = tr ? 1 : ZP, ldq = tr ? ZP : 1;
                    const int mult = nq - nr;                      // dummy_row = True (mgm:333-349)
                    // The reference normalises in the log domain: z -= logsumexp over rows / columns, alternately, sk_iter times, then
                    // exp - one fp64 exp per element per step: the projector was FP64-pipe bound (~100 k cycles per iteration, a third
                    // of the solver on the bench workload).  Here only the first TWO steps (a row and a column normalisation) run in
                    // the log domain; they bring every entry to <= 0 with column sums 1 and row sums >= 1 / 33, so K = exp(z_2) cannot
                    // lose a row or a column to underflow (entries that do underflow are < 1e-300 of their row's and column's mass:
                    // the log form's own final exp returns 0 for them too).  The remaining steps use the SCALING form of the same
                    // iteration: exp(z_k)_rq = a_r K_rq b_q, row step a_r = 1 / sum_q K_rq b_q (dummy rows, all alike: a_d = 1 /
                    // sum_q kd_q b_q), column step b_q = 1 / (sum_r a_r K_rq + mult a_d kd_q) - multiply-adds, one thread per line, no
                    // cross-lane reductions.  Identical in exact arithmetic; ~1e-15 relative apart in fp64.
                    for (int q = tid; q < nq; q += GAGM_THREADS) padv[q] = -100.0;
                    __syncthreads();
                    const int klog = p.sk_iter < 2 ? p.sk_iter : 2;
                    for (int k = 0; k < klog; ++k) {
                        sinkhorn_step(Z, ldr, ldq, nr, nq, padv, mult, k, nullptr, warp, GAGM_WARPS, lane);
                        __syncthreads();
                    }
                    for (int e = tid; e < n * NU; e += GAGM_THREADS) Z[(e / NU) * ZP + (e % NU)] = exp(Z[(e / NU) * ZP + (e % NU)]);   // K (or the result)
                    for (int q = tid; q < nq; q += GAGM_THREADS) { padv[q] = exp(padv[q]); bvec[q] = 1.0; }                          // kd, b
                    if (tid <= nr && tid < NU + 8) avec[tid] = 1.0;
                    __syncthreads();
                    for (int k = klog; k < p.sk_iter; ++k) {
                        if ((k & 1) == 0) {
                            if (tid < nr) {
                                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                                const double *zr = Z + tid * ldr;
                                int q = 0;
                                for (; q + 3 < nq; q += 4) {
                                    s0 = fma(zr[q * ldq], bvec[q], s0); s1 = fma(zr[(q + 1) * ldq], bvec[q + 1], s1);
                                    s2 = fma(zr[(q + 2) * ldq], bvec[q + 2], s2); s3 = fma(zr[(q + 3) * ldq], bvec[q + 3], s3);
                                }
                                for (; q < nq; ++q) s0 = fma(zr[q * ldq], bvec[q], s0);
                                avec[tid] = 1.0 / ((s0 + s1) + (s2 + s3));
                            } else if (tid == 32 && mult > 0) {
                                double s0 = 0.0;
                                for (int q = 0; q < nq; ++q) s0 = fma(padv[q], bvec[q], s0);
                                avec[nr] = 1.0 / s0;
                            }
                        } else if (tid < nq) {
                            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                            const double *zq = Z + tid * ldq;
                            int r = 0;
                            for (; r + 3 < nr; r += 4) {
                                s0 = fma(avec[r], zq[r * ldr], s0); s1 = fma(avec[r + 1], zq[(r + 1) * ldr], s1);
                                s2 = fma(avec[r + 2], zq[(r + 2) * ldr], s2); s3 = fma(avec[r + 3], zq[(r + 3) * ldr], s3);
                            }
                            for (; r < nr; ++r) s0 = fma(avec[r], zq[r * ldr], s0);
                            bvec[tid] = 1.0 / (((s0 + s1) + (s2 + s3)) + (mult > 0 ? (double)mult * avec[nr] * padv[tid] : 0.0));
                        }
                        __syncthreads();
                    }
                    // U_new (node-major n x NU) = a K b on the real rows
                    for (int e = tid; e < n * NU; e += GAGM_THREADS) {
                        const int node = e / NU, slot = e % NU;
                        const int r = tr ? slot : node, q = tr ? node : slot;
                        Ug[e] = avec[r] * Z[node * ZP + slot] * bvec[q];
                    }
                    ck_proj += clock64() - ck_pj;
                } else {
                    // SciPy refuses a cost matrix with NaN / inf ("matrix contains invalid numeric entries", utils/hungarian.py:34 ->
                    // linear_sum_assignment); the shortest-path search would not terminate on one.  Non-finite entries are zeroed so
                    // that the solve ends, and info[15] tells the host to raise the same error.
                    int bad = 0;
                    for (int e = tid; e < n * NU; e += GAGM_THREADS) {
                        Ug[e] = 0.0;
                        double &zv = Z[(e / NU) * ZP + (e % NU)];
                        if (!(fabs(zv) <= 1.79e308)) { zv = 0.0; bad = 1; }
                    }
                    if (__syncthreads_or(bad) && tid == 0 && p.info) p.info[15] = 1;
                    if (warp == 0) {
                        const long long ck_l = clock64();
                        const double *Zc = Z;
                        const int lf = p.lap_fast >= 3 ? 0 : p.lap_fast;
                        const bool lean = p.lap_fast >= 3, bfm = p.lap_fast == 4;
                        if (n <= NU) {
                            if (!(lean && lap_lean_warp(n, NU, LapSmemNegCostP{Zc, ZP, 1}, *lapw, bfm)))
                                lap_solve_warp(n, NU, LapSmemNegCost{lap_smem_u32(Zc), ZP, 1}, *lapw, lf);
                            for (int i = lane; i < n; i += 32) Ug[i * NU + lapw->col4row[i]] = 1.0;
                        } else {
                            if (!(lean && lap_lean_warp(NU, n, LapSmemNegCostP{Zc, 1, ZP}, *lapw, bfm)))
                                lap_solve_warp(NU, n, LapSmemNegCost{lap_smem_u32(Zc), 1, ZP}, *lapw, lf);
                            for (int i = lane; i < NU; i += 32) Ug[lapw->col4row[i] * NU + i] = 1.0;
                        }
                        ck_lap += clock64() - ck_l;
                    }
                }
                __syncthreads();
                HSEG(5);
                if (G == 2 && g == 0) {                            // mgm:358-359
                    for (int e = tid; e < n * NU; e += GAGM_THREADS) Ug[e] = (e / NU == e % NU) ? 1.0 : 0.0;
                    __syncthreads();
                }
                const bool lean = hf && projector == 1 && it_hg >= 2;      // U_t and U_{t-1} are permutations too: norms by counting
                last_lean = lean;
                if (projector == 1 && tid < NU) {                  // node_of of U_{t+1} for the next (Hungarian-stage) iteration
                    int nd = -1;
                    for (int r = 0; r < n; ++r) if (Ug[r * NU + tid] != 0.0) nd = r;
                    if (hf) {
                        int *dst = nodeof_s + ((it_total + 1) & 1) * (GAGM_MAX_C * NU) + g * NU + tid;
                        for (int cc = 0; cc < C; ++cc) *cluster.map_shared_rank(dst, cc) = nd;
                        if (lean) {                                // entries are 0 / 1: ||U' - U||^2 = slots whose node changed, per side
                            d1 += nd != nd_cur ? (double)((nd >= 0) + (nd_cur >= 0)) : 0.0;
                            d2 += nd != nd_prev ? (double)((nd >= 0) + (nd_prev >= 0)) : 0.0;
                        }
                        nd_prev = nd_cur; nd_cur = nd;
                    } else nodeof_nxt[g * NU + tid] = nd;
                }
                if (lean) {
                    if (p.trace && it_total < p.trace_cap)
                        for (int e = tid; e < n * NU; e += GAGM_THREADS) p.trace[(size_t)(it_total + 1) * UB + (size_t)o * NU + e] = Ug[e];
                } else
                for (int e = tid; e < n * NU; e += GAGM_THREADS) {
                    const double un = Ug[e];
                    const double a = un - __ldcg(Ul + (size_t)o * NU + e);
                    const double b = un - __ldcg(Ul2 + (size_t)o * NU + e);
                    d1 = fma(a, a, d1);
                    d2 = fma(b, b, d2);
                    Un[(size_t)o * NU + e] = un;
                    if (p.trace && it_total < p.trace_cap) {
                        if (it_total == 0) p.trace[(size_t)o * NU + e] = __ldcg(Ul + (size_t)o * NU + e);
                        p.trace[(size_t)(it_total + 1) * UB + (size_t)o * NU + e] = un;
                    }
                }
                __syncthreads();
                HSEG(6);
            }
            d1 = warp_sum(d1); d2 = warp_sum(d2);
            if (lane == 0) { red[warp] = d1; red[GAGM_WARPS + warp] = d2; }
            __syncthreads();
            double *nx = normx + (it_total & 1) * (2 * GAGM_MAX_C);
            if (tid == 0) {
                double a = 0.0, b = 0.0;
                for (int w = 0; w < GAGM_WARPS; ++w) { a += red[w]; b += red[GAGM_WARPS + w]; }
                if (hf) {
                    for (int cc = 0; cc < C; ++cc) { double *r = cluster.map_shared_rank(nx, cc); r[2 * c] = a; r[2 * c + 1] = b; }
                } else { p.normpart[2 * c] = a; p.normpart[2 * c + 1] = b; }
            }
            if (!last_lean) __threadfence();                       // U_{t+1} / norm partials in global memory
            HSEG(7);
            { const long long ck_b = clock64(); cluster_sync_all(); ck_bar += clock64() - ck_b; }
            HSEG(8);
            double n1 = 0.0, n2 = 0.0;
            if (hf) {
#pragma unroll
                for (int cc = 0; cc < GAGM_MAX_C; ++cc) if (cc < C) { n1 += nx[2 * cc]; n2 += nx[2 * cc + 1]; }
            } else
            for (int cc = 0; cc < C; ++cc) { n1 += __ldcg(p.normpart + 2 * cc); n2 += __ldcg(p.normpart + 2 * cc + 1); }
            if (p.trace_meta && c == 0 && tid == 0 && it_total < p.trace_cap) {
                p.trace_meta[2 * it_total] = (double)projector; p.trace_meta[2 * it_total + 1] = tau;
            }
            ++it_total;
            if (projector == 1) ck_hung += clock64() - ck_it;
            binU = (projector == 1);
            if (projector == 0) ++it_sk; else { ++it_hg; n_lap += G; }
            if (p.mode == 1) { stop_all = true; break; }
            HSEG(9);
            if (sqrt(n1) < p.tol || n2 == 0.0) break;              // mgm:361
        }
        if (stop_all) break;
        // projection control (mgm:373-383); "not converged" with hung_iter is a no-op (mgm:364-366)
        if (projector == 1) break;
        ++n_stage;
        if (tau > p.min_tau) tau *= p.sk_gamma;
        else projector = 1;
    }

    const double *Uf = p.Ubuf + (size_t)cur * UB;
    if (last_lean) {                                           // the last iterations kept U only as node_of
        const int o = p.node_off[c], n = p.node_off[c + 1] - o;
        for (int e = tid; e < n * NU; e += GAGM_THREADS) p.U_out[(size_t)o * NU + e] = 0.f;
        __syncthreads();
        if (tid < NU && nd_cur >= 0) p.U_out[(size_t)(o + nd_cur) * NU + tid] = 1.f;
    } else
    for (int g = c; g < G; g += C) {
        const int o = p.node_off[g], n = p.node_off[g + 1] - o;
        for (int e = tid; e < n * NU; e += GAGM_THREADS) p.U_out[(size_t)o * NU + e] = (float)Uf[(size_t)o * NU + e];
    }
    if (c == 0 && tid == 0 && p.info) {
        p.info[0] = it_total; p.info[1] = it_sk; p.info[2] = it_hg; p.info[3] = n_lap; p.info[4] = n_stage;
        p.info[5] = lapw->stat_steps; p.info[6] = lapw->stat_hops; p.info[7] = lapw->stat_fast_fallback;     // graph 0's LAPs: Dijkstra steps, path hops
        p.info[8] = (int)((clock64() - ck_start) >> 10); p.info[9] = (int)(ck_hung >> 10); p.info[10] = (int)(ck_lap >> 10);
        p.info[11] = (int)(ck_bar >> 10);
        p.info[12] = (int)(ck_p1 >> 10); p.info[13] = (int)(ck_v >> 10); p.info[14] = (int)(ck_proj >> 10);
        for (int i = 0; i < 10; ++i) g_gagm_prof[i] = hs[i];
        for (int i = 0; i < 6; ++i) g_gagm_prof[10 + i] = lapw->stat_ck[i];
        for (int i = 0; i < 8; ++i) g_gagm_prof[16 + i] = lapw->stat_free[i];
    }
}

static size_t gagm_smem_bytes() { return GAGM_FIXED_SMEM; }

}  // namespace ttdg

using namespace ttdg;

static int g_lap_fast = -1;     // -1: read TTDG_LAP_FAST at first use

// Hungarian projections inside the GA-GM solver (lap.cuh): 0 = SciPy-order solve only, 1 / 2 = row-reduction / auction start +
// uniqueness certificate, 3 (default) = lean certified solve, 4 = its label-correcting variant; the SciPy-order solve is the
// fall-back.  Results are identical by construction; info[7] counts the fall-backs of graph 0.  Returns the previous setting.
extern "C" int ttdg_gagm_set_lap_fast(int on) {
    const int prev = g_lap_fast < 0 ? 3 : g_lap_fast;
    g_lap_fast = (on >= 1 && on <= 4) ? on : 0;              // 2: Jacobi-auction start instead of the row reduction; 3: lean certified solve
    return prev;
}

// Diagnostic: cycles CTA 0 spent in the segments of the Hungarian-stage iterations of the last solve (synchronises the device):
// {T build, Q gather, V1 = A Q, V2 = W U, V store, projection, U store + norms, norm reduce, cluster barrier, tail} and, inside the
// lean LAPs of graph 0, {auction scans, bid resolution, augmentations, certificate, its reachability part, its Kahn part} and
// [16..23] the free rows at the start of each auction round / after the last, summed over graph 0's LAPs.
extern "C" int ttdg_gagm_read_profile(int64_t *out24) {
    long long h[24];
    cudaError_t e = cudaMemcpyFromSymbol(h, g_gagm_prof, sizeof(h));
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < 24; ++i) out24[i] = (int64_t)h[i];
    return 0;
}

extern "C" int64_t ttdg_gagm_scratch_bytes(int M, int G) {
    (void)G;
    return (int64_t)(3 * (int64_t)M * NU + GAGM_MAX_C * NU * NU + 2 * GAGM_MAX_C) * (int64_t)sizeof(double) +
           (int64_t)(2 * GAGM_MAX_G * NU) * (int64_t)sizeof(int32_t);
}

extern "C" int ttdg_gagm_solve(const float *A, const float *W, const float *U0, const int32_t *ms_h, int G, int M,
                               int n_univ, double init_tau, double min_tau, double sk_gamma, int max_iter, int sk_iter,
                               double converge_tol, double quad_weight, int mode, int step_projector, float *U,
                               int32_t *info, void *scratch, double *trace, double *trace_meta, int trace_cap,
                               void *stream) {
    TTDG_CHECK_ARG(A && W && U0 && ms_h && U && scratch && G >= 1 && M >= 1 && max_iter >= 1 && sk_iter >= 0);
    TTDG_CHECK_ARG(init_tau > 0 && (mode == 0 || mode == 1) && (step_projector == 0 || step_projector == 1));
    if (n_univ != NU || G > GAGM_MAX_G) return TTDG_E_LIMIT;
    GagmParams p;
    p.A = A; p.W = W; p.U0 = U0; p.U_out = U; p.info = info;
    p.G = G; p.M = M; p.C = G < GAGM_MAX_C ? G : GAGM_MAX_C;
    p.node_off[0] = 0;
    for (int g = 0; g < G; ++g) {
        if (ms_h[g] < 1 || ms_h[g] > GAGM_MAX_N) return TTDG_E_LIMIT;
        p.node_off[g + 1] = p.node_off[g] + ms_h[g];
    }
    if (p.node_off[G] != M) return TTDG_E_ARG;
    {   // mgm:330-353: equal sizes -> plain reshape (never square-transposed); ragged -> padded batch G x max_n x 32
        bool equal = true; int mx = 0;
        for (int g = 0; g < G; ++g) { equal = equal && ms_h[g] == ms_h[0]; mx = ms_h[g] > mx ? ms_h[g] : mx; }
        p.sq_transposed = (!equal && mx > NU) ? 1 : 0;
    }
    p.Ubuf = reinterpret_cast<double *>(scratch);
    p.Tpart = p.Ubuf + 3 * (size_t)M * NU;
    p.normpart = p.Tpart + GAGM_MAX_C * NU * NU;
    p.nodeof = reinterpret_cast<int32_t *>(p.normpart + 2 * GAGM_MAX_C);
    p.init_tau = init_tau; p.min_tau = min_tau; p.sk_gamma = sk_gamma; p.tol = converge_tol; p.quad_weight = quad_weight;
    p.max_iter = max_iter; p.sk_iter = sk_iter; p.mode = mode; p.step_projector = step_projector;
    p.trace = trace; p.trace_meta = trace_meta; p.trace_cap = trace ? trace_cap : 0;
    if (g_lap_fast < 0) { const char *e = getenv("TTDG_LAP_FAST"); g_lap_fast = (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 3; }
    p.lap_fast = g_lap_fast;

    size_t smem = gagm_smem_bytes();
    p.uall = 0;
    static int s_uall = -1;
    if (s_uall < 0) { const char *ev = getenv("TTDG_GAGM_UALL"); s_uall = (ev && ev[0] >= '0' && ev[0] <= '2') ? ev[0] - '0' : 2; }
    // 2: pitch GAGM_UP + FP64 tensor-core product (TTDG_GAGM_UALL=1 keeps the vector form), 1: pitch NU, vector form
    if (s_uall >= 2 && smem + (size_t)M * GAGM_UP * sizeof(double) <= 227 * 1024) { p.uall = 2; smem += (size_t)M * GAGM_UP * sizeof(double); }
    else if (s_uall && smem + (size_t)M * NU * sizeof(double) <= 227 * 1024) { p.uall = 1; smem += (size_t)M * NU * sizeof(double); }
    // Hungarian-stage cache (reuses the U-all region): diagonal blocks of A for all graphs + one graph's rows of W, fp32
    static int s_hfast = -1;
    if (s_hfast < 0) { const char *ev = getenv("TTDG_GAGM_HFAST"); s_hfast = (ev && ev[0] >= '0' && ev[0] <= '2') ? ev[0] - '0' : 1; }   // 2: without the fp32 cache
    p.hfast = 0; p.hcache = 0;
    for (int h = 0; h <= GAGM_MAX_C; ++h) p.aoff[h] = 0;
    if (s_hfast && G <= GAGM_MAX_C) {
        p.hfast = 1;
        int mx = 0;
        for (int g = 0; g < G; ++g) { p.aoff[g + 1] = p.aoff[g] + ms_h[g] * ms_h[g]; mx = ms_h[g] > mx ? ms_h[g] : mx; }
        const size_t need = ((size_t)((p.aoff[G] + 3) & ~3) + (size_t)mx * M) * sizeof(float);
        const size_t base = gagm_smem_bytes();
        if (s_hfast != 2 && base + need <= 227 * 1024) {
            p.hcache = 1;
            if (base + need > smem) smem = base + need;
        }
    }
    cudaError_t e = cudaFuncSetAttribute(gagm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.C);
    cfg.blockDim = dim3(GAGM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ttdg::count_launches(1);
    e = cudaLaunchKernelEx(&cfg, gagm_kernel, p);
    return (int)e;
}
