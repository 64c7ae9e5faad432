/* ttdg_b200.h - C ABI of libttdg_sm100.so, the B200 (sm_100a) device library behind the TTDG-MGM
 * test-time-adaptation hot path.
 *
 * Conventions (last row):
 *   - every pointer is a DEVICE pointer unless its name ends in _h; tensors are dense row-major;
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns at once;
 *   - return value: 0 = ok, >0 = cudaError_t of the launch, <0 = argument error (TTDG_E_*);
 *   - nothing is allocated and no pointer is retained: callers pass scratch buffers whose sizes come
 *     from the matching *_scratch_bytes() function (host-only, no GPU needed);
 *   - "ragged" batches are described by int32 offset arrays of length count+1 (prefix sums).
 *
 * Each entry point cites the reference interface it replaces (paths relative to
 * /root/reference/adapteacher/modeling/GModule unless stated).
 *
 * Arithmetic contract of the matching stage ("precise mode"): fp32 in / fp32 out,
 * fp64 internally with a single rounding at each documented point, so results do not depend on
 * summation order and the discrete solver (LAP inside GA-GM) is reproducible bit for bit.
 */
#ifndef TTDG_B200_H
#define TTDG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTDG_E_ARG (-1)       /* bad argument (null pointer, negative size, ...)        */
#define TTDG_E_LIMIT (-2)     /* size above a compiled-in limit (see ttdg_limits)        */

/* library / build info; never touches the GPU */
int ttdg_version(void);                         /* MAJOR*10000 + MINOR*100 + PATCH */
const char *ttdg_build_info(void);              /* "sm_100a nvcc 12.9 ..." */
int ttdg_limit(const char *name);               /* "lap_max_dim", "sinkhorn_small_max_dim", "gagm_max_graphs", ... ; -1 if unknown */

/* ---------------------------------------------------------------------------------------------
 * Sinkhorn.  Replaces utils/sinkhorn.py:58-87 -> pygmtools.sinkhorn(backend='pytorch') (0.3.8).
 * Per-item semantics: the item is an n1 x n2 matrix stored with leading dimension
 * `ld` at `s + item_off[b]` (element offsets, int64).  If transpose[b] != 0 the stored matrix is read as its
 * transpose (so the working matrix always has rows <= cols); working matrix / tau; if dummy_row the
 * (cols-rows) missing rows are filled with -100; `max_iter` alternating normalisations (even = over the
 * columns of each row, odd = over the rows of each column); exp; written back in the stored layout.
 * --------------------------------------------------------------------------------------------- */

/* small matrices (rows, cols <= ttdg_limit("sinkhorn_small_max_dim")): one CTA per item, matrix resident
 * in shared memory, fp64 internal.  dims[b] = {n1, n2, ld, transpose} as stored. */
int ttdg_sinkhorn_small_fwd(const float *s, float *out, const int64_t *item_off, const int32_t *dims4,
                            int n_items, double tau, int max_iter, int dummy_row, void *stream);
/* backward of the above: grad_in = d(sum(out * grad_out)) / d s.  Recomputes the forward in-kernel. */
int ttdg_sinkhorn_small_bwd(const float *s, const float *grad_out, float *grad_in, const int64_t *item_off,
                            const int32_t *dims4, int n_items, double tau, int max_iter, int dummy_row,
                            void *stream);

/* large matrices: HBM-streaming fp32 path (the N = 256/512/1024 microbenchmark of BASELINE.json).
 * Uniform batch of `batch` dense n1 x n2 matrices, n1 <= n2, no dummy rows needed (n1 == n2) or
 * dummy_row with n1 < n2.  scratch: ttdg_sinkhorn_stream_scratch_bytes(batch, n1, n2).
 * One launch of a persistent kernel: each CTA owns whole matrices and runs all iterations on them. */
int64_t ttdg_sinkhorn_stream_scratch_bytes(int batch, int n1, int n2);
int ttdg_sinkhorn_stream_fwd(const float *s, float *out, int batch, int n1, int n2, float tau, int max_iter,
                             int dummy_row, void *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------
 * LAP.  Replaces utils/hungarian.py:8-65 -> scipy.optimize.linear_sum_assignment(-s) (fp64, Crouse
 * shortest augmenting path, SciPy tie-breaking).  perm (same layout as s) receives
 * 0/1 float32.  One warp per item; rows, cols <= ttdg_limit("lap_max_dim").  dims3[b] = {n1, n2, ld}.
 * --------------------------------------------------------------------------------------------- */
int ttdg_lap_solve(const float *s, float *perm, const int64_t *item_off, const int32_t *dims3, int n_items,
                   void *stream);

/* ---------------------------------------------------------------------------------------------
 * Attention adjacency.  Replaces MGM3_unsup._forward_intra_graph (multi_graph_matching.py:571-574) ->
 * MultiHeadAttention.forward v2 (utils/attentions.py:60-86), keeping only the attention map that
 * mgm:498-502 uses:  A[blk g] = dropout(softmax((x Wq^T + bq)(x Wk^T + bk)^T / 16)), diagonal zeroed,
 * everything outside the diagonal blocks zeroed.  A is M x M (M = node_off[G]).
 * dropout: keep_mask != NULL -> explicit M-row ragged masks (mask_off[g] element offsets, n_g x n_g 0/1
 * floats); else if p_drop > 0 -> Philox4x32-10 keyed by (seed, offset); p_drop == 0 -> eval mode.
 * scratch: ttdg_attn_scratch_bytes(M).
 * --------------------------------------------------------------------------------------------- */
int64_t ttdg_attn_scratch_bytes(int M);
int ttdg_attn_adjacency(const float *nodes, const int32_t *node_off, int G, const float *wq, const float *bq,
                        const float *wk, const float *bk, const float *keep_mask, const int64_t *mask_off,
                        float p_drop, uint64_t seed, uint64_t offset, float *A, void *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Learned affinity.  Replaces MGM3_unsup._forward_aff (mgm:576-582) -> Affinity.forward
 * (utils/affinity.py:44-57) in separable form (no N1 x N2 x 512 tensor):
 *     a = (X Ps^T) W0a^T,  c = (Y Pt^T) W0b^T + b0,  M_ij = sum_k w1_k relu(a_ik + c_jk) + b1
 * for every listed (src, tgt) graph pair.  nodes: M x 256.  pairs: int32 {src, tgt} x n_pairs.
 * out_off[p]: element offset of pair p's n_src x n_tgt block in `out`.
 * scratch (fp64 a, c, projections): ttdg_affinity_scratch_bytes(M).
 * --------------------------------------------------------------------------------------------- */
int64_t ttdg_affinity_scratch_bytes(int M);
int ttdg_affinity_fwd(const float *nodes, const int32_t *node_off, int G, const float *w_sr, const float *w_tg,
                      const float *w0, const float *b0, const float *w1, const float *b1, const int32_t *pairs,
                      const int64_t *out_off, int n_pairs, float *out, void *scratch, void *stream);
/* backward: given grad_out (same ragged layout as out) accumulates into grad_nodes (M x 256) and the six
 * parameter gradients (all fp32, must be zero-initialised or hold a running sum). `scratch` must be the
 * buffer the forward filled (a, c are reused). scratch2: ttdg_affinity_bwd_scratch_bytes(M). */
int64_t ttdg_affinity_bwd_scratch_bytes(int M);
int ttdg_affinity_bwd(const float *nodes, const int32_t *node_off, int G, const float *w_sr, const float *w_tg,
                      const float *w0, const float *w1, const int32_t *pairs, const int64_t *out_off, int n_pairs,
                      const float *grad_out, float *grad_nodes, float *g_w_sr, float *g_w_tg, float *g_w0,
                      float *g_b0, float *g_w1, float *g_b1, void *scratch, void *scratch2, void *stream);

/* U0 = nodes @ universe^T  (mgm:531-532): nodes M x 256, universe n_univ x 256 -> M x n_univ. */
int ttdg_universe_init(const float *nodes, int M, const float *universe, int n_univ, int dim, float *U0,
                       void *stream);

/* ---------------------------------------------------------------------------------------------
 * GA-GM solver.  Replaces GA_GM.forward + gagm (multi_graph_matching.py:223-244, 300-389) for the
 * configuration the hot path uses (num_clusters = 1, projector0 = 'sinkhorn', hung_iter = True) including
 * the per-iteration projector: batched Sinkhorn (mgm:330-353) or per-graph Hungarian (mgm:324-328), the
 * G == 2 identity quirk (mgm:358-359) and both convergence tests (mgm:361) - all on the device, no
 * host round trip.  One persistent CTA.  A, W: M x M; U0, U: M x 32; ms: int32[G].
 * mode: 0 = full solve; 1 = exactly one iteration with projector `step_projector` (0 sinkhorn, 1 hungarian)
 * at temperature init_tau (teacher-forced parity tests).
 * info (int32[8], device): {iterations, sinkhorn-stage iterations, hungarian-stage iterations, LAP calls,
 *                           sinkhorn stages, converged-flag of last stage, 0, 0}.
 * scratch: ttdg_gagm_scratch_bytes(M, G).  Scalars are doubles because the reference keeps tau etc. as
 * Python floats (mgm:307,379).
 * --------------------------------------------------------------------------------------------- */
int64_t ttdg_gagm_scratch_bytes(int M, int G);
int ttdg_gagm_solve(const float *A, const float *W, const float *U0, const int32_t *ms, int G, int M,
                    double init_tau, double min_tau, double sk_gamma, int max_iter, int sk_iter,
                    double converge_tol, double quad_weight, int mode, int step_projector, float *U,
                    int32_t *info, void *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Matching loss.  Replaces collect_intra_class_matching_wrapper + the 'perm' loss loop (mgm:543-564,
 * 594-633) -> PermutationLoss / BCEFocalLoss (utils/losses.py:83-103, 419-455):
 *     loss = mean over pairs i1<i2 of mean_ij focal_bce(clamp(S_ij), (U_i1 U_i2^T)_ij)
 * S is read from Wds (M x M) with the orientation rule of mgm:620-623.  loss: 1 float (device).
 * The backward writes grad_Wds (M x M, zero outside the touched blocks) scaled by *grad_loss.
 * --------------------------------------------------------------------------------------------- */
int ttdg_matching_loss_fwd(const float *Wds, const float *U, const int32_t *node_off, int G, int M, int n_univ,
                           float *loss, int32_t *flags, void *stream);
int ttdg_matching_loss_bwd(const float *Wds, const float *U, const int32_t *node_off, int G, int M, int n_univ,
                           const float *grad_loss, float *grad_Wds, void *stream);
/* generic focal BCE on one matrix (utils/losses.py:83-103): mean reduction. */
int ttdg_focal_bce_fwd(const float *p, const float *y, int64_t n, float *loss, void *stream);
int ttdg_focal_bce_bwd(const float *p, const float *y, int64_t n, const float *grad_loss, float *grad_p,
                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TTDG_B200_H */
