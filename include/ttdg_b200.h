/* ttdg_b200.h - C ABI of libttdg_sm100.so, the B200 (sm_100a) device library behind the TTDG-MGM
 * test-time-adaptation hot path.
 *
 * Conventions (last row):
 *   - every pointer is a DEVICE pointer unless its name ends in _h; tensors are dense row-major fp32
 *     unless stated; "scratch" buffers are opaque device memory sized by the matching *_scratch_bytes();
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns at once;
 *   - return value: 0 = ok, >0 = cudaError_t of the launch, <0 = argument error (TTDG_E_*);
 *   - nothing is allocated and no pointer is retained;
 *   - ragged batches are described by small int32/int64 descriptor arrays in DEVICE memory.
 *
 * Each entry point cites the reference interface it replaces (paths relative to
 * /root/reference/adapteacher/modeling/GModule unless stated).
 *
 * Arithmetic contract of the matching stage ("precise mode"): fp32 in / fp32 out,
 * fp64 internally, so results do not depend on summation order and the discrete solver (LAP inside
 * GA-GM) is reproducible bit for bit.  The HBM-streaming Sinkhorn (large matrices) is fp32.
 */
#ifndef TTDG_B200_H
#define TTDG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTDG_E_ARG (-1)   /* bad argument (null pointer, negative size, ...)  */
#define TTDG_E_LIMIT (-2) /* size above a compiled-in limit (see ttdg_limit)  */

/* library / build info; never touches the GPU */
int ttdg_version(void);            /* MAJOR*10000 + MINOR*100 + PATCH */
const char *ttdg_build_info(void); /* "sm_100a nvcc 12.9 ..." */
int ttdg_limit(const char *name);
long long ttdg_launch_count(void); /* kernels launched through this library since load (all entry points) */
  /* "small_max_dim", "lap_max_dim", "gagm_max_graphs", "univ", "feat_dim"; -1 if unknown */

/* ---------------------------------------------------------------------------------------------
 * Sinkhorn.  Replaces utils/sinkhorn.py:58-87 -> pygmtools.sinkhorn(backend='pytorch') (0.3.8).
 * Per-item semantics: the stored n1 x n2 matrix is worked on in the orientation
 * with rows <= cols (transposed when n2 < n1; a SQUARE item is transposed iff flag bit 0 is set - pygmtools
 * transposes a whole padded batch when its padded shape is tall, and only re-transposes the items that are
 * strictly wide, so square items inherit the batch orientation); divided by tau; if dummy_row the (cols-rows) missing rows
 * are filled with -100; `max_iter` alternating normalisations (even = each row over its columns, odd =
 * each column over its rows); exp; written back in the stored orientation.
 * --------------------------------------------------------------------------------------------- */

/* Small matrices (n1, n2 <= ttdg_limit("small_max_dim")): one CTA per item, matrix resident in shared
 * memory, fp64 internal.  items: int64[n_items][9] =
 *   { s_off, out_off, mirror_off, n1, n2, ld_s, ld_out, ld_mirror, flags }   (element offsets / leading dims)
 * out receives the n1 x n2 result at out_off; if mirror_off >= 0 its TRANSPOSE (n2 x n1) is also written
 * at mirror_off (MGM3_unsup stores both Wds[src,tgt] and Wds[tgt,src], mgm:523-525).
 * max_dim: host-known upper bound of every n1, n2 (sizes the shared memory; items above it are skipped). */
int ttdg_sinkhorn_small_fwd(const float *s, float *out, const int64_t *items, int n_items, int max_dim,
                            double tau, int max_iter, int dummy_row, void *stream);
/* Backward: grad_in = d(sum(out * grad_out)) / d s.  Recomputes the forward in-kernel.  items: int64[n][9] =
 *   { s_off, gout_off, gin_off, n1, n2, ld_s, ld_gout, ld_gin, flags }. */
int ttdg_sinkhorn_small_bwd(const float *s, const float *grad_out, float *grad_in, const int64_t *items,
                            int n_items, int max_dim, double tau, int max_iter, int dummy_row, void *stream);

/* Large matrices: fp32 path for the N = 256/512/1024 micro-benchmark of BASELINE.json.  Uniform batch of
 * dense n1 x n2 matrices (n1 <= n2; n1 < n2 needs dummy_row = 0 or 1 as in the reference; n2 % 4 == 0).
 * One thread-block cluster (1..8 CTAs) per matrix keeps the matrix in distributed shared memory for all
 * iterations (row/column potentials: the matrix is never rewritten, no intermediate HBM traffic); rows that do
 * not fit (N = 1024) are re-read through L2.  scratch: ttdg_sinkhorn_stream_scratch_bytes. */
int64_t ttdg_sinkhorn_stream_scratch_bytes(int batch, int n1, int n2);
int ttdg_sinkhorn_stream_fwd(const float *s, float *out, int batch, int n1, int n2, float tau, int max_iter,
                             int dummy_row, void *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------
 * LAP.  Replaces utils/hungarian.py:8-65 -> scipy.optimize.linear_sum_assignment(-s) (fp64, Crouse
 * shortest augmenting path, SciPy tie-breaking).  One warp per item;
 * n1, n2 <= ttdg_limit("lap_max_dim").  items: int64[n][6] = { s_off, perm_off, n1, n2, ld_s, ld_perm };
 * perm receives a 0/1 float32 matrix of the max-weight assignment.
 * --------------------------------------------------------------------------------------------- */
int ttdg_lap_solve(const float *s, float *perm, const int64_t *items, int n_items, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Y = X W^T (+ b): the small dense layers of the matching head (attention q/k projections
 * attentions.py:66-69, affinity projections affinity.py:48-49, universe init mgm:531-532).
 * X: m x k (ldx), W: n x k (ldw), Y: m x n (ldy), fp32 storage, fp64 accumulation, one rounding.
 * --------------------------------------------------------------------------------------------- */
int ttdg_linear_f64acc(const float *X, int ldx, const float *W, int ldw, const float *b, float *Y, int ldy,
                       int m, int n, int k, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Attention adjacency.  Replaces MGM3_unsup._forward_intra_graph (mgm:571-574) ->
 * MultiHeadAttention.forward v2 (utils/attentions.py:60-86), keeping only the attention map that
 * mgm:498-502 uses:  A[blk g] = dropout(softmax(q_g k_g^T * scale)), diagonal zeroed, zero outside the
 * diagonal blocks.  S = q k^T (M x M fp32 logits; q, k from ttdg_linear_f64acc, product from
 * ttdg_gemm_f64acc) - only its diagonal blocks are read.  scale = (256 // heads) ** -0.5 = 1/16.
 * A: M x M, fully written.
 * dropout: keep_mask != NULL -> explicit ragged masks (n_g x n_g 0/1 floats at mask_off[g]);
 * else if p_drop > 0 -> Philox4x32-10 keyed by (seed, offset + element index); p_drop == 0 -> eval mode.
 * --------------------------------------------------------------------------------------------- */
int ttdg_attn_adjacency(const float *S, const int32_t *node_off, int G, int M, float scale,
                        const float *keep_mask, const int64_t *mask_off, float p_drop, uint64_t seed,
                        uint64_t offset, float *A, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Learned affinity.  Replaces MGM3_unsup._forward_aff (mgm:576-582) -> Affinity.forward
 * (utils/affinity.py:44-57) in separable form (no N1 x N2 x 512 tensor):
 *     a = Xp W0a^T,  c = Yp W0b^T + b0,  M_ij = sum_k w1_k relu(a_ik + c_jk) + b1
 * with Xp = X Ps^T, Yp = Y Pt^T (ttdg_linear_f64acc), W0 = [W0a | W0b] (hidden x 2 dim).
 * ac: M x (2 hidden) fp64 = [a | c] per node, produced by ttdg_affinity_hidden.
 * pairs: int64[n_pairs][4] = { src_row0, n_src, tgt_row0, n_tgt }; out_off[p] = element offset of pair p's
 * n_src x n_tgt block in `out` (blocks stored back to back).  max_n_src: host-known max of n_src.
 * --------------------------------------------------------------------------------------------- */
int ttdg_affinity_hidden(const float *Xp, const float *Yp, const float *w0, const float *b0, int M, int dim,
                         int hidden, double *ac, void *stream);
int ttdg_affinity_pairs_fwd(const double *ac, const float *w1, const float *b1, const int64_t *pairs,
                            const int64_t *out_off, int n_pairs, int max_n_src, int hidden, float *out,
                            void *stream);
/* backward of the pair stage: g_ac (M x 2 hidden fp64, fully written), g_w1 (hidden) and g_b1 (1) (fp32,
 * overwritten).  Deterministic (no atomics): each node row gathers from the pairs it belongs to.
 * grad_out has the layout of `out` (grad_out_total elements); max_n: host-known max graph size.
 * scratch: ttdg_affinity_bwd_scratch_bytes(M, hidden). */
int64_t ttdg_affinity_bwd_scratch_bytes(int M, int hidden);
int ttdg_affinity_pairs_bwd(const double *ac, const float *w1, const int64_t *pairs, const int64_t *out_off,
                            int n_pairs, int hidden, int M, int max_n, const float *grad_out,
                            int64_t grad_out_total, double *g_ac, float *g_w1, float *g_b1, void *scratch,
                            void *stream);
/* dense helper (fp64 accumulation): C (m x n) = op(A) (m x k) op(B) (k x n) [+ C], row-major operands given
 * as fp32 or fp64 (a_is_f64 ...), op = transpose when trans* != 0.  Used for the attention logits and the
 * affinity autograd: g_w0 = g_ac^T [Xp | Yp], g_Xp = g_a W0a, g_Ps = g_Xp^T X, g_nodes = g_Xp Ps + g_Yp Pt. */
int ttdg_gemm_f64acc(int transA, int transB, int m, int n, int k, const void *A, int a_is_f64, int lda,
                     const void *B, int b_is_f64, int ldb, void *C, int c_is_f64, int ldc, int accumulate,
                     void *stream);

/* Hungarian projections inside ttdg_gagm_solve (also env TTDG_LAP_FAST): 0 = every projection walks SciPy's
 * shortest-augmenting-path order; 1 = start from a row reduction (2 = from a Jacobi auction with epsilon 0), certify that the
 * optimum is unique by a margin (no tight edge to a free column, tight-edge digraph acyclic) and fall back to the SciPy-order
 * solve otherwise; 3 (default) = the lean certified solve (auction rounds + Dijkstra without SciPy's bookkeeping + the same
 * certificate and fall-back), 4 = the same with label-correcting rounds instead of Dijkstra.  Same results by construction
 * (mgm:324-328 -> utils/hungarian.py:58-65), checked on every iteration of the trajectory.  Returns the previous setting. */
int ttdg_gagm_set_lap_fast(int on);
/* Diagnostic (tools/run_kernels.py gagm_*): cycles CTA 0 spent in the segments of the Hungarian-stage iterations of the last
 * ttdg_gagm_solve - out24 = {T build, Q gather, V1 = A Q, V2 = W U, V store, projection, U store + norms, norm reduce,
 * cluster barrier, tail; inside graph 0's lean LAPs: auction scans, bid resolution, augmentations, certificate, its
 * reachability part, its Kahn part}, [16..23] = free rows at the start of each auction round / after the last, summed over
 * graph 0's LAPs.  Synchronises the device. */
int ttdg_gagm_read_profile(int64_t *out24);
/* ---------------------------------------------------------------------------------------------
 * GA-GM solver.  Replaces GA_GM.forward + gagm (multi_graph_matching.py:223-244, 300-389) for the
 * configuration the hot path uses (num_clusters = 1, projector0 = 'sinkhorn', hung_iter = True) including
 * the per-iteration projector: batched Sinkhorn (mgm:330-353) or per-graph Hungarian (mgm:324-328), the
 * G == 2 identity quirk (mgm:358-359) and both convergence tests (mgm:361) - all on the device, no
 * host round trip.  One thread-block cluster (one CTA per graph, up to 8; more graphs share CTAs).
 * A, W: M x M; U0, U: M x n_univ (n_univ == 32); ms_h: HOST int32[G].
 * mode: 0 = full solve; 1 = exactly one iteration with projector `step_projector` (0 sinkhorn, 1 hungarian)
 * at temperature init_tau (teacher-forced parity tests).
 * info (int32[16], device): {iterations, sinkhorn-stage iterations, hungarian-stage iterations, LAP calls,
 *                           sinkhorn stages, Dijkstra steps / path hops / certificate fall-backs of graph 0's LAPs,
 *                           cycle accounting of CTA 0 in units of 1024 cycles: whole kernel, Hungarian-stage iterations,
 *                           inside the LAP, waiting at cluster barriers; Sinkhorn-stage iterations: phase 1, V products, projector;
 *                           [15] = 1 if a Hungarian projection met a NaN / inf cost (zeroed so that the solve ends; SciPy raises
 *                           "matrix contains invalid numeric entries" there, the host mirror does the same in check_flags())}.
 * scratch: ttdg_gagm_scratch_bytes(M, G).
 * trace (optional, may be NULL): fp64[(trace_cap + 1)][M][32] receives U_t of every iteration t <= trace_cap
 * (U_0 = U0) and trace_meta fp64[trace_cap][2] = {projector, tau} - lets a test verify EVERY iteration of the
 * trajectory against the oracle's single step (the trajectory as a whole is chaotic).
 * --------------------------------------------------------------------------------------------- */
int64_t ttdg_gagm_scratch_bytes(int M, int G);
int ttdg_gagm_solve(const float *A, const float *W, const float *U0, const int32_t *ms_h, int G, int M,
                    int n_univ, double init_tau, double min_tau, double sk_gamma, int max_iter, int sk_iter,
                    double converge_tol, double quad_weight, int mode, int step_projector, float *U,
                    int32_t *info, void *scratch, double *trace, double *trace_meta, int trace_cap, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Matching loss.  Replaces collect_intra_class_matching_wrapper + the 'perm' loss loop (mgm:543-564,
 * 594-633) -> PermutationLoss / BCEFocalLoss (utils/losses.py:83-103, 419-455):
 *     loss = mean over pairs i1<i2 of mean_ij focal_bce(clamp(S_ij), (U_i1 U_i2^T)_ij)
 * S_ij is read from Wds (M x M) at the block the pairwise Sinkhorn wrote (rows of graph i2, columns of graph
 * i1 - mgm:507-525, :620-623).  loss: 1 float.  flags: int32[1], set non-zero if an S or target value left
 * [0, 1] (the reference asserts, losses.py:437-442).  The backward writes grad_Wds (M x M, zero outside
 * the touched blocks) scaled by *grad_loss.
 * --------------------------------------------------------------------------------------------- */
int64_t ttdg_matching_loss_scratch_bytes(int G);
int ttdg_matching_loss_fwd(const float *Wds, const float *U, const int32_t *node_off, int G, int M, int n_univ,
                           float *loss, int32_t *flags, void *scratch, void *stream);
int ttdg_matching_loss_bwd(const float *Wds, const float *U, const int32_t *node_off, int G, int M, int n_univ,
                           const float *grad_loss, float *grad_Wds, void *stream);
/* generic focal BCE on one dense array (utils/losses.py:83-103): mean reduction. */
int64_t ttdg_focal_bce_scratch_bytes(void);
int ttdg_focal_bce_fwd(const float *p, const float *y, int64_t n, float *loss, void *scratch, void *stream);
int ttdg_focal_bce_bwd(const float *p, const float *y, int64_t n, const float *grad_loss, float *grad_p,
                       void *stream);

/* ---------------------------------------------------------------------------------------------
 * Node sampler.  Replaces PrototypeComputation.__call__ (build_graph.py:160-250; targets :70-115,
 * locations :133-157).  Levels l = 0..4 with strides 4, 8, 16, 32, 64; lvl_hw_h = HOST int32[5][2] (H_l, W_l).
 * boxes: float[total_boxes][4] xyxy, classes int64[total_boxes], box_off int32[B+1] - one entry per LISTED
 * image; images without boxes must already be dropped by the caller, and listed image b reads feature-map
 * image b (the reference skips box-less images at :79 but indexes features by list position at :181).
 * ttdg_sampler_select: label[b][loc] (int32, L = sum_l H_l W_l locations per image), counts[b][l] = number of
 *   kept nodes, sel_idx[b][l][max_per_level] = flat location index (inside the image) of every kept node.
 * ttdg_sampler_gather: nodes (n_total x C) and labels (int64) in the reference order (image, level, location);
 *   node_off = device int32[B*5+1] prefix sums of counts.  Feature maps are read through strides
 *   (feat_strides_h = HOST int64[5][3] = {image, channel, pixel} element strides) so NCHW and NHWC both work;
 *   feat_ptrs_h = HOST array of 5 device pointers.
 * ttdg_sampler_scatter_bwd: gfeat[l][b, :, pixel] += grad_nodes[k, :] (kept locations are unique).
 * --------------------------------------------------------------------------------------------- */
int ttdg_sampler_select(const float *boxes, const int64_t *classes, const int32_t *box_off, int B,
                        const int32_t *lvl_hw_h, int sample_dist, int max_per_level, int32_t *label,
                        int32_t *counts, int32_t *sel_idx, void *stream);
int ttdg_sampler_gather(const float *const *feat_ptrs_h, const int64_t *feat_strides_h, const int32_t *lvl_hw_h,
                        int B, int C, int n_total, const int32_t *label, const int32_t *sel_idx,
                        int max_per_level, const int32_t *node_off, float *nodes, int64_t *labels_out,
                        void *stream);
int ttdg_sampler_scatter_bwd(const float *grad_nodes, float *const *gfeat_ptrs_h, const int64_t *feat_strides_h,
                             const int32_t *lvl_hw_h, int B, int C, int n_total, const int32_t *sel_idx,
                             int max_per_level, const int32_t *node_off, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused SGD step over one flat fp32 bucket.  Replaces the torch.optim.SGD the caller steps at
 * engine/trainer.py:480-482 (Detectron2 build_optimizer: momentum 0.9, weight decay 1e-4, lr from
 * configs/test_segment.yaml:28):  g = grad_scale * g + wd * p;  m = first_step ? g : momentum * m + g;
 * p -= lr * m.  grad_scale carries the 1 / world_size of the gradient all-reduce.  p, g, m 16-byte aligned.
 * --------------------------------------------------------------------------------------------- */
int ttdg_sgd_step(float *p, const float *g, float *m, int64_t n, float lr, float momentum, float weight_decay,
                  float grad_scale, int first_step, void *stream);

/* =============================================================================================
 * Detector (Detectron2 0.5 Mask R-CNN R50-FPN as configured by configs/Base-RCNN-FPN.yaml + test_segment.yaml;
 * driven by meta_arch/rcnn.py:219-226,331-357 and :181-182).  Activations are NHWC fp32; convolution weights are
 * [R][S][Cin][Cout] (the host wrapper permutes torch's [Cout][Cin][R][S]); channel counts are multiples of 4.
 * ============================================================================================= */

/* Convolution forward with fused epilogue  y = relu?( conv(x) * scale[c] + bias[c] + residual )  - FrozenBN folded
 * into (scale, bias) (d2 FrozenBatchNorm2d), conv bias, the bottleneck shortcut add (res_mode 1: residual has y's
 * shape) or the FPN top-down add (res_mode 2: residual is the half-resolution map, nearest-upsampled on the fly).
 * Replaces the cuDNN calls behind d2 ResNet/FPN/RPN-head/mask-head convs and the box-head FCs (as 1x1 convs on
 * N = rows, H = W = 1). */
int ttdg_conv_fwd(const float *x, const float *w, const float *scale, const float *bias, const float *residual,
                  int res_mode, int relu, int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad,
                  float *y, void *stream);
/* Data gradient (autograd of the above for the layers the TTT loss reaches: res3-res5 and FPN).
 * (N, H, W, Cin) is the forward input geometry; stride 1 any R x S, or 1x1 stride 2 (dx must be zero-filled). */
int ttdg_conv_dgrad(const float *dy, const float *w, int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                    int pad, float *dx, void *stream);
/* Weight gradient, ACCUMULATED into dw [R][S][Cin][Cout] (split over pixels, fp32 atomics). Cin % 128 == 0. */
int ttdg_conv_wgrad(const float *x, const float *dy, int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                    int pad, float *dw, void *stream);
/* out = (y > 0 ? g : 0) * scale[c]: backward through ReLU (mask from the stored output y, may be NULL) and the
 * FrozenBN scale (may be NULL). */
int ttdg_relu_bn_bwd(const float *g, const float *y, const float *scale, int C, int64_t numel, float *out, void *stream);
/* the same in one pass with both results: out_pre = g * [y > 0] (gradient of the residual branch), out_conv = out_pre * scale */
int ttdg_relu_bn_bwd2(const float *g, const float *y, const float *scale, int C, int64_t numel, float *out_pre, float *out_conv,
                      void *stream);
/* out[c] += sum over pixels of g[pixel][c]  (conv bias gradient) */
int ttdg_bias_grad(const float *g, int64_t pixels, int C, float *out, void *stream);
int ttdg_maxpool3x3s2(const float *x, int N, int H, int W, int C, float *y, void *stream);           /* ResNet stem pool */
/* mode 0: y[small] = x[big at even pixels] (p6 = max_pool2d(p5, 1, 2)); 1: y[big even] += x[small] (its backward);
 * 2: y[small] += sum of 2x2 children of x[big] (backward of the FPN nearest upsample).  Hs x Ws = small grid. */
int ttdg_resample2(const float *x, float *y, int N, int Hs, int Ws, int C, int mode, void *stream);
/* uint8 N x 3 x H x W planar -> fp32 NHWC with 4 channels (4th = 0), minus PIXEL_MEAN (d2 preprocess_image).  Output rows
 * hold Wp >= left + W pixels: `left` zero pixels, the image, zeros up to Wp (left = 0, Wp = W: plain layout). */
int ttdg_preprocess(const unsigned char *img_u8, int N, int H, int W, int Wp, int left, float mean0, float mean1, float mean2,
                    float *out, void *stream);

/* Tensor-core version of ttdg_conv_fwd / stride-1 ttdg_conv_dgrad: tcgen05.mma.kind::tf32 fed by TMA (activations as a 4-D
 * NHWC tensor map - filter taps are coordinate shifts, padding is TMA's out-of-bounds zero fill), fp32 accumulators in
 * TMEM, same fused epilogue.  Persistent: one CTA per SM walks its (pixel tile, N tile) items.  Requires Cin % 32 == 0, Cout % 64 == 0, stride 1 (ttdg_conv_tc_supported).
 *   x          : N x H x W x Cin, fp32.
 *   wk_hi, wk_lo: weights K-major [taps][n][k], split into hi = tf32(w) and lo = w - hi: forward = [R*S][Cout][Cin]
 *                (ttdg_weight_transpose_split of the [R][S][Cin][Cout] parameter); data gradient (flip = 1,
 *                pad = R-1-pad_fwd, x = dY) = the parameter array itself read as [R*S][n = Cin_fwd][k = Cout_fwd], split
 *                with ttdg_tf32_split.  wk_lo == NULL -> single-pass TF32.  Otherwise "3xTF32": hi*hi + lo*hi + hi*lo
 *                gives fp32-grade products (parity config); the activations are split inside the kernel's pipeline,
 *                between the TMA arrival and the MMA, into TENSOR MEMORY (the A operand of the MMAs comes from TMEM).
 *   (Cin, Cout) are the GEMM's k and n extents.
 *   Strided 1x1 convs (the first block of res3 / res4 / res5, STRIDE_IN_1X1): in_stride = 2 reads x[n, 2 ho, 2 wo] through
 *   TMA element strides (forward); out_stride = 2 stores the result for (ho, wo) at (2 ho, 2 wo) of a zero-filled
 *   outH x outW map (data gradient).  R = S = 1, pad = 0 only; outH / outW are ignored when out_stride == 1. */
int ttdg_conv_tc_supported(int Cin, int Cout, int stride);
/* Thread-block-cluster size along the pixel tiles for ttdg_conv_tc: the CTAs of a cluster share the weight tile, each
 * loads 1/cl of it and TMA multicasts the slice to all of them.  cl = 1 (default; also env TTDG_TC_CLUSTER), 2 or 4.
 * Results do not depend on it.  Returns the previous value, or TTDG_E_ARG. */
int ttdg_conv_tc_set_cluster(int cl);
/* Epilogue of ttdg_conv_tc / ttdg_conv_tc_bf16 / ttdg_stem_tc (env TTDG_TC_EPI): 0 = every thread stores its own pixel row (32
 * lines per warp access); 1 = each epilogue warp transposes its 32 pixel rows through 4 KB of shared memory so that global
 * loads (residual) and stores are 128-byte row segments; 2 = 1 unless the layer adds a residual; 3 (default) = the tile goes
 * through shared memory as TMA boxes (TMA store, TMA residual load) where the layer allows it - fp32 output at stride 1,
 * residual absent or fp32 at the output's resolution - else as 2.  Same arithmetic in the same order: results are
 * bit-identical.  Returns the previous value, or TTDG_E_ARG. */
int ttdg_conv_tc_set_epilogue(int mode);
/* Persistent tensor-core kernels (ttdg_conv_tc*, ttdg_wgrad_tc*, ttdg_stem_tc*) launch one CTA per SM and give every CTA the same
 * share of the tiles.  When another kernel holds some SMs for a long time - the GA-GM solver's cluster, 8 SMs for ~10 ms, while the
 * evaluation pass of the previous dataset runs beside it (adapteacher/engine/trainer.py) - CTAs that find no SM would start only when
 * others finish and double the kernel's duration; n > 0 caps the grid of the launches that follow at n SMs, 0 (default) = all.
 * Host-side launch parameter only; returns the previous value. */
int ttdg_set_sm_limit(int n);
/* Diagnostics: while dev_buf != NULL, CTA 0 of every ttdg_conv_tc* launch records clock64() at 8 points of its first `items`
 * tiles into dev_buf[item * 8 + slot] (slot 0 / 1: first / last k-block's TMA issue, 2: MMA warp owns the accumulator,
 * 3: last k-block's operands ready, 4: last commit issued, 7 / 5 / 6: epilogue warp 0 starts the tile / has drained the
 * accumulator / has issued its stores).  dev_buf = NULL or items = 0 switches it off (the default). */
int ttdg_conv_tc_set_trace(long long *dev_buf, int items);
int ttdg_conv_tc(const float *x, const float *wk_hi, const float *wk_lo, const float *scale, const float *bias,
                 const float *residual, int res_mode, int relu, int flip, int N, int H, int W, int Cin, int Cout, int R,
                 int S, int pad, int in_stride, int out_stride, int outH, int outW, float *y, void *stream);
/* Weight gradient on tensor cores (stride-1 convs and 1x1 stride-2 convs, Cin % 128 == 0, Cout % 64 == 0): dw [R][S][Cin][Cout] +=
 * sum_pixels X[pixel + tap][ci] dY[pixel][co].  Both operands are MN-major in memory (pixels = GEMM k = slow dimension):
 * TMA boxes of {32 channels, 32 pixels}.  precise != 0 = 3xTF32: X is staged unswizzled and split by the kernel into TENSOR
 * MEMORY (A operand from TMEM, K-major by construction), dY is split in place in shared memory with the 128-byte / 32-byte-
 * atom swizzle (the only MN-major layout tcgen05 takes for 32-bit operands); 0 = single-pass TF32 with both operands from
 * shared memory.  Accumulated into dw with coalesced fp32 reductions over the pixel splits (dw may be the parameter's slice
 * of the flat gradient bucket).  Replaces the weight half of torch autograd's conv
 * backward behind loss.backward() (engine/trainer.py:481). */
int ttdg_wgrad_tc_supported(int Cin, int Cout, int stride);
int ttdg_wgrad_tc(const float *x, const float *dy, int precise, int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                  int pad, float *dw, void *stream);
/* d2 BasicStem (7x7 stride 2 pad 3, 3 -> 64, FrozenBN, ReLU) on tensor cores from the padded image of ttdg_preprocess
 * (left = 3, Wp >= W + 8): the 28 floats of one filter row are a contiguous window, read as overlapping TMA boxes.
 * wk_hi / wk_lo [7][64][32]: wk[r][co][4 s + c] = w[r][s][c][co], zero for s = 7 (wk_lo NULL = single-pass TF32).
 * y: N x H/2 x W/2 x 64; H, W even. */
int ttdg_stem_tc(const float *x_pad, int Wp, const float *wk_hi, const float *wk_lo, const float *scale, const float *bias,
                 int relu, int N, int H, int W, float *y, void *stream);
int ttdg_tf32_split(const float *x, float *hi, float *lo, int64_t numel, void *stream);
/* w [taps][Cin][Cout] -> wt_hi, wt_lo (may be NULL) [taps][Cout][Cin] */
int ttdg_weight_transpose_split(const float *w, int taps, int Cin, int Cout, float *wt_hi, float *wt_lo, void *stream);
/* All tensor-core weight copies of the adapted convolutions in ONE launch after an optimizer step (instead of one small launch
 * per layer and variant on first use).  jobs_dev: device int64[njobs][8] = { src, dst_hi, dst_lo (0 = none), taps, Cin, Cout,
 * mode, first_tile } with mode 0 = K-major transposed tf32 hi / lo (forward), 1 = same-layout hi / lo (data gradient),
 * 2 = K-major transposed bf16 into dst_hi; tiles are 32 x 32 per tap, first_tile ascending, total_tiles their sum. */
int ttdg_weight_refresh(const int64_t *jobs_dev, int njobs, int64_t total_tiles, void *stream);

/* ---------------------------------------------------------------------------------------------
 * bf16 backbone (BASELINE.json configs[2]: "bf16 backbone / fp32 Sinkhorn").  Replaces the same Detectron2 ResNet-50-FPN
 * convolutions as ttdg_conv_tc (meta_arch/rcnn.py:226, configs/Base-RCNN-FPN.yaml:3-8) with bf16 activations and weights in
 * HBM: tcgen05.mma.kind::f16, fp32 accumulation in TMEM, fp32 epilogue (FrozenBN scale / bias, residual, ReLU).
 * x: N x H x W x Cin bf16 (Cin % 64 == 0); wk: [taps][n][k] bf16 K-major (ttdg_weight_transpose_bf16 of the fp32 master
 * weights, which stay in the optimizer's flat bucket); residual: bf16 when res_bf16 else fp32; y: bf16 when out_bf16 else fp32
 * (the FPN output convolutions hand an fp32 pyramid to the heads and to the matching stage).  Other arguments as ttdg_conv_tc.
 * ttdg_stem_tc2 = ttdg_stem_tc with an optional bf16 output; the two elementwise helpers are the stem max pool on bf16 maps
 * and the ReLU-mask / FrozenBN-scale backward with the stored bf16 output as the mask (gradients stay fp32).
 * --------------------------------------------------------------------------------------------- */
int ttdg_conv_tc_bf16(const void *x, const void *wk, const float *scale, const float *bias, const void *residual, int res_bf16,
                      int res_mode, int relu, int flip, int N, int H, int W, int Cin, int Cout, int R, int S, int pad,
                      int in_stride, int out_stride, int outH, int outW, void *y, int out_bf16, void *stream);
int ttdg_weight_transpose_bf16(const float *w, int taps, int Cin, int Cout, void *wt_bf16, void *stream);
int ttdg_stem_tc2(const float *x_pad, int Wp, const float *wk_hi, const float *wk_lo, const float *scale, const float *bias,
                  int relu, int N, int H, int W, void *y, int out_bf16, void *stream);
int ttdg_maxpool3x3s2_bf16(const void *x, int N, int H, int W, int C, void *y, void *stream);
int ttdg_relu_bn_bwd_bf16y(const float *g, const void *y_bf16, const float *scale, int C, int64_t numel, float *out, void *stream);

/* RPN: decode + clip the anchors selected per level (d2 find_top_rpn_proposals / Box2BoxTransform.apply_deltas).
 * idx [n_img][k] indexes (pixel * A + a); deltas = NHWC head output, channel a * 4 + c, row pitch ld_deltas;
 * cell_anchors_h = HOST float[A][4].  valid = finite and non-empty after clipping. */
int ttdg_rpn_decode(const float *deltas, int ld_deltas, const int64_t *idx, int n_img, int k, int Hl, int Wl, int A,
                    int stride, const float *cell_anchors_h, float img_h, float img_w, float *boxes,
                    unsigned char *valid, void *stream);
/* Box head output (FastRCNNOutputLayers.inference up to the score filter): softmax over K+1 class scores,
 * class-specific deltas with weights (10, 10, 5, 5), clip; cand_scores = -1 where score <= thresh. */
int ttdg_box_predict(const float *cls, int ld_cls, const float *reg, int ld_reg, const float *proposals, int R, int K,
                     float img_h, float img_w, float score_thresh, float *cand_boxes, float *cand_scores, void *stream);
/* Per-category NMS (torchvision batched_nms semantics) for `batch` images at once: boxes_sorted [batch][n][4] already
 * sorted by descending score, category [batch][n] (boxes of different categories never suppress each other).
 * keep [batch][max_keep] receives indices in order, n_keep [batch] their number.
 * scratch: ttdg_nms_scratch_bytes(batch, n) (used for n > 2560: suppression-mask kernel over the upper triangle + sweep;
 * smaller problems run in one CTA per image with boxes and state in shared memory). */
int64_t ttdg_nms_scratch_bytes(int batch, int n);
int ttdg_nms(const float *boxes_sorted, const int32_t *category, int batch, int n, float iou_thresh, int max_keep,
             int32_t *keep, int32_t *n_keep, void *scratch, void *stream);
/* ---------------------------------------------------------------------------------------------
 * Device-side candidate selection (no host round trip between the RPN head and the box head).  Replaces Detectron2
 * `find_top_rpn_proposals` (proposal_generator/rpn.py:52-54 -> predict_proposals) and the candidate ordering of
 * `fast_rcnn_inference_single_image` (roi_heads/roi_heads.py:173-205), which the reference runs as torch sort / top-k /
 * boolean indexing with data-dependent shapes.
 * ttdg_rpn_select: per (image, level) radix-select top-k of the H*W*A objectness logits, sorted descending (ties: lower anchor
 *   index first), anchors decoded (Box2BoxTransform weights 1, clamp log(1000/16)), clipped to the image's own size.
 *   logits_h / deltas_h: HOST arrays of n_levels device pointers (N x H x W x ld, channel a resp. 4a + c); lvl_hw_h {H, W} per
 *   level, k_h[l] = min(H*W*A, pre_nms_topk) <= 2048; img_hw_h {h, w} per image.  Outputs [n_img][sum k]: boxes, scores, valid
 *   (finite and non-empty), level l occupying [off_l, off_l + k_l).
 * ttdg_sort_candidates: per image, candidates ordered by (valid, score descending, position ascending); category for ttdg_nms =
 *   cats_in[position] (cat_mod == 0) or position % cat_mod, invalid -> unique negative; valid == NULL means score > 0.
 * ttdg_gather_kept: the kept (ttdg_nms) valid candidates padded to max_keep rows + per-image counts (device).
 * ttdg_rois_from_padded / ttdg_mask_padded_candidates: padded proposals -> RoIAlign rois; candidates of padding rows -> -1.
 * --------------------------------------------------------------------------------------------- */
int ttdg_rpn_select(const void *const *logits_h, const void *const *deltas_h, const int32_t *lvl_hw_h, const int32_t *strides_h,
                    const int32_t *k_h, int n_levels, int ld_logits, int ld_deltas, int A, const float *cell_anchors_h, int n_img,
                    const float *img_hw_h, float *boxes, float *scores, unsigned char *valid, void *stream);
int ttdg_sort_candidates(const float *boxes, const float *scores, const unsigned char *valid, const int32_t *cats_in, int cat_mod,
                         int n_img, int n, float invalid_score, float *boxes_out, float *scores_out, int32_t *cats_out,
                         int32_t *n_valid, void *stream);
int ttdg_gather_kept(const float *boxes, const float *scores, const int32_t *cats, const int32_t *keep, const int32_t *n_keep,
                     const int32_t *n_valid, int n_img, int n, int max_keep, float pad_score, float *boxes_out, float *scores_out,
                     int64_t *cats_out, int32_t *counts, void *stream);
/* RPN NMS without the cross-level sweep (find_top_rpn_proposals applies batched_nms with the level as category, so levels never
 * interact): ttdg_rpn_nms_levels runs one shared-memory greedy NMS per (image, level) on ttdg_rpn_select's output (every level's
 * segment of k_h[l] candidates is sorted by score; invalid candidates are skipped) and writes kept[img][pos] (uint8);
 * ttdg_top_candidates orders the candidates with valid[] != 0 by (score descending, position ascending) and writes the first
 * n_out of them, padded with a zero box / pad_score, plus counts[img] = min(number valid, n_out).  Together they equal
 * ttdg_sort_candidates + ttdg_nms + ttdg_gather_kept on the level-concatenated list. */
int ttdg_rpn_nms_levels(const float *boxes, const unsigned char *valid, const int32_t *k_h, int n_levels, int n_img,
                        float iou_thresh, int max_keep, unsigned char *kept, void *stream);
int ttdg_top_candidates(const float *boxes, const float *scores, const unsigned char *valid, int n_img, int n, int n_out,
                        float pad_score, float *boxes_out, float *scores_out, int32_t *counts, void *stream);
int ttdg_rois_from_padded(const float *boxes, const int32_t *counts, int n_img, int P, float *rois, void *stream);
int ttdg_mask_padded_candidates(float *cand_scores, const int32_t *counts, int n_img, int P, int K, void *stream);

/* ROIPooler + ROIAlignV2 (aligned, sampling_ratio 0) over p2..p5: rois [n][5] = {image, x0, y0, x1, y1};
 * out [n][pooled][pooled][C].  feat_ptrs_h = HOST array of 4 device pointers, lvl_hw_h = HOST int32[4][2]. */
int ttdg_roi_align(const float *const *feat_ptrs_h, const int32_t *lvl_hw_h, const float *rois, int n_rois, int C,
                   int pooled, float *out, void *stream);
/* x [R][H][W][(a, b, c)] -> y [R][2H][2W][c]  (the 2x2 stride-2 deconv of the mask head, computed as a 1x1 conv) */
int ttdg_pixel_shuffle2(const float *x, int R, int H, int W, int C, float *y, void *stream);
/* detector_postprocess / paste_masks_in_image: out[r] = bilinear(sigmoid(logits[r, :, :, class[r]])) >= threshold
 * over the H x W image (grid_sample semantics, align_corners = False). */
int ttdg_mask_paste(const float *logits, int ld_logits, int M, const float *boxes, const int64_t *classes, int R, int H,
                    int W, float threshold, unsigned char *out, void *stream);

/* On-device half of DiceEvaluator.process (adapteacher/evaluation/dice_metric.py:25-92): for BINARY masks Dice (:57-58),
 * the E-measure (:110-143) and the S-measure (:146-240) are functions of pixel counts, so the <= 100 x H x W predicted
 * masks per image never leave the GPU.
 *   ttdg_mask_gt_stats  : gt G x H x W (uint8, nonzero = set) -> stats int64[G][5] = {n, sum of rows, sum of columns,
 *                         split_y, split_x}, split = int(round(centre of mass)) + 1 with round-half-to-even (:227-229)
 *   ttdg_mask_pair_counts: pairs int32[n_pairs][2] = (prediction index, ground-truth index) ->
 *                         counts int64[n_pairs][4 quadrants: TL, TR, BL, BR][n11, n10, n01, n00] (prediction first),
 *                         quadrants split at the ground truth's (split_y, split_x). */
int ttdg_mask_gt_stats(const unsigned char *gt, int G, int H, int W, int64_t *stats, void *stream);
int ttdg_mask_pair_counts(const unsigned char *pred, const unsigned char *gt, const int32_t *pairs, int n_pairs,
                          const int64_t *gt_stats, int H, int W, int64_t *counts, void *stream);

/* ---- test data path: image resize on the device.  Replaces the host-side PIL resize of d2's
 * DatasetMapper(cfg, False) -> ResizeShortestEdge -> ResizeTransform.apply_image (reference adapteacher/data/build.py:122-154),
 * bit-exact with PIL.Image.resize((w, h), Image.BILINEAR) on uint8 images (Pillow's 8-bit resampler: antialiasing triangle filter,
 * 22-bit fixed-point coefficients, horizontal pass then vertical pass, each clipped to uint8).
 * ttdg_resize_ksize / ttdg_resize_coeffs_u8 are HOST functions (no device work): the coefficient table of one axis, built in
 * double precision with Pillow's expressions.  bounds_h: 2 * out_size int32 (first input index, count); kk_h: out_size * ksize.
 * ttdg_resize_bilinear_u8: src H x W x C uint8 interleaved (C = 1, 3, 4) -> dst nh x nw interleaved (planar = 0) or as C planes
 * (planar = 1: the uint8 C x H x W tensor ttdg_preprocess reads), channels reversed when flip (INPUT.FORMAT "BGR").  The x tables
 * (device copies for (W, nw)) are ignored when nw == W, the y tables when nh == H; tmp: H * nw * C bytes of device scratch. */
int ttdg_resize_ksize(int in_size, int out_size);
int ttdg_resize_coeffs_u8(int in_size, int out_size, int32_t *bounds_h, int32_t *kk_h);
int ttdg_resize_bilinear_u8(const unsigned char *src, int H, int W, int C, const int32_t *bounds_x, const int32_t *kk_x, int ksize_x,
                            const int32_t *bounds_y, const int32_t *kk_y, int ksize_y, int nh, int nw, unsigned char *tmp,
                            unsigned char *dst, int planar, int flip, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TTDG_B200_H */
