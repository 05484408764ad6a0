/* pleas_b200 — C ABI of the B200-native PLeaS-Merging merge hot path.
 *
 * Plain pointers, sizes and a CUDA stream handle; no torch types, no exceptions, no
 * allocation of caller-visible memory.  All pointers are DEVICE pointers unless a
 * parameter is documented as host memory.  Every function enqueues on `stream` (a
 * cudaStream_t passed as void*; NULL = legacy default stream), is safe to capture in a
 * CUDA graph unless noted, and returns
 *     0   success
 *    <0   invalid argument (PLB_EINVAL ...)
 *    >0   CUDA error code (cudaError_t) raised while launching
 * plb_last_error_string() describes the last failure on the calling thread.
 *
 * The reference (SewoongLab/PLeaS-Merging) is pure Python; each entry point names the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add.
 */
#ifndef PLEAS_B200_H
#define PLEAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLB_OK 0
#define PLB_EINVAL (-1)
#define PLB_ESIZE (-2)   /* problem too large for the kernel's on-chip working set */
#define PLB_EALIGN (-3)  /* pointer / leading dimension not aligned as required */

/* cross-statistic modes (pleas/methods/activation_matching.py:14-46) */
#define PLB_MODE_INNER 0      /* G = X Y^T            cross_features_inner_product :14-28 */
#define PLB_MODE_NEG_CDIST 1  /* -sqrt(max(qa+qb-2G,0)) cross_features_cdist      :31-46 */
#define PLB_MODE_CORR 2       /* Pearson correlation of the unit pairs from G, the row sums and the row sums of
                                 squares: (G - sa sb/K) / sqrt((qa - sa^2/K)(qb - sb^2/K)); 0 for a unit without
                                 variance.  Not in the reference (it ships the two siblings above): the third
                                 statistic BASELINE.json's north_star names; plb_cross_finalize_corr */

int plb_version(void);
const char *plb_last_error_string(void);

/* ---------------------------------------------------------------------------------------
 * Packed operand planes.
 *
 * A logical operand is a row-major view X[rows, K] of a tensor laid out as
 * [outer][rows][inner] (K = outer*inner, k = o*inner + i): the reference's
 * `movedim(x, a, 0).reshape(x.shape[a], -1)` (activation_matching.py:26-27, 44-45) without
 * the transposed copy.  plb_pack_split writes it as two fp32 planes hi = tf32(x) and
 * lo = tf32(x - hi) in the tcgen05 K-major no-swizzle core-matrix order
 *     plane[kb][g][j][r][e],  row = 8 g + r,  k = 16 kb + 4 j + e
 * with `row_groups` (>= ceil(rows/8)) groups of 8 rows per 16-wide k-block, so that any
 * (up to 128·t rows x 16 k) operand tile is one contiguous run that a single cp.async.bulk can
 * land in shared memory ready for tcgen05.mma; the GEMM copies only the row groups that exist.
 * Rows >= rows inside the last group and k >= K are written as 0.
 * --------------------------------------------------------------------------------------- */

/* bytes of ONE plane for a [rows, K] operand; *row_groups / *k_blocks receive the padded
 * geometry (row_groups = ceil(rows/8), k_blocks = ceil(K/16)).  Host-only helper. */
int64_t plb_plane_bytes(int64_t rows, int64_t K, int32_t *row_groups, int32_t *k_blocks);

/* Splits and packs one operand; optionally accumulates per-row sum of squares and sum into
 * fp64 vectors (atomic adds: zero them first).  `row_index` (may be NULL) gathers rows:
 * packed row r reads source row row_index[r] (int64) — used to apply a permutation on the
 * fly (utils.py:233-246 without the index_select copy).  kb_offset places the operand at
 * k-block kb_offset of a plane holding several K-concatenated operands (weight matching's
 * sum over state axes, weight_matching.py:66-75). */
int plb_pack_split(const float *x, int64_t outer, int64_t src_rows, int64_t inner,
                   const int64_t *row_index, int64_t rows,
                   float *hi, float *lo, int32_t row_groups, int32_t kb_offset,
                   double *row_sumsq, double *row_sum, void *stream);

/* Both operands of one activation tap (the same tap of the two models: equal [outer][rows][inner]
 * geometry, all rows, no gather) in one launch — the form activation matching uses for every tap
 * (activation_matching.py:87-92 inserts one cross_features call per tapped node pair). */
int plb_pack_split_pair(const float *xa, const float *xb, int64_t outer, int64_t src_rows, int64_t inner,
                        float *hi_a, float *lo_a, float *hi_b, float *lo_b, int32_t row_groups,
                        int32_t kb_offset, double *sumsq_a, double *sumsq_b, void *stream);
/* same, additionally accumulating the rows' plain sums (correlation statistic) */
int plb_pack_split_pair_sums(const float *xa, const float *xb, int64_t outer, int64_t src_rows, int64_t inner,
                             float *hi_a, float *lo_a, float *hi_b, float *lo_b, int32_t row_groups,
                             int32_t kb_offset, double *sumsq_a, double *sumsq_b, double *sum_a,
                             double *sum_b, void *stream);

/* Two-source gather-average + im2col pack for the PLeaS normal equations
 * (pleas_merging.py:116-123, 146-147: X-bar = cat[(x1[bi1]+x2[bi2])/2, x1[bi1c], x2[bi2c]]).
 * Packed row f = (c, dy, dx) of the merged layer input, k = (n, ho, wo).  chan1/chan2 give,
 * per merged channel c, the source channel in x1 / x2 (-1 = absent) and scale1/scale2 the
 * weights (0.5/0.5 merged, 1/0 or 0/1 separate).  x1, x2: [N, Cin, H, W].  With
 * ones_row != 0 one extra all-ones row (the bias feature) is appended. */
int plb_pack_im2col(const float *x1, const float *x2, int64_t N, int64_t cin_src, int64_t H, int64_t W,
                    const int32_t *chan1, const int32_t *chan2, const float *scale1, const float *scale2,
                    int64_t cmerged, int32_t kh, int32_t kw, int32_t stride_h, int32_t stride_w,
                    int32_t pad_h, int32_t pad_w, int32_t dil_h, int32_t dil_w, int64_t Ho, int64_t Wo,
                    int32_t ones_row, float *hi, float *lo, int32_t row_groups, int32_t kb_offset,
                    void *stream);

/* Fused cross-Gram of one NARROW tap (C <= 128 channels, C % 8 == 0, inner % 16 == 0): reads the
 * two fp32 activations x, y [outer][C][inner] directly (no packed planes: 4 instead of 20 bytes
 * of HBM traffic per element), splits to tf32 hi/lo on the way into shared memory and contracts
 * with the 3xTF32 tcgen05 pipeline.  Writes partial[s] = X[:, Ks] Y[:, Ks]^T for `splits` K ranges
 * as [splits][128][bn] fp32 with bn = 64 (C <= 64) or 128, to be reduced by plb_cross_finalize
 * (ld_m = 128, ld_n = bn), and adds the rows' sums of squares to the fp64 vectors (atomics: zero
 * them first; both or neither).  chain_kb = k-blocks per tensor-core accumulation chain (4).
 * Replaces cross_features_inner_product / cross_features_cdist (activation_matching.py:14-46)
 * for the HBM-bound taps. */
int plb_gram_direct(const float *x, const float *y, int64_t outer, int64_t C, int64_t inner,
                    float *partial, int32_t splits, int32_t chain_kb, double *row_sumsq_x,
                    double *row_sumsq_y, void *stream);

/* TMA-fed fused cross-Gram of one tap of ANY width, straight from the two fp32 activations
 * x, y [outer][C][inner] (inner % 4 == 0: the tensor is then a legal 3-D tensor map): no packed
 * planes.  TMA boxes of (rows x 32 k) land in 128-byte-swizzled shared memory, converter warps
 * derive the tf32 lo plane in place, the contraction is the 3xTF32 tcgen05 pipeline with in-kernel
 * promotion.  C <= 128: one CTA per work item (tile 128 x 64 / 128 x 128); C > 128: a CTA PAIR
 * (cluster of 2, tcgen05 cta_group::2) per 256 x 256 tile.  Writes partial[s] = X[:, Ks] Y[:, Ks]^T
 * for `splits` K ranges as [splits][ld_m][ld_n] fp32 (geometry from plb_gram_tma_geometry), to be
 * reduced by plb_cross_finalize, and adds the rows' sums of squares to the fp64 vectors (atomics:
 * zero them first; both or neither); row_sum_x / row_sum_y (may be NULL) additionally receive the
 * rows' plain sums for the correlation statistic.  chain_kb = 16-wide k-blocks per accumulation
 * chain (4).  Replaces cross_features_inner_product / cross_features_cdist
 * (activation_matching.py:14-46) together with its movedim/reshape copy (:26-27, 44-45). */
int plb_gram_tma(const float *x, const float *y, int64_t outer, int64_t C, int64_t inner,
                 float *partial, int32_t splits, int32_t chain_kb, double *row_sumsq_x,
                 double *row_sumsq_y, double *row_sum_x, double *row_sum_y, void *stream);

/* Experiments only (profiles/experiments/tma_trace.py): CTA 0 of every later plb_gram_tma launch records four
 * clock64 stamps per pipeline box into dev_buf (>= 1024 uint64); NULL switches tracing off. */
int plb_debug_set_trace(unsigned long long *dev_buf);

/* Host-only: tiling plb_gram_tma uses for C rows: CTAs per work item (1 or 2), output tiles and the
 * padded leading dimensions of the partial tiles. */
int plb_gram_tma_geometry(int64_t C, int32_t *cta_group, int32_t *m_tiles, int32_t *n_tiles,
                          int32_t *ld_m, int32_t *ld_n);

/* ---------------------------------------------------------------------------------------
 * 3xTF32 tcgen05 GEMM over packed planes:  partial[s] = A[:, Ks] B[:, Ks]^T per K split s.
 * One table entry per problem; problems sharing a tile width are launched together
 * (grouped launch).  The table lives in DEVICE memory (build it once per plan).
 * --------------------------------------------------------------------------------------- */
typedef struct PlbGemmProblem {
  const float *a_hi, *a_lo; /* A planes */
  const float *b_hi, *b_lo; /* B planes */
  float *partial;           /* [splits][m_tiles*128][n_tiles*bn] fp32 */
  int32_t a_row_groups, b_row_groups;
  int32_t k_blocks;         /* 16-wide k-blocks to contract */
  int32_t m_tiles, n_tiles, splits;
  int32_t cta_begin;        /* first CTA of this problem in the grouped grid */
  int32_t symmetric;        /* 1: A and B are the same operand; only the tiles touching the lower
                               triangle are enumerated: row-tile mt has min(n_tiles, (128*mt+127)/bn+1) */
} PlbGemmProblem;

/* bn in {64,128,256}.  total_ctas = sum over problems of (active tiles)*splits.
 * impl: 16*chain_kb = persistent tcgen05 3xTF32 kernel with in-kernel promotion every chain_kb
 * k-blocks (product path; total_ctas is then the number of work items, the grid is capped at
 * the SM count); 0 = one-CTA-per-work-item tcgen05 kernel (no promotion: the caller bounds the
 * chain with `splits`); 1 = SIMT fp32 FMA reference kernel on the same planes (cross-check only). */
int plb_gemm_grouped(const PlbGemmProblem *problems_dev, int32_t n_problems, int32_t total_ctas,
                     int32_t bn, int32_t impl, void *stream);

/* Sums the K-split partials and applies the cross-statistic epilogue
 *   cost[i, j] (+)= f(G_ij, qa_i, qb_j)      i < M, j < N
 * (activation_matching.py:28, 46; per-group accumulation :129-134).  accumulate = 0
 * overwrites.  cost64 != NULL selects an fp64 accumulator (normal equations).  sym_bn != 0:
 * the GEMM ran a symmetric problem with tile width sym_bn and only produced the tiles touching
 * the lower triangle; only those tiles of `cost` are written (they cover the lower triangle
 * including the diagonal) — the caller mirrors the accumulator when it needs the full matrix. */
int plb_cross_finalize(const float *partial, int32_t splits, int64_t ld_m, int64_t ld_n,
                       int64_t M, int64_t N, const double *qa, const double *qb, int32_t mode,
                       float *cost, double *cost64, int64_t ldc, int32_t accumulate, int32_t sym_bn,
                       void *stream);

/* All taps of a calibration batch in ONE launch (activation_matching.py:123-134: the per-tap cross features of a
 * batch, summed per permutation group).  For every group the taps tap_begin..tap_end-1 (in that order) are
 * reduced over their K splits, passed through the statistic's epilogue and added up; each cost entry is read
 * and written once: cost[i, j] (+)= sum_t f_t(G_t[i, j]).  Tables live in device memory; blocks of group g are
 * block_begin .. block_begin + B(n) - 1 with B(n) = ceil(n / 256) * ceil(n / 16) for n >= 256, else
 * ceil(n / 64) * ceil(n / 64); total_blocks is their sum.  sa / sb / K are only read in
 * PLB_MODE_CORR, qa / qb not in PLB_MODE_INNER. */
typedef struct PlbFinalizeTap {
  const float *partial;        /* [splits][ld_m][ld_n] fp32 */
  const double *qa, *qb;       /* row sums of squares */
  const double *sa, *sb;       /* row sums */
  int64_t ld_m, ld_n, K;
  int32_t splits;
  int32_t n_affine;            /* derived taps: the statistic of (s_a x + t_a, s_b y + t_b) per unit, e.g. the
                                  output of an eval-mode BatchNorm behind this tap, formed from the SAME Gram, row
                                  sums and sums of squares (needs sa / sb) — no second contraction */
  const double *affine;        /* n_affine x [s_a (n) | t_a (n) | s_b (n) | t_b (n)], n = the group's size */
} PlbFinalizeTap;
typedef struct PlbFinalizeGroup {
  float *cost;                 /* [n][ldc] fp32 */
  int64_t ldc;
  int32_t n, tap_begin, tap_end, block_begin;
} PlbFinalizeGroup;
int plb_cross_finalize_grouped(const PlbFinalizeTap *taps_dev, const PlbFinalizeGroup *groups_dev,
                               int32_t n_groups, int32_t total_blocks, int32_t mode, int32_t accumulate,
                               void *stream);

/* Correlation epilogue (PLB_MODE_CORR): cost[i, j] (+)= corr(x_i, y_j) from the K-split partials of
 * G = X Y^T, the rows' sums of squares qa/qb and sums sa/sb (fp64, as accumulated by plb_gram_tma /
 * plb_pack_split) and the contraction length K.  Same partial layout as plb_cross_finalize. */
int plb_cross_finalize_corr(const float *partial, int32_t splits, int64_t ld_m, int64_t ld_n,
                            int64_t M, int64_t N, const double *qa, const double *qb,
                            const double *sa, const double *sb, int64_t K, float *cost, int64_t ldc,
                            int32_t accumulate, void *stream);

/* ---------------------------------------------------------------------------------------
 * Batched linear sum assignment — replaces scipy_solve_lsa (pleas/core/solvers.py:18-33,
 * SciPy's Crouse shortest-augmenting-path in float64) including its tie-breaking rules, so
 * the returned assignment is identical to SciPy's.  One CTA per problem.
 * cost[p]: [n_p, n_p] fp32 row-major with leading dimension ld_p; col4row[p]: int64[n_p];
 * objective[p]: sum_i cost[i, col4row[i]] in fp64; status[p]: 0 ok, 1 infeasible, 2 invalid input
 * (NaN / -inf entries, as SciPy rejects them; or a leading dimension of 2^30 elements and more).
 * The four arrays-of-pointers / sizes are DEVICE arrays of length n_problems. n_p <= 4096.
 * --------------------------------------------------------------------------------------- */
int plb_lap_solve_batched(const float *const *cost, const int32_t *n, const int32_t *ld,
                          int64_t *const *col4row, double *objective, int32_t *status,
                          int32_t n_problems, int32_t max_n, int32_t maximize, void *stream);

/* The same solver started from column duals of a related problem (weight_matching.py:59-91 re-solves every group
 * once per sweep while the costs change little): v_in[p] (DEVICE array of device pointers, entries may be NULL = cold
 * start) holds fp64[n_p] column duals in the solver's minimisation form, scaled by v_scale; rows whose arg-min
 * reduced-cost column is free keep it, the others are inserted by shortest augmenting paths as usual.  The result
 * is an optimal assignment and equals plb_lap_solve_batched's whenever the optimum is unique (SciPy's tie-breaking
 * is only reproduced by the cold start).  v_out[p] (may be NULL) receives the final column duals. */
int plb_lap_solve_batched_warm(const float *const *cost, const int32_t *n, const int32_t *ld,
                               int64_t *const *col4row, double *objective, int32_t *status,
                               int32_t n_problems, int32_t max_n, int32_t maximize,
                               const double *const *v_in, double v_scale, double *const *v_out, void *stream);

/* ---------------------------------------------------------------------------------------
 * partial_merge building blocks (pleas/methods/partial_matching.py:47-176)
 * --------------------------------------------------------------------------------------- */

/* get_blocks for one group (:76-86): c_i = cost[i, P_i] (P = arange when identity != 0),
 * threshold = torch.quantile(c, ratio) (linear interpolation, fp32), mask = c >= threshold;
 * writes the order-preserving index lists Q[mask], P[mask], Q[~mask], P[~mask] (int64,
 * capacity n each) and counts[0] = number merged.  n <= 4096. */
int plb_get_blocks(const float *cost, int64_t ldc, const int64_t *perm, int32_t n, float ratio,
                   int32_t identity, int64_t *q_merged, int64_t *p_merged, int64_t *q_sep,
                   int64_t *p_sep, int32_t *counts, void *stream);

/* build_partial_merge_model's tensor assembly (:112-176) for one tensor viewed as
 * [O, I, R] (R = trailing elements per (o, i), e.g. kh*kw).  Output [no+2mo, ni+2mi, R].
 * Index lists may be NULL (= that axis is not in a permutation group: identity, no
 * separate part).  in_axis_only != 0 selects the 1-axis concatenation along axis 1 (:122-129),
 * where separate input units are copied un-halved. */
int plb_block_merge(const float *w1, const float *w2, int64_t O, int64_t I, int64_t R,
                    const int64_t *bo1, const int64_t *bo2, const int64_t *bo1c, const int64_t *bo2c,
                    int64_t no, int64_t mo,
                    const int64_t *bi1, const int64_t *bi2, const int64_t *bi1c, const int64_t *bi2c,
                    int64_t ni, int64_t mi, int32_t in_axis_only, float *out, void *stream);

/* apply_perm for one state axis (pleas/core/utils.py:244): out[o, p, i] = in[o, P[p], i]. */
int plb_gather_axis(const float *in, float *out, int64_t outer, int64_t n, int64_t inner,
                    const int64_t *P, void *stream);

/* perm composition (weight_matching.py:85): out[i] = a[b[i]] on int64. */
int plb_compose_perm(const int64_t *a, const int64_t *b, int64_t *out, int64_t n, void *stream);

/* weight_matching's progress test (weight_matching.py:80-81): *flag |= (newL > oldL + 1e-12) with
 * newL = sum_i A[i,P_i], oldL = sum_i A[i,i] compared as FLOAT32 values like the reference's torch
 * sums (each sum is formed in fp64 and rounded once); *gain (optional) = newL - oldL in fp64. */
int plb_wm_progress(const float *A, int64_t ld, const int64_t *P, int32_t n, int32_t *flag,
                    double *gain, void *stream);

/* ---------------------------------------------------------------------------------------
 * Forward convolution of the SOURCE models inside the calibration / PLeaS loops.
 *
 * The reference runs the two source models through ATen's fp32 convolutions
 * (pleas/methods/activation_matching.py:123 `gm_cross(x)`, pleas_merging.py:262-281 the hooked
 * forwards); on a B200 those are cuDNN's SIMT fp32 kernels and 60 % of a calibration step.  This
 * is the same arithmetic as an implicit GEMM on the tensor cores with the 3xTF32 split of the
 * Gram kernels (fp32-level accuracy, fp32 accumulation):
 *     out[n, co, oh, ow] = bias[co] + sum_{ci,kh,kw} x[n, ci, oh*s - ph + kh, ow*s - pw + kw] w[co, ci, kh, kw]
 * NCHW fp32 contiguous, groups = 1, dilation = 1, zero padding.
 *
 * plb_conv_pack_weights splits w into tf32 hi / lo planes in the order the kernel's tensor map
 * reads: planes[2][taps][Cout][Kc].  Cin % 32 == 0: taps = KH*KW, Kc = Cin, plane[t][co][ci].
 * Otherwise ("flat" form, e.g. the 3-channel stem): taps = 1, KWP = KW rounded up to a power of two,
 * Kc = ceil32(Cin*KH*KWP), plane[0][co][(ci*KH + kh)*KWP + kw], zero for kw >= KW and beyond Cin*KH rows: a
 * kernel row's taps are consecutive k, so the kernel's gather resolves (ci, kh) once per row.  Needs
 * Kc <= 512.  plb_conv_packed_floats returns the number of
 * floats of the whole packed buffer (host-only helper).  Weights are constants of the source
 * models: pack once, reuse for every batch. */
/* experiments only: per-box clock64 stamps of CTA 0 of later plb_conv2d_forward launches (NULL switches it off) */
int plb_conv_debug_set_trace(unsigned long long *dev_buf);
int64_t plb_conv_packed_floats(int64_t Cout, int64_t Cin, int32_t KH, int32_t KW);
int plb_conv_pack_weights(const float *w, int64_t Cout, int64_t Cin, int32_t KH, int32_t KW, float *packed,
                          void *stream);

/* One launch for up to two convolutions of identical geometry (the same layer of the two source
 * models).  x / out / bias (bias entries may be NULL) are HOST arrays of `nprob` device pointers. */
int plb_conv2d_forward(const float *const *x, const float *const *packed_w, const float *const *bias,
                       float *const *out, int32_t nprob, int64_t NB, int64_t Cin, int64_t IH, int64_t IW,
                       int64_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad_h, int32_t pad_w,
                       void *stream);

/* Same, with the per-channel affine layer behind the convolution fused into the epilogue: besides out, the
 * kernel writes out2[n, co, oh, ow] = f(scale[co] * out + shift[co]), f = ReLU when `relu`, else identity — an
 * eval-mode BatchNorm (scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale) with its ReLU,
 * i.e. the `conv -> bn -> relu` chain of the source models in one pass over the activation instead of three.
 * scale_shift / out2 are HOST arrays of nprob device pointers; scale_shift[i] is Cout interleaved (scale, shift)
 * fp32 pairs.  Both NULL: plain convolution. */
int plb_conv2d_affine_forward(const float *const *x, const float *const *packed_w, const float *const *bias,
                              float *const *out, const float *const *scale_shift, float *const *out2,
                              int32_t relu, int32_t nprob, int64_t NB, int64_t Cin, int64_t IH, int64_t IW,
                              int64_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad_h, int32_t pad_w,
                              void *stream);

/* ---------------------------------------------------------------------------------------
 * PLeaS closed form: solve (G + ridge*I) X = B for symmetric positive definite G (fp64,
 * column/row symmetric so layout-agnostic), in place: G is overwritten by its Cholesky
 * factor (lower), B [n, nrhs] row-major by the solution.  Replaces the Adam loop of
 * pleas/methods/pleas_merging.py:357-375.  info (device int32): 0 ok, k>0 = pivot k-1 was
 * not positive (after the ridge).  Not graph-capturable (launch count depends on n).
 * --------------------------------------------------------------------------------------- */
int plb_chol_solve(double *G, int64_t n, double *B, int64_t nrhs, double ridge, int32_t *info,
                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PLEAS_B200_H */
