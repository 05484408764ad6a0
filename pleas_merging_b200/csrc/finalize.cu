// Split-K reduction + cross-statistic epilogue.
//   cost[i, j] (+)= f(G_ij, qa_i, qb_j),   G = sum over K splits of the GEMM partial tiles
// f = identity            : cross_features_inner_product (activation_matching.py:14-28) and the
//                           weight-matching / normal-equation Grams
// f = -sqrt(max(.,0))     : cross_features_cdist (activation_matching.py:31-46) — ATen's
//                           _euclidean_dist: qa_i + qb_j - 2 G_ij, clamp_min(0), sqrt, then neg
// Splits are summed in a fixed order, so the result is deterministic.  HBM-bound: reads
// splits*M*N fp32, writes M*N.
#include "common.cuh"

namespace plb {

// Two shapes of the same epilogue.
//  * many K splits (SL = 8): 8 threads sum one output element's partials in parallel (fp64, fixed
//    order => deterministic) and combine through shared memory; block = 32 columns x 8 lanes.
//  * up to 48 splits: the kernel is a pure stream (read partial + read-modify-write cost), so
//    each thread owns kRows rows of one column and issues all its loads before using any of them
//    (one element per thread left the kernel latency-bound at 0.7 TB/s); block = 256 columns.
// Symmetric problems only computed the tiles touching the lower triangle: those are the only ones
// written (the caller mirrors the accumulator once at the end), so every access stays coalesced.
constexpr int kRows = 4;

template <typename OutT>
__device__ __forceinline__ void epilogue_store(double gs, int64_t i, int64_t j, const double *qa, const double *qb,
                                               int mode, OutT *cost, int64_t ldc, int accumulate, OutT old) {
  const float g = (float)gs;
  float v = g;
  if (mode == PLB_MODE_NEG_CDIST) {
    const float d2 = ((float)qa[i] + (float)qb[j]) - 2.0f * g;
    v = -sqrtf(fmaxf(d2, 0.f));
  }
  const OutT add = (sizeof(OutT) == 8 && mode == PLB_MODE_INNER) ? (OutT)gs : (OutT)v;
  cost[i * ldc + j] = accumulate ? (OutT)(old + add) : add;
}

template <typename OutT>
__global__ void __launch_bounds__(256) cross_finalize_stream_kernel(const float *__restrict__ partial, int splits,
                                                                    int64_t ld_m, int64_t ld_n, int64_t M, int64_t N,
                                                                    const double *__restrict__ qa,
                                                                    const double *__restrict__ qb, int mode,
                                                                    OutT *__restrict__ cost, int64_t ldc,
                                                                    int accumulate, int sym_bn) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * kRows;
  if (j >= N) return;
  const int64_t split_stride = ld_m * ld_n;
  bool live[kRows];
  const float *p[kRows];
  OutT old[kRows];
  double gs[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int64_t i = i0 + r;
    live[r] = i < M && !(sym_bn > 0 && (128 * (i / 128) + 127 < (int64_t)sym_bn * (j / sym_bn)));
    // dead rows read a valid dummy location (row i0's column j is inside the padded partial tile)
    p[r] = partial + (live[r] ? i : i0) * ld_n + j;
    old[r] = (live[r] && accumulate) ? cost[i * ldc + j] : (OutT)0;
    gs[r] = 0.0;
  }
  int s = 0;
  for (; s + 1 < splits; s += 2) {  // 2 splits x kRows rows = 8 independent loads in flight
    float v0[kRows], v1[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      v0[r] = p[r][(int64_t)s * split_stride];
      v1[r] = p[r][(int64_t)(s + 1) * split_stride];
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) gs[r] += (double)v0[r] + (double)v1[r];
  }
  if (s < splits) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) gs[r] += (double)p[r][(int64_t)s * split_stride];
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r)
    if (live[r]) epilogue_store<OutT>(gs[r], i0 + r, j, qa, qb, mode, cost, ldc, accumulate, old[r]);
}

template <typename OutT>
__global__ void __launch_bounds__(256) cross_finalize_reduce_kernel(const float *__restrict__ partial, int splits,
                                                                    int64_t ld_m, int64_t ld_n, int64_t M, int64_t N,
                                                                    const double *__restrict__ qa,
                                                                    const double *__restrict__ qb, int mode,
                                                                    OutT *__restrict__ cost, int64_t ldc,
                                                                    int accumulate, int sym_bn) {
  constexpr int SL = 8, COLS = 32;
  __shared__ double red[SL][COLS + 1];
  const int tx = threadIdx.x % COLS, ty = threadIdx.x / COLS;
  const int64_t j = (int64_t)blockIdx.x * COLS + tx;
  const int64_t i = blockIdx.y;
  const int64_t split_stride = ld_m * ld_n;
  const bool live = j < N && !(sym_bn > 0 && (128 * (i / 128) + 127 < (int64_t)sym_bn * (j / sym_bn)));
  double gs = 0.0;
  if (live) {
    const float *p = partial + i * ld_n + j;
    int s = ty;
    for (; s + 3 * SL < splits; s += 4 * SL) {  // 4 independent loads in flight per thread
      const float a = p[(int64_t)s * split_stride], b = p[(int64_t)(s + SL) * split_stride];
      const float c = p[(int64_t)(s + 2 * SL) * split_stride], d = p[(int64_t)(s + 3 * SL) * split_stride];
      gs += ((double)a + (double)b) + ((double)c + (double)d);
    }
    for (; s < splits; s += SL) gs += (double)p[(int64_t)s * split_stride];
  }
  red[ty][tx] = gs;
  __syncthreads();
  if (ty != 0 || !live) return;
#pragma unroll
  for (int k = 1; k < SL; ++k) gs += red[k][tx];
  epilogue_store<OutT>(gs, i, j, qa, qb, mode, cost, ldc, accumulate, accumulate ? cost[i * ldc + j] : (OutT)0);
}

// Correlation epilogue: Pearson coefficient of unit i of model A and unit j of model B over the K positions of
// the tap, from the cross-Gram and the rows' first and second moments — all in fp64 (the covariance is a
// difference of nearly equal numbers whenever the means are large against the spread, e.g. after a ReLU).
__global__ void __launch_bounds__(256) cross_finalize_corr_kernel(const float *__restrict__ partial, int splits,
                                                                  int64_t ld_m, int64_t ld_n, int64_t M, int64_t N,
                                                                  const double *__restrict__ qa,
                                                                  const double *__restrict__ qb,
                                                                  const double *__restrict__ sa,
                                                                  const double *__restrict__ sb, double inv_k,
                                                                  float *__restrict__ cost, int64_t ldc,
                                                                  int accumulate) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= N) return;
  const float *p = partial + i * ld_n + j;
  const int64_t split_stride = ld_m * ld_n;
  double g = 0.0;
  for (int s = 0; s < splits; ++s) g += (double)p[(int64_t)s * split_stride];
  const double ma = sa[i], mb = sb[j];
  const double cov = g - ma * mb * inv_k;
  const double va = qa[i] - ma * ma * inv_k, vb = qb[j] - mb * mb * inv_k;
  // a unit without variance on this batch (a dead ReLU channel) correlates with nothing: 0, not NaN,
  // so the assignment problem stays well defined
  const double eps_a = 1e-12 * qa[i], eps_b = 1e-12 * qb[j];
  float v = 0.f;
  if (va > eps_a && vb > eps_b) v = (float)(cov / sqrt(va * vb));
  cost[i * ldc + j] = accumulate ? cost[i * ldc + j] + v : v;
}

template <typename OutT>
static void launch_finalize(const float *partial, int splits, int64_t ld_m, int64_t ld_n, int64_t M, int64_t N,
                            const double *qa, const double *qb, int mode, OutT *cost, int64_t ldc, int accumulate,
                            int sym_bn, cudaStream_t s) {
  if (splits > 48) {
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)M);
    cross_finalize_reduce_kernel<OutT><<<grid, 256, 0, s>>>(partial, splits, ld_m, ld_n, M, N, qa, qb, mode, cost,
                                                            ldc, accumulate, sym_bn);
  } else {
    dim3 grid((unsigned)ceil_div(N, 256), (unsigned)ceil_div(M, kRows));
    cross_finalize_stream_kernel<OutT><<<grid, 256, 0, s>>>(partial, splits, ld_m, ld_n, M, N, qa, qb, mode, cost,
                                                            ldc, accumulate, sym_bn);
  }
}

}  // namespace plb

extern "C" int plb_cross_finalize(const float *partial, int32_t splits, int64_t ld_m, int64_t ld_n, int64_t M,
                                  int64_t N, const double *qa, const double *qb, int32_t mode, float *cost,
                                  double *cost64, int64_t ldc, int32_t accumulate, int32_t sym_bn, void *stream) {
  using namespace plb;
  PLB_REQUIRE(partial && (cost || cost64), PLB_EINVAL, "plb_cross_finalize: null pointer");
  PLB_REQUIRE(splits > 0 && M > 0 && N > 0 && M <= ld_m && N <= ld_n && ldc >= N, PLB_EINVAL,
              "plb_cross_finalize: bad geometry");
  PLB_REQUIRE(mode == PLB_MODE_INNER || (mode == PLB_MODE_NEG_CDIST && qa && qb), PLB_EINVAL,
              "plb_cross_finalize: mode needs row norms");
  PLB_REQUIRE(M <= 65535, PLB_ESIZE, "plb_cross_finalize: M too large");
  PLB_REQUIRE(sym_bn == 0 || (M == N && (sym_bn == 64 || sym_bn == 128 || sym_bn == 256)), PLB_EINVAL,
              "plb_cross_finalize: symmetric finalize needs a square problem and the GEMM's tile width");
  cudaStream_t s = (cudaStream_t)stream;
  if (cost64)
    launch_finalize<double>(partial, splits, ld_m, ld_n, M, N, qa, qb, mode, cost64, ldc, accumulate, sym_bn, s);
  else
    launch_finalize<float>(partial, splits, ld_m, ld_n, M, N, qa, qb, mode, cost, ldc, accumulate, sym_bn, s);
  return launch_status("cross_finalize");
}

namespace plb {

// One launch for ALL taps of a calibration batch.  A block owns a tile of one permutation group's cost matrix
// (thread = 4 consecutive columns x 4 rows; consecutive lanes = consecutive float4 of a partial row) and walks the
// group's taps in order: each tap's K-split partials are streamed as float4 with 8 independent 16-byte loads in
// flight per thread and summed in fp64 in a fixed order, the tap's statistic epilogue is applied, the taps' values
// are added up in registers and the cost entries are read and written ONCE per batch (the per-tap kernels did one
// read-modify-write per tap and cost 174 launches per ResNet-50 batch).  No atomics, no shared memory.
// Tile: 256 columns x 16 rows from 256 units up, 64 columns x 64 rows below (block = 256 threads either way).
constexpr int kGRows = 4, kGVec = 4;

__global__ void __launch_bounds__(256) cross_finalize_grouped_kernel(const PlbFinalizeTap *__restrict__ taps,
                                                                     const PlbFinalizeGroup *__restrict__ groups,
                                                                     int n_groups, int mode, int accumulate) {
  int lo = 0, hi = n_groups - 1;
  while (lo < hi) {  // last group whose block_begin <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (groups[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PlbFinalizeGroup g = groups[lo];
  const int local = (int)blockIdx.x - g.block_begin;
  const int cols = g.n >= 256 ? 256 : 64;                 // columns per block
  const int tpr = cols / kGVec;                           // threads per tile row: 64 or 16
  const int tile_rows = (256 / tpr) * kGRows;             // 16 or 64
  const int col_blocks = (g.n + cols - 1) / cols;
  const int tx = threadIdx.x % tpr, ty = threadIdx.x / tpr;
  const int64_t j = (int64_t)(local % col_blocks) * cols + tx * kGVec;
  const int64_t i0 = (int64_t)(local / col_blocks) * tile_rows + ty * kGRows;
  if (j >= g.n || i0 >= g.n) return;
  bool live[kGRows];
  float total[kGRows][kGVec];
#pragma unroll
  for (int r = 0; r < kGRows; ++r) {
    live[r] = i0 + r < g.n;
#pragma unroll
    for (int c = 0; c < kGVec; ++c) total[r][c] = 0.f;
  }
  for (int t = g.tap_begin; t < g.tap_end; ++t) {
    const PlbFinalizeTap tp = taps[t];
    const int64_t split_stride = tp.ld_m * tp.ld_n;  // ld_n is a multiple of 64: every float4 below is aligned and in range
    const float4 *p[kGRows];
    double gs[kGRows][kGVec];
#pragma unroll
    for (int r = 0; r < kGRows; ++r) {
      p[r] = reinterpret_cast<const float4 *>(tp.partial + (live[r] ? i0 + r : i0) * tp.ld_n + j);  // dead rows re-read row i0
#pragma unroll
      for (int c = 0; c < kGVec; ++c) gs[r][c] = 0.0;
    }
    const int64_t stride4 = split_stride / 4;
    int s2 = 0;
    for (; s2 + 1 < tp.splits; s2 += 2) {  // 2 splits x 4 rows = 8 independent 16-byte loads in flight
      float4 v0[kGRows], v1[kGRows];
#pragma unroll
      for (int r = 0; r < kGRows; ++r) {
        v0[r] = p[r][(int64_t)s2 * stride4];
        v1[r] = p[r][(int64_t)(s2 + 1) * stride4];
      }
#pragma unroll
      for (int r = 0; r < kGRows; ++r) {
        gs[r][0] += (double)v0[r].x + (double)v1[r].x;
        gs[r][1] += (double)v0[r].y + (double)v1[r].y;
        gs[r][2] += (double)v0[r].z + (double)v1[r].z;
        gs[r][3] += (double)v0[r].w + (double)v1[r].w;
      }
    }
    if (s2 < tp.splits) {
#pragma unroll
      for (int r = 0; r < kGRows; ++r) {
        const float4 v = p[r][(int64_t)s2 * stride4];
        gs[r][0] += (double)v.x;
        gs[r][1] += (double)v.y;
        gs[r][2] += (double)v.z;
        gs[r][3] += (double)v.w;
      }
    }
#pragma unroll
    for (int r = 0; r < kGRows; ++r) {
      if (!live[r]) continue;
      const int64_t i = i0 + r;
#pragma unroll
      for (int c = 0; c < kGVec; ++c) {
        if (j + c >= g.n) continue;
        float v;
        if (mode == PLB_MODE_NEG_CDIST) {
          const float gf = (float)gs[r][c];
          const float d2 = ((float)tp.qa[i] + (float)tp.qb[j + c]) - 2.0f * gf;
          v = -sqrtf(fmaxf(d2, 0.f));
        } else if (mode == PLB_MODE_CORR) {
          const double inv_k = 1.0 / (double)tp.K;
          const double ma = tp.sa[i], mb = tp.sb[j + c];
          const double cov = gs[r][c] - ma * mb * inv_k;
          const double va = tp.qa[i] - ma * ma * inv_k, vb = tp.qb[j + c] - mb * mb * inv_k;
          v = (va > 1e-12 * tp.qa[i] && vb > 1e-12 * tp.qb[j + c]) ? (float)(cov / sqrt(va * vb)) : 0.f;
        } else {
          v = (float)gs[r][c];
        }
        total[r][c] += v;
        // derived taps: the same statistic of the per-unit affine images s_a x_i + t_a, s_b y_j + t_b (an eval-mode
        // BatchNorm behind this tap) from the Gram entry, the row sums S and the sums of squares q, in fp64:
        //   <x', y'>  = s_a s_b G + s_a t_b S_a + t_a s_b S_b + K t_a t_b
        //   |x'-y'|^2 = s_a^2 q_a + s_b^2 q_b - 2 s_a s_b G + K (t_a - t_b)^2 + 2 (t_a - t_b)(s_a S_a - s_b S_b)
        //   corr(x', y') = sign(s_a s_b) corr(x, y)
        for (int a = 0; a < tp.n_affine; ++a) {
          const double *av = tp.affine + (int64_t)a * 4 * g.n;
          const double s_a = av[i], t_a = av[g.n + i], s_b = av[2 * g.n + j + c], t_b = av[3 * (int64_t)g.n + j + c];
          float w;
          if (mode == PLB_MODE_NEG_CDIST) {
            const double dt = t_a - t_b;
            const double d2 = s_a * s_a * tp.qa[i] + s_b * s_b * tp.qb[j + c] - 2.0 * s_a * s_b * gs[r][c] +
                              (double)tp.K * dt * dt + 2.0 * dt * (s_a * tp.sa[i] - s_b * tp.sb[j + c]);
            w = -sqrtf(fmaxf((float)d2, 0.f));
          } else if (mode == PLB_MODE_CORR) {
            const double sg = s_a * s_b;
            w = sg > 0.0 ? v : (sg < 0.0 ? -v : 0.f);
          } else {
            w = (float)(s_a * s_b * gs[r][c] + s_a * t_b * tp.sa[i] + t_a * s_b * tp.sb[j + c] +
                        (double)tp.K * t_a * t_b);
          }
          total[r][c] += w;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kGRows; ++r) {
    if (!live[r]) continue;
#pragma unroll
    for (int c = 0; c < kGVec; ++c) {
      if (j + c >= g.n) continue;
      float *cp = g.cost + (i0 + r) * g.ldc + j + c;
      *cp = accumulate ? *cp + total[r][c] : total[r][c];
    }
  }
}

}  // namespace plb

extern "C" int plb_cross_finalize_grouped(const PlbFinalizeTap *taps_dev, const PlbFinalizeGroup *groups_dev,
                                          int32_t n_groups, int32_t total_blocks, int32_t mode, int32_t accumulate,
                                          void *stream) {
  using namespace plb;
  PLB_REQUIRE(taps_dev && groups_dev && n_groups > 0 && total_blocks > 0, PLB_EINVAL,
              "plb_cross_finalize_grouped: empty table");
  PLB_REQUIRE(mode == PLB_MODE_INNER || mode == PLB_MODE_NEG_CDIST || mode == PLB_MODE_CORR, PLB_EINVAL,
              "plb_cross_finalize_grouped: unknown mode");
  cross_finalize_grouped_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>(taps_dev, groups_dev, n_groups, mode,
                                                                               accumulate);
  return launch_status("cross_finalize_grouped_kernel");
}

extern "C" int plb_cross_finalize_corr(const float *partial, int32_t splits, int64_t ld_m, int64_t ld_n, int64_t M,
                                       int64_t N, const double *qa, const double *qb, const double *sa,
                                       const double *sb, int64_t K, float *cost, int64_t ldc, int32_t accumulate,
                                       void *stream) {
  using namespace plb;
  PLB_REQUIRE(partial && cost && qa && qb && sa && sb, PLB_EINVAL, "plb_cross_finalize_corr: null pointer");
  PLB_REQUIRE(splits > 0 && M > 0 && N > 0 && M <= ld_m && N <= ld_n && ldc >= N && K > 0, PLB_EINVAL,
              "plb_cross_finalize_corr: bad geometry");
  PLB_REQUIRE(M <= 65535, PLB_ESIZE, "plb_cross_finalize_corr: M too large");
  dim3 grid((unsigned)ceil_div(N, 256), (unsigned)M);
  cross_finalize_corr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(partial, splits, ld_m, ld_n, M, N, qa, qb, sa, sb,
                                                                     1.0 / (double)K, cost, ldc, accumulate);
  return launch_status("cross_finalize_corr_kernel");
}
