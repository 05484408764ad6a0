// PLeaS closed-form solve: (G + ridge I) X = B by blocked right-looking Cholesky and two
// blocked triangular solves, fp64, on device.  Replaces the 401-step Adam loop of
// pleas/methods/pleas_merging.py:357-375 for the per-layer least-squares objective
// (:281-283); G = U^T U is the Gram of the im2col'd merged-layer input, B the gradient at the
// partial_merge init, so X is the minimum-ridge update of the layer weights.
//
// fp64 because normal equations square the condition number; the work (n^3/3 + 2 n^2 nrhs,
// n <= 4608) is a few ms per layer on B200's fp64 pipe and is L2/HBM-bound at this blocking
// (panel width 32, 64x64 trailing tiles), so plain SIMT DFMA is the right tool — no tensor
// cores here.  Two-level blocking (inner steps of 32 inside outer panels of 128) keeps the
// trailing matrix's read-modify-write traffic and the launch count down.
#include "common.cuh"

namespace plb {

constexpr int NB = 32;   // inner step: diagonal blocks live in shared memory / registers
constexpr int OB = 128;  // outer panel: one trailing update per OB columns

__global__ void add_ridge_kernel(double *G, int64_t n, double ridge) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) G[i * n + i] += ridge;
}

// Factor the nb x nb diagonal block at j0 (lower, in place) and store the strictly-lower part of
// inv(L_jj) transposed into the block's strictly-upper triangle (scratch space nobody reads; the
// inverse's diagonal is 1 / L_ii).  The panel and right-hand-side solves then become small
// matrix products without dependent chains or divisions.
// 32 x 32 threads, thread (i, j) owns entry [i][j]; one barrier per step:
//   factor:  step k applies the rank-1 update a[i][j] -= a[i][k] a[j][k] / a[k][k] to the trailing
//            entries directly from the UNSCALED column k, and stores L[i][k] = a[i][k] / sqrt(a[k][k])
//            on the side (column k is never touched again);
//   inverse: step k finalises row k of X = inv(L) (X[k][j] = (I - S)[k][j] / L[k][k]) and pushes
//            it into the running sums S[i][j] += L[i][k] X[k][j] of the rows below.
__global__ void __launch_bounds__(1024) potrf_diag_kernel(double *G, int64_t n, int j0, int nb, int32_t *info) {
  __shared__ double a[NB][NB + 1];
  __shared__ double l[NB][NB + 1];
  __shared__ double x[NB][NB + 1];
  const int i = threadIdx.y, j = threadIdx.x;
  a[i][j] = (i < nb && j <= i) ? G[(int64_t)(j0 + i) * n + j0 + j] : (i == j ? 1.0 : 0.0);  // identity padding
  l[i][j] = 0.0;
  __syncthreads();
  for (int k = 0; k < NB; ++k) {
    double piv = a[k][k];
    if (!(piv > 0.0)) {
      if (i == 0 && j == 0 && k < nb) atomicCAS(info, 0, j0 + k + 1);
      piv = 1.0;
    }
    const double aik = a[i][k], ajk = a[j][k];
    if (j == k && i >= k) l[i][k] = aik / sqrt(piv);
    if (i > k && j > k && j <= i) a[i][j] -= aik * ajk / piv;  // never touches column k: no barrier before
    __syncthreads();
  }
  double s = 0.0;  // S[i][j]
  for (int k = 0; k < NB; ++k) {
    if (i == k) x[k][j] = (j <= k) ? (((j == k) ? 1.0 : 0.0) - s) / l[k][k] : 0.0;
    __syncthreads();
    if (i > k) s = fma(l[i][k], x[k][j], s);
  }
  __syncthreads();
  if (i < nb && j < nb) G[(int64_t)(j0 + i) * n + j0 + j] = (j <= i) ? l[i][j] : x[j][i];  // upper: inv(L)^T
}

// loads inv(L_jj) (lower triangular incl. diagonal) from the layout potrf_diag_kernel left behind
__device__ __forceinline__ void load_inverse_block(double (*li)[NB + 1], const double *G, int64_t n, int j0, int nb) {
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    double v = 0.0;
    if (r < nb && c < nb) {
      if (c < r) v = G[(int64_t)(j0 + c) * n + j0 + r];
      else if (c == r) v = 1.0 / G[(int64_t)(j0 + r) * n + j0 + r];
    }
    li[r][c] = v;
  }
}

// rows below the diagonal block: A[i, j0:j0+nb] <- A[i, j0:j0+nb] * inv(L_jj)^T
// (thread = one row; out[c] = sum_{k <= c} x[k] * inv[c][k]: independent dot products)
__global__ void __launch_bounds__(128) trsm_panel_kernel(double *G, int64_t n, int j0, int nb) {
  __shared__ double li[NB][NB + 1];
  load_inverse_block(li, G, n, j0, nb);
  __syncthreads();
  const int64_t i = (int64_t)j0 + nb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[NB];
  double *row = G + i * n + j0;
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = (c < nb) ? row[c] : 0.0;
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < NB; k += 2) {
      if (k <= c) s0 = fma(x[k], li[c][k], s0);
      if (k + 1 <= c) s1 = fma(x[k + 1], li[c][k + 1], s1);
    }
    if (c < nb) row[c] = s0 + s1;
  }
}

// generic 64x64-tile rank-kdim update  C[i, j] -= sum_c A(i, c) * B(c, j),  c < kdim
//   A(i, c) = A[i * ars + c * acs],  B(c, j) = B[c * brs + j * bcs],  C row-major ldc
// kdim is consumed in chunks of NB staged through shared memory (global loads walk the unit-stride
// index of each operand fastest, so both the panel and its transpose are read coalesced); C is
// read and written once per launch.  lower_only skips tiles strictly above the diagonal
// (trailing SYRK update).
__global__ void __launch_bounds__(256) rank_update_kernel(double *__restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                          const double *__restrict__ A, int64_t ars, int64_t acs,
                                                          const double *__restrict__ B, int64_t brs, int64_t bcs,
                                                          int kdim, int lower_only) {
  if (lower_only && blockIdx.x > blockIdx.y) return;
  __shared__ double sa[64][NB + 1];
  __shared__ double sb[NB][64 + 1];
  const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
  const int tid = threadIdx.x;
  // thread micro-tile: rows tr..tr+3, columns tx + 16 q — the 16 lanes of a half-warp read 16
  // consecutive doubles of sb (conflict-free) and both half-warps share them; sa reads broadcast
  // (measured: 12.7 -> 9.6 ms of updates in a K = 4608 solve; a 128x64 / 8x4 register-prefetch
  // variant was slower — fewer resident warps to hide the fp64 latency)
  const int tr = (tid / 16) * 4, tx = tid % 16;
  double acc[4][4] = {};
  for (int c0 = 0; c0 < kdim; c0 += NB) {
    const int nb = min(NB, kdim - c0);
    if (c0) __syncthreads();
    for (int e = tid; e < 64 * NB; e += 256) {
      int r, c;
      if (acs == 1) { r = e / NB; c = e % NB; } else { c = e / 64; r = e % 64; }
      sa[r][c] = (i0 + r < M && c < nb) ? A[(i0 + r) * ars + (int64_t)(c0 + c) * acs] : 0.0;
    }
    for (int e = tid; e < NB * 64; e += 256) {
      int c, j;
      if (brs == 1) { j = e / NB; c = e % NB; } else { c = e / 64; j = e % 64; }
      sb[c][j] = (j0 + j < N && c < nb) ? B[(int64_t)(c0 + c) * brs + (j0 + j) * bcs] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < NB; ++c) {
      double av[4], bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) av[r] = sa[tr + r][c];
#pragma unroll
      for (int q = 0; q < 4; ++q) bv[q] = sb[c][tx + 16 * q];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = fma(av[r], bv[q], acc[r][q]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t i = i0 + tr + r, j = j0 + tx + 16 * q;
      if (i < M && j < N) C[i * ldc + j] -= acc[r][q];
    }
}

// diagonal-block triangular solve on the right-hand sides as a product with the stored inverse:
//   forward  X = inv(L_jj) B_blk,   backward  X = inv(L_jj)^T B_blk.
// block = 32 rhs columns x 8 row-quads: thread (tx, ty) produces rows 4 ty .. 4 ty + 3 of column tx
__global__ void __launch_bounds__(256) trsv_block_kernel(const double *__restrict__ G, int64_t n, int j0, int nb,
                                                         double *__restrict__ B, int64_t nrhs, int transposed) {
  __shared__ double li[NB][NB + 1];
  __shared__ double xb[NB][32 + 1];
  load_inverse_block(li, G, n, j0, nb);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + tx;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = ty * 4 + q;
    xb[r][tx] = (r < nb && j < nrhs) ? B[(int64_t)(j0 + r) * nrhs + j] : 0.0;
  }
  __syncthreads();
  double out[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
  for (int k = 0; k < NB; ++k) {
    const double xv = xb[k][tx];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty * 4 + q;
      out[q] = fma(transposed ? li[k][r] : li[r][k], xv, out[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = ty * 4 + q;
    if (r < nb && j < nrhs) B[(int64_t)(j0 + r) * nrhs + j] = out[q];
  }
}

}  // namespace plb

extern "C" int plb_chol_solve(double *G, int64_t n, double *B, int64_t nrhs, double ridge, int32_t *info,
                              void *stream) {
  using namespace plb;
  PLB_REQUIRE(G && info && n > 0, PLB_EINVAL, "plb_chol_solve: bad arguments");
  PLB_REQUIRE(nrhs == 0 || B, PLB_EINVAL, "plb_chol_solve: null right-hand side");
  PLB_REQUIRE(n < ((int64_t)1 << 30), PLB_ESIZE, "plb_chol_solve: n too large");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) {
    set_error("plb_chol_solve: memset: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (ridge != 0.0) add_ridge_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(G, n, ridge);
  auto update = [&](double *C, int64_t ldc, int64_t M, int64_t N, const double *A, int64_t ars, int64_t acs,
                    const double *B, int64_t brs, int64_t bcs, int kdim, int lower_only) {
    if (M > 0 && N > 0 && kdim > 0)
      rank_update_kernel<<<dim3((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64)), 256, 0, s>>>(
          C, ldc, M, N, A, ars, acs, B, brs, bcs, kdim, lower_only);
  };
  // Two-level blocking: NB-wide steps (diagonal factor / triangular solve in shared memory) only
  // update the columns of their own OB-wide outer panel; the trailing matrix is updated once per outer
  // panel with k = OB, which cuts its read-modify-write traffic and launch count by OB / NB.
  // ---- factor G = L L^T (lower, in place; entries above the diagonal are scratch)
  for (int64_t J0 = 0; J0 < n; J0 += OB) {
    const int64_t W = (n - J0 < OB) ? (n - J0) : OB;
    for (int64_t j0 = J0; j0 < J0 + W; j0 += NB) {
      const int nb = (int)((J0 + W - j0 < NB) ? (J0 + W - j0) : NB);
      potrf_diag_kernel<<<1, dim3(NB, NB), 0, s>>>(G, n, (int)j0, nb, info);
      const int64_t m = n - j0 - nb;
      if (m > 0) {
        trsm_panel_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(G, n, (int)j0, nb);
        // columns j0+nb .. J0+W of the rows below: G[i, k] -= sum_c panel[i, c] * panel[k, c]
        double *panel = G + (j0 + nb) * n + j0;
        update(G + (j0 + nb) * n + (j0 + nb), n, m, J0 + W - j0 - nb, panel, n, 1, panel, 1, n, nb, 0);
      }
    }
    const int64_t m = n - J0 - W;
    double *panel = G + (J0 + W) * n + J0;
    update(G + (J0 + W) * n + (J0 + W), n, m, m, panel, n, 1, panel, 1, n, (int)W, 1);
  }
  if (nrhs > 0) {
    // ---- forward: L Z = B
    for (int64_t J0 = 0; J0 < n; J0 += OB) {
      const int64_t W = (n - J0 < OB) ? (n - J0) : OB;
      for (int64_t j0 = J0; j0 < J0 + W; j0 += NB) {
        const int nb = (int)((J0 + W - j0 < NB) ? (J0 + W - j0) : NB);
        trsv_block_kernel<<<(unsigned)ceil_div(nrhs, 32), 256, 0, s>>>(G, n, (int)j0, nb, B, nrhs, 0);
        update(B + (j0 + nb) * nrhs, nrhs, J0 + W - j0 - nb, nrhs, G + (j0 + nb) * n + j0, n, 1, B + j0 * nrhs, nrhs, 1,
               nb, 0);
      }
      update(B + (J0 + W) * nrhs, nrhs, n - J0 - W, nrhs, G + (J0 + W) * n + J0, n, 1, B + J0 * nrhs, nrhs, 1, (int)W,
             0);
    }
    // ---- backward: L^T X = Z
    for (int64_t J0 = ((n - 1) / OB) * OB; J0 >= 0; J0 -= OB) {
      const int64_t W = (n - J0 < OB) ? (n - J0) : OB;
      for (int64_t j0 = J0 + ((W - 1) / NB) * NB; j0 >= J0; j0 -= NB) {
        const int nb = (int)((J0 + W - j0 < NB) ? (J0 + W - j0) : NB);
        trsv_block_kernel<<<(unsigned)ceil_div(nrhs, 32), 256, 0, s>>>(G, n, (int)j0, nb, B, nrhs, 1);
        // rows J0 .. j0 of this outer block: B[i, :] -= sum_c L[j0 + c, i] * X[j0 + c, :]
        update(B + J0 * nrhs, nrhs, j0 - J0, nrhs, G + j0 * n + J0, 1, n, B + j0 * nrhs, nrhs, 1, nb, 0);
      }
      // rows above the outer block
      update(B, nrhs, J0, nrhs, G + J0 * n, 1, n, B + J0 * nrhs, nrhs, 1, (int)W, 0);
    }
  }
  return launch_status("plb_chol_solve");
}
