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

// factor the nb x nb diagonal block at j0 (lower) in shared memory
__global__ void __launch_bounds__(1024) potrf_diag_kernel(double *G, int64_t n, int j0, int nb, int32_t *info) {
  __shared__ double a[NB][NB + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;  // a[tx][ty]: row tx, col ty
  if (tx < nb && ty < nb) a[tx][ty] = (ty <= tx) ? G[(int64_t)(j0 + tx) * n + j0 + ty] : 0.0;
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    if (tx == k && ty == k) {
      double piv = a[k][k];
      if (!(piv > 0.0)) {
        atomicCAS(info, 0, j0 + k + 1);
        piv = 1.0;
      }
      a[k][k] = sqrt(piv);
    }
    __syncthreads();
    if (ty == k && tx > k && tx < nb) a[tx][k] /= a[k][k];
    __syncthreads();
    if (tx > k && ty > k && ty <= tx && tx < nb) a[tx][ty] -= a[tx][k] * a[ty][k];
    __syncthreads();
  }
  if (tx < nb && ty <= tx) G[(int64_t)(j0 + tx) * n + j0 + ty] = a[tx][ty];
}

// rows below the diagonal block: A[i, j0:j0+nb] <- A[i, j0:j0+nb] * L_jj^{-T}
__global__ void __launch_bounds__(128) trsm_panel_kernel(double *G, int64_t n, int j0, int nb) {
  __shared__ double l[NB][NB + 1];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int r = e / nb, c = e % nb;
    l[r][c] = G[(int64_t)(j0 + r) * n + j0 + c];
  }
  __syncthreads();
  const int64_t i = (int64_t)j0 + nb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[NB];
  double *row = G + i * n + j0;
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = (c < nb) ? row[c] : 0.0;
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    if (c < nb) {
      double s = x[c];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        if (k < c) s -= x[k] * l[c][k];
      x[c] = s / l[c][c];
    }
  }
#pragma unroll
  for (int c = 0; c < NB; ++c)
    if (c < nb) row[c] = x[c];
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
  const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;
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
      for (int q = 0; q < 4; ++q) bv[q] = sb[c][tc + q];
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
      const int64_t i = i0 + tr + r, j = j0 + tc + q;
      if (i < M && j < N) C[i * ldc + j] -= acc[r][q];
    }
}

// diagonal-block triangular solve on the right-hand sides: one thread per rhs column
__global__ void __launch_bounds__(128) trsv_block_kernel(const double *__restrict__ G, int64_t n, int j0, int nb,
                                                         double *__restrict__ B, int64_t nrhs, int transposed) {
  __shared__ double l[NB][NB + 1];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int r = e / nb, c = e % nb;
    l[r][c] = G[(int64_t)(j0 + r) * n + j0 + c];
  }
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nrhs) return;
  double x[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) x[r] = (r < nb) ? B[(int64_t)(j0 + r) * nrhs + j] : 0.0;
  if (!transposed) {
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      if (r < nb) {
        double s = x[r];
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k < r) s -= l[r][k] * x[k];
        x[r] = s / l[r][r];
      }
    }
  } else {
#pragma unroll
    for (int rr = 0; rr < NB; ++rr) {
      const int r = NB - 1 - rr;
      if (r < nb) {
        double s = x[r];
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k > r && k < nb) s -= l[k][r] * x[k];
        x[r] = s / l[r][r];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NB; ++r)
    if (r < nb) B[(int64_t)(j0 + r) * nrhs + j] = x[r];
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
        trsv_block_kernel<<<(unsigned)ceil_div(nrhs, 128), 128, 0, s>>>(G, n, (int)j0, nb, B, nrhs, 0);
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
        trsv_block_kernel<<<(unsigned)ceil_div(nrhs, 128), 128, 0, s>>>(G, n, (int)j0, nb, B, nrhs, 1);
        // rows J0 .. j0 of this outer block: B[i, :] -= sum_c L[j0 + c, i] * X[j0 + c, :]
        update(B + J0 * nrhs, nrhs, j0 - J0, nrhs, G + j0 * n + J0, 1, n, B + j0 * nrhs, nrhs, 1, nb, 0);
      }
      // rows above the outer block
      update(B, nrhs, J0, nrhs, G + J0 * n, 1, n, B + J0 * nrhs, nrhs, 1, (int)W, 0);
    }
  }
  return launch_status("plb_chol_solve");
}
