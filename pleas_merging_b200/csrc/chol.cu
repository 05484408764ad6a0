// PLeaS closed-form solve: (G + ridge I) X = B by blocked right-looking Cholesky and two
// blocked triangular solves, fp64, on device.  Replaces the 401-step Adam loop of
// pleas/methods/pleas_merging.py:357-375 for the per-layer least-squares objective
// (:281-283); G = U^T U is the Gram of the im2col'd merged-layer input, B the gradient at the
// partial_merge init, so X is the minimum-ridge update of the layer weights.
//
// fp64 because normal equations square the condition number; the work (n^3/3 + 2 n^2 nrhs,
// n <= 4608) is a few ms per layer on B200's fp64 pipe and is L2/HBM-bound at this blocking
// (panel width 32, 64x64 trailing tiles), so plain SIMT DFMA is the right tool — no tensor
// cores here.
#include "common.cuh"

namespace plb {

constexpr int NB = 32;

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

// generic 64x64-tile rank-nb update  C[i, j] -= sum_c A(i, c) * B(c, j)
//   A(i, c) = A[i * ars + c * acs],  B(c, j) = B[c * brs + j * bcs],  C row-major ldc
// lower_only skips tiles strictly above the diagonal (trailing SYRK update).
__global__ void __launch_bounds__(256) rank_update_kernel(double *__restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                          const double *__restrict__ A, int64_t ars, int64_t acs,
                                                          const double *__restrict__ B, int64_t brs, int64_t bcs,
                                                          int nb, int lower_only) {
  if (lower_only && blockIdx.x > blockIdx.y) return;
  __shared__ double sa[64][NB + 1];
  __shared__ double sb[NB][64 + 1];
  const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
  const int tid = threadIdx.x;
  for (int e = tid; e < 64 * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    sa[r][c] = (i0 + r < M && c < nb) ? A[(i0 + r) * ars + (int64_t)c * acs] : 0.0;
  }
  for (int e = tid; e < NB * 64; e += 256) {
    const int c = e / 64, j = e % 64;
    sb[c][j] = (j0 + j < N && c < nb) ? B[(int64_t)c * brs + (j0 + j) * bcs] : 0.0;
  }
  __syncthreads();
  const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;
  double acc[4][4] = {};
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
  // ---- factor G = L L^T (lower, in place)
  for (int64_t j0 = 0; j0 < n; j0 += NB) {
    const int nb = (int)((n - j0 < NB) ? (n - j0) : NB);
    potrf_diag_kernel<<<1, dim3(NB, NB), 0, s>>>(G, n, (int)j0, nb, info);
    const int64_t m = n - j0 - nb;
    if (m > 0) {
      trsm_panel_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(G, n, (int)j0, nb);
      double *panel = G + (j0 + nb) * n + j0;
      const unsigned t = (unsigned)ceil_div(m, 64);
      // trailing[i, k] -= sum_c panel[i, c] * panel[k, c]
      rank_update_kernel<<<dim3(t, t), 256, 0, s>>>(G + (j0 + nb) * n + (j0 + nb), n, m, m, panel, n, 1, panel, 1, n,
                                                    nb, 1);
    }
  }
  if (nrhs > 0) {
    // ---- forward: L Z = B
    for (int64_t j0 = 0; j0 < n; j0 += NB) {
      const int nb = (int)((n - j0 < NB) ? (n - j0) : NB);
      trsv_block_kernel<<<(unsigned)ceil_div(nrhs, 128), 128, 0, s>>>(G, n, (int)j0, nb, B, nrhs, 0);
      const int64_t m = n - j0 - nb;
      if (m > 0)
        rank_update_kernel<<<dim3((unsigned)ceil_div(nrhs, 64), (unsigned)ceil_div(m, 64)), 256, 0, s>>>(
            B + (j0 + nb) * nrhs, nrhs, m, nrhs, G + (j0 + nb) * n + j0, n, 1, B + j0 * nrhs, nrhs, 1, nb, 0);
    }
    // ---- backward: L^T X = Z
    for (int64_t j0 = ((n - 1) / NB) * NB; j0 >= 0; j0 -= NB) {
      const int nb = (int)((n - j0 < NB) ? (n - j0) : NB);
      trsv_block_kernel<<<(unsigned)ceil_div(nrhs, 128), 128, 0, s>>>(G, n, (int)j0, nb, B, nrhs, 1);
      if (j0 > 0)  // rows above: B[i, :] -= sum_c L[j0 + c, i] * X[j0 + c, :]
        rank_update_kernel<<<dim3((unsigned)ceil_div(nrhs, 64), (unsigned)ceil_div(j0, 64)), 256, 0, s>>>(
            B, nrhs, j0, nrhs, G + j0 * n, 1, n, B + j0 * nrhs, nrhs, 1, nb, 0);
    }
  }
  return launch_status("plb_chol_solve");
}
