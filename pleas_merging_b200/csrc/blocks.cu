// partial_merge building blocks and the small index kernels of weight matching — all
// HBM/latency-bound integer and gather work (pleas/methods/partial_matching.py:47-176,
// pleas/core/utils.py:233-246, pleas/methods/weight_matching.py:80-88).
#include <math.h>

#include "common.cuh"

namespace plb {

// ------------------------------------------------------------------------------- get_blocks
// One CTA per group.  Matched costs -> bitonic sort in shared memory -> torch.quantile
// threshold (ATen quantile_compute: rank = q*(n-1) in fp32, two-sided lerp) -> mask ->
// order-preserving compaction (block scan).
constexpr int kMaxBlocksN = 4096;

__global__ void __launch_bounds__(1024, 1) get_blocks_kernel(const float *__restrict__ cost, int64_t ldc,
                                                             const int64_t *__restrict__ perm, int n, float ratio,
                                                             int identity, int64_t *__restrict__ qm,
                                                             int64_t *__restrict__ pm, int64_t *__restrict__ qs,
                                                             int64_t *__restrict__ ps, int32_t *__restrict__ counts) {
  __shared__ float vals[kMaxBlocksN];
  __shared__ float sorted[kMaxBlocksN];
  __shared__ int warp_sums[32];
  __shared__ float thr_s;
  const int tid = threadIdx.x;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = tid; i < np2; i += blockDim.x) {
    float c = INFINITY;
    if (i < n) {
      const int64_t p = identity ? (int64_t)i : perm[i];
      c = cost[(int64_t)i * ldc + p];
      vals[i] = c;
    }
    sorted[i] = c;
  }
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = sorted[i], b = sorted[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            sorted[i] = b;
            sorted[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    const float rank = __fmul_rn(ratio, (float)(n - 1));
    const float below = floorf(rank);
    const float w = __fsub_rn(rank, below);
    const float a = sorted[(int)below], b = sorted[(int)ceilf(rank)];
    const float diff = __fsub_rn(b, a);
    // ATen lerp (Lerp.h), compiled with fused multiply-add like torch's own kernels
    thr_s = (w < 0.5f) ? fmaf(w, diff, a) : fmaf(-diff, __fsub_rn(1.0f, w), b);
  }
  __syncthreads();
  const float thr = thr_s;
  // each thread owns 4 consecutive units so the compaction preserves index order
  const int base = tid * 4;
  int m[4], cnt = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = base + e;
    m[e] = (i < n) && (vals[i] >= thr);
    cnt += m[e];
  }
  int incl = cnt;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = warp_sums[lane];
    int iv = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, iv, d);
      if (lane >= d) iv += t;
    }
    warp_sums[lane] = iv - v;  // exclusive
    if (lane == 31) counts[0] = iv;
  }
  __syncthreads();
  int pos_m = warp_sums[warp] + incl - cnt;  // merged units before `base`
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = base + e;
    if (i >= n) break;
    const int64_t p = identity ? (int64_t)i : perm[i];
    if (m[e]) {
      qm[pos_m] = i;
      pm[pos_m] = p;
      ++pos_m;
    } else {
      const int pos_s = i - pos_m;
      qs[pos_s] = i;
      ps[pos_s] = p;
    }
  }
}

// ------------------------------------------------------------------------------- block merge
struct AxisBlocks {
  const int64_t *b1, *b2, *b1c, *b2c;
  int64_t nm, ms;  // merged, separate counts
};
// returns class (0 merged, 1 model-1 separate, 2 model-2 separate) and the source indices
__device__ __forceinline__ int classify(const AxisBlocks &ab, int64_t idx, int64_t &s1, int64_t &s2) {
  if (ab.b1 == nullptr) {
    s1 = s2 = idx;
    return 0;
  }
  if (idx < ab.nm) {
    s1 = ab.b1[idx];
    s2 = ab.b2[idx];
    return 0;
  }
  if (idx < ab.nm + ab.ms) {
    s1 = ab.b1c[idx - ab.nm];
    s2 = -1;
    return 1;
  }
  s1 = -1;
  s2 = ab.b2c[idx - ab.nm - ab.ms];
  return 2;
}

__global__ void __launch_bounds__(256) block_merge_kernel(const float *__restrict__ w1, const float *__restrict__ w2,
                                                          int64_t I, int64_t R, AxisBlocks bo, AxisBlocks bi,
                                                          int64_t Iout, int in_axis_only, float *__restrict__ out,
                                                          int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t rr = e % R;
  const int64_t t = e / R;
  const int64_t ip = t % Iout, op = t / Iout;
  int64_t o1, o2, i1, i2;
  const int co = classify(bo, op, o1, o2);
  const int ci = classify(bi, ip, i1, i2);
  float v = 0.f;
  if (co == 0 && ci == 0) {
    v = (w1[(o1 * I + i1) * R + rr] + w2[(o2 * I + i2) * R + rr]) / 2.0f;
  } else if (co == 0) {  // separate input feeding a merged output: activations are averaged
    const float s = (ci == 1) ? w1[(o1 * I + i1) * R + rr] : w2[(o2 * I + i2) * R + rr];
    v = in_axis_only ? s : s / 2.0f;
  } else if (ci == 0) {  // merged input feeding a separate output
    v = (co == 1) ? w1[(o1 * I + i1) * R + rr] : w2[(o2 * I + i2) * R + rr];
  } else if (co == ci) {
    v = (co == 1) ? w1[(o1 * I + i1) * R + rr] : w2[(o2 * I + i2) * R + rr];
  }
  out[e] = v;
}

__global__ void __launch_bounds__(256) gather_axis_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                          int64_t n, int64_t inner, const int64_t *__restrict__ P,
                                                          int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t i = e % inner;
  const int64_t t = e / inner;
  const int64_t p = t % n, o = t / n;
  out[e] = in[(o * n + P[p]) * inner + i];
}

__global__ void compose_perm_kernel(const int64_t *__restrict__ a, const int64_t *__restrict__ b,
                                    int64_t *__restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[b[i]];
}

__global__ void __launch_bounds__(256) wm_progress_kernel(const float *__restrict__ A, int64_t ld,
                                                          const int64_t *__restrict__ P, int n, int32_t *flag,
                                                          double *gain) {
  __shared__ double red[2][8];
  double o = 0.0, nw = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    o += (double)A[(int64_t)i * ld + i];
    nw += (double)A[(int64_t)i * ld + P[i]];
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    o += __shfl_xor_sync(0xffffffffu, o, m);
    nw += __shfl_xor_sync(0xffffffffu, nw, m);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = o;
    red[1][threadIdx.x >> 5] = nw;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double so = 0.0, sn = 0.0;
    for (int wq = 0; wq < 8; ++wq) {
      so += red[0][wq];
      sn += red[1][wq];
    }
    // the reference compares torch fp32 sums (weight_matching.py:80-81: newL > oldL + 1e-12 on float32
    // tensors): a gain below one ulp of the sum is NOT progress.  The sums are formed in fp64 and rounded
    // to fp32 once, so GEMM-level noise on near-tied groups cannot trigger extra sweeps.
    const float so32 = (float)so, sn32 = (float)sn;
    if (sn32 > so32 + 1e-12f) *flag = 1;
    if (gain) *gain = sn - so;
  }
}

}  // namespace plb

extern "C" int plb_get_blocks(const float *cost, int64_t ldc, const int64_t *perm, int32_t n, float ratio,
                              int32_t identity, int64_t *q_merged, int64_t *p_merged, int64_t *q_sep,
                              int64_t *p_sep, int32_t *counts, void *stream) {
  using namespace plb;
  PLB_REQUIRE(cost && (perm || identity) && q_merged && p_merged && q_sep && p_sep && counts, PLB_EINVAL,
              "plb_get_blocks: null pointer");
  PLB_REQUIRE(n > 0 && n <= kMaxBlocksN, PLB_ESIZE, "plb_get_blocks: n must be in [1, 4096]");
  PLB_REQUIRE(ratio >= 0.f && ratio <= 1.f, PLB_EINVAL, "plb_get_blocks: ratio must be in [0, 1]");
  get_blocks_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(cost, ldc, perm, n, ratio, identity, q_merged, p_merged,
                                                         q_sep, p_sep, counts);
  return launch_status("get_blocks_kernel");
}

extern "C" int plb_block_merge(const float *w1, const float *w2, int64_t O, int64_t I, int64_t R, const int64_t *bo1,
                               const int64_t *bo2, const int64_t *bo1c, const int64_t *bo2c, int64_t no, int64_t mo,
                               const int64_t *bi1, const int64_t *bi2, const int64_t *bi1c, const int64_t *bi2c,
                               int64_t ni, int64_t mi, int32_t in_axis_only, float *out, void *stream) {
  using namespace plb;
  PLB_REQUIRE(w1 && w2 && out, PLB_EINVAL, "plb_block_merge: null pointer");
  PLB_REQUIRE(O > 0 && I > 0 && R > 0, PLB_EINVAL, "plb_block_merge: empty tensor");
  AxisBlocks bo{bo1, bo2, bo1c, bo2c, no, mo}, bi{bi1, bi2, bi1c, bi2c, ni, mi};
  const int64_t Oout = bo1 ? no + 2 * mo : O;
  const int64_t Iout = bi1 ? ni + 2 * mi : I;
  PLB_REQUIRE(!bo1 || (bo2 && (mo == 0 || (bo1c && bo2c))), PLB_EINVAL, "plb_block_merge: incomplete output blocks");
  PLB_REQUIRE(!bi1 || (bi2 && (mi == 0 || (bi1c && bi2c))), PLB_EINVAL, "plb_block_merge: incomplete input blocks");
  const int64_t total = Oout * Iout * R;
  if (total == 0) return PLB_OK;
  block_merge_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w1, w2, I, R, bo, bi, Iout,
                                                                                      in_axis_only, out, total);
  return launch_status("block_merge_kernel");
}

extern "C" int plb_gather_axis(const float *in, float *out, int64_t outer, int64_t n, int64_t inner, const int64_t *P,
                               void *stream) {
  using namespace plb;
  PLB_REQUIRE(in && out && P, PLB_EINVAL, "plb_gather_axis: null pointer");
  PLB_REQUIRE(in != out, PLB_EINVAL, "plb_gather_axis: in-place gather is not supported");
  const int64_t total = outer * n * inner;
  if (total <= 0) return PLB_OK;
  gather_axis_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, inner, P, total);
  return launch_status("gather_axis_kernel");
}

extern "C" int plb_compose_perm(const int64_t *a, const int64_t *b, int64_t *out, int64_t n, void *stream) {
  using namespace plb;
  PLB_REQUIRE(a && b && out && out != a, PLB_EINVAL, "plb_compose_perm: bad pointers");
  if (n <= 0) return PLB_OK;
  compose_perm_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  return launch_status("compose_perm_kernel");
}

extern "C" int plb_wm_progress(const float *A, int64_t ld, const int64_t *P, int32_t n, int32_t *flag, double *gain,
                               void *stream) {
  using namespace plb;
  PLB_REQUIRE(A && P && flag && n > 0, PLB_EINVAL, "plb_wm_progress: bad arguments");
  wm_progress_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(A, ld, P, n, flag, gain);
  return launch_status("wm_progress_kernel");
}
