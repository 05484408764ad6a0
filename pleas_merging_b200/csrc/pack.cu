// Operand packing for the 3xTF32 tcgen05 GEMM (gemm.cu): one pass over the source tensor
// that (a) replaces the reference's `movedim(x, a, 0).reshape(C, -1)` transposed copy
// (activation_matching.py:26-27, 44-45), (b) splits every fp32 value into tf32 hi + tf32 lo,
// (c) writes both in tcgen05's K-major core-matrix order and (d) accumulates the per-row
// sum of squares / sum the -cdist and correlation epilogues need, so the activation is read
// from HBM exactly once.
//
// Thread mapping: a warp owns one (8 rows x 16 k) panel per iteration; lane = 8*jj + r reads
// 4 consecutive k of row r (jj-th 16-byte chunk) and writes them as one float4, so a warp's
// store is one contiguous 512-byte panel per plane and its load is 8 rows x 64 contiguous
// bytes; the 8 warps of a block walk consecutive k-blocks, i.e. 512 contiguous bytes per row.
#include "common.cuh"

namespace plb {

constexpr int kKbPerBlock = 64;

__device__ __forceinline__ void split_store(const float (&v)[4], float *hi, float *lo, int64_t off) {
  float4 h, l;
  h.x = to_tf32(v[0]); l.x = to_tf32(v[0] - h.x);
  h.y = to_tf32(v[1]); l.y = to_tf32(v[1] - h.y);
  h.z = to_tf32(v[2]); l.z = to_tf32(v[2] - h.z);
  h.w = to_tf32(v[3]); l.w = to_tf32(v[3] - h.w);
  *reinterpret_cast<float4 *>(hi + off) = h;
  *reinterpret_cast<float4 *>(lo + off) = l;
}

// block-level reduction of the per-thread row statistics and one fp64 atomic per row per block
__device__ __forceinline__ void reduce_row_stats(float s2, float s1, int row, int rows, double *sumsq, double *sum) {
  __shared__ float red[2][8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  if (lane < 8) {
    red[0][warp][lane] = s2;
    red[1][warp][lane] = s1;
  }
  __syncthreads();
  if (threadIdx.x < 8 && row < rows) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) {
      a += (double)red[0][wq][threadIdx.x];
      b += (double)red[1][wq][threadIdx.x];
    }
    if (sumsq) atomicAdd(sumsq + row, a);
    if (sum) atomicAdd(sum + row, b);
  }
}

template <bool VEC4>
__device__ __forceinline__ void pack_split_body(const float *__restrict__ x, int64_t src_rows, uint32_t inner,
                                                uint32_t K, const int64_t *__restrict__ row_index, int rows,
                                                float *__restrict__ hi, float *__restrict__ lo, int row_groups,
                                                int kb_offset, int k_blocks, double *__restrict__ sumsq,
                                                double *__restrict__ sum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane & 7, jj = lane >> 3;
  const int g = blockIdx.y;
  const int row = g * 8 + r;
  const bool row_ok = row < rows;
  const int64_t src_row = row_ok ? (row_index ? row_index[row] : (int64_t)row) : 0;
  const int kb_end = min(k_blocks, (int)(blockIdx.x + 1) * kKbPerBlock);
  float s2 = 0.f, s1 = 0.f;
#pragma unroll 2
  for (int kb = blockIdx.x * kKbPerBlock + warp; kb < kb_end; kb += 8) {
    const uint32_t k = (uint32_t)kb * kPackK + jj * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (row_ok) {
      if (VEC4) {
        if (k < K) {
          const uint32_t o = k / inner, i = k - o * inner;
          const float4 t = __ldg(reinterpret_cast<const float4 *>(x + ((int64_t)o * src_rows + src_row) * inner + i));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t ke = k + e;
          if (ke < K) {
            const uint32_t o = ke / inner, i = ke - o * inner;
            v[e] = __ldg(x + ((int64_t)o * src_rows + src_row) * inner + i);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s2 = fmaf(v[e], v[e], s2);
      s1 += v[e];
    }
    split_store(v, hi, lo, panel_offset(kb_offset + kb, g, row_groups) + (jj * 8 + r) * 4);
  }
  if (sumsq || sum) reduce_row_stats(s2, s1, g * 8 + (int)threadIdx.x, rows, sumsq, sum);
}

template <bool VEC4>
__global__ void __launch_bounds__(256) pack_split_kernel(const float *__restrict__ x, int64_t src_rows, uint32_t inner,
                                                         uint32_t K, const int64_t *__restrict__ row_index, int rows,
                                                         float *__restrict__ hi, float *__restrict__ lo,
                                                         int row_groups, int kb_offset, int k_blocks,
                                                         double *__restrict__ sumsq, double *__restrict__ sum) {
  pack_split_body<VEC4>(x, src_rows, inner, K, row_index, rows, hi, lo, row_groups, kb_offset, k_blocks, sumsq, sum);
}

// Both operands of one tap (same geometry) in ONE launch: blockIdx.z selects the operand.  Half the
// launches and twice the CTAs per launch for the many small taps (C = 256..2048 at 14x14 / 7x7), whose
// single-operand packs are ~10 us launches that never fill the machine (ncu: 1.3-3.7 TB/s vs 5.4-6.5
// TB/s for the large ones, profiles/launches_r01c.csv.gz).
struct PackPair {
  const float *x[2];
  float *hi[2], *lo[2];
  double *sumsq[2];
  double *sum[2];
};

template <bool VEC4>
__global__ void __launch_bounds__(256) pack_split_pair_kernel(PackPair pp, int64_t src_rows, uint32_t inner,
                                                              uint32_t K, int rows, int row_groups, int kb_offset,
                                                              int k_blocks) {
  const int z = blockIdx.z;
  pack_split_body<VEC4>(pp.x[z], src_rows, inner, K, nullptr, rows, pp.hi[z], pp.lo[z], row_groups, kb_offset,
                        k_blocks, pp.sumsq[z], pp.sum[z]);
}

struct Im2colGeom {
  int64_t N, cin_src, H, W, Ho, Wo, cmerged;
  int kh, kw, sh, sw, ph, pw, dh, dw, ones_row;
};

// FAST1X1: 1x1 kernel, stride 1, no padding, Ho*Wo % 4 == 0, 16-byte aligned sources — the four k of a thread are
// four consecutive positions of one image: one 128-bit load per source instead of four decoded scalar gathers
// (36 of the 53 convolution inputs of a ResNet-50 and all 54 target packs: 57 % of the bytes this kernel writes).
template <bool FAST1X1>
__global__ void __launch_bounds__(256) pack_im2col_kernel(const float *__restrict__ x1, const float *__restrict__ x2,
                                                          const int32_t *__restrict__ chan1,
                                                          const int32_t *__restrict__ chan2,
                                                          const float *__restrict__ scale1,
                                                          const float *__restrict__ scale2, Im2colGeom gm, int rows,
                                                          uint32_t K, float *__restrict__ hi, float *__restrict__ lo,
                                                          int row_groups, int kb_offset, int k_blocks) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane & 7, jj = lane >> 3;
  const int g = blockIdx.y;
  const int row = g * 8 + r;
  const int taps = gm.kh * gm.kw;
  const int feat_rows = (int)gm.cmerged * taps;
  // decode the feature row once: (c, dy, dx)
  int c1 = -1, c2 = -1, dy = 0, dx = 0;
  float w1 = 0.f, w2 = 0.f;
  const bool is_ones = gm.ones_row && row == feat_rows;
  if (row < feat_rows) {
    const int c = row / taps, t = row - c * taps;
    dy = t / gm.kw;
    dx = t - dy * gm.kw;
    c1 = chan1 ? chan1[c] : c;
    c2 = chan2 ? chan2[c] : -1;
    w1 = scale1 ? scale1[c] : 1.f;
    w2 = scale2 ? scale2[c] : 0.f;
  }
  const uint32_t HoWo = (uint32_t)(gm.Ho * gm.Wo);
  const int64_t plane = gm.H * gm.W;
  const int kb_end = min(k_blocks, (int)(blockIdx.x + 1) * kKbPerBlock);
  const int Wo = (int)gm.Wo, Ho = (int)gm.Ho;
  const int64_t img = (int64_t)gm.cin_src * plane;
  const float *b1 = (c1 >= 0) ? x1 + (int64_t)c1 * plane : nullptr;
  const float *b2 = (c2 >= 0) ? x2 + (int64_t)c2 * plane : nullptr;
#pragma unroll 2
  for (int kb = blockIdx.x * kKbPerBlock + warp; kb < kb_end; kb += 8) {
    const uint32_t k = (uint32_t)kb * kPackK + jj * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < rows && k < K) {
      if (is_ones) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (k + e < K) ? 1.f : 0.f;
      } else {
        if constexpr (FAST1X1) {
          const uint32_t n = k / HoWo;
          const int64_t off = (int64_t)n * img + (int64_t)(k - n * HoWo);
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 a = b1 ? __ldg(reinterpret_cast<const float4 *>(b1 + off)) : z;
          const float4 b = b2 ? __ldg(reinterpret_cast<const float4 *>(b2 + off)) : z;
          v[0] = fmaf(w1, a.x, w2 * b.x);  // same rounding as the generic path below
          v[1] = fmaf(w1, a.y, w2 * b.y);
          v[2] = fmaf(w1, a.z, w2 * b.z);
          v[3] = fmaf(w1, a.w, w2 * b.w);
        } else {
        // decode (n, ho, wo) once, then step along the output row
        uint32_t n = k / HoWo;
        const uint32_t p = k - n * HoWo;
        int ho = (int)(p / (uint32_t)Wo), wo = (int)(p - (uint32_t)ho * (uint32_t)Wo);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (k + e < K) {
            const int h = ho * gm.sh - gm.ph + dy * gm.dh, wq = wo * gm.sw - gm.pw + dx * gm.dw;
            if (h >= 0 && h < gm.H && wq >= 0 && wq < gm.W) {
              const int64_t off = (int64_t)n * img + (int64_t)h * gm.W + wq;
              const float a = b1 ? __ldg(b1 + off) : 0.f;
              const float b = b2 ? __ldg(b2 + off) : 0.f;
              v[e] = fmaf(w1, a, w2 * b);  // (a + b) / 2 for merged channels: one rounding of the sum
            }
          }
          if (++wo == Wo) {
            wo = 0;
            if (++ho == Ho) {
              ho = 0;
              ++n;
            }
          }
        }
        }
      }
    }
    split_store(v, hi, lo, panel_offset(kb_offset + kb, g, row_groups) + (jj * 8 + r) * 4);
  }
}

}  // namespace plb

extern "C" int plb_pack_split(const float *x, int64_t outer, int64_t src_rows, int64_t inner,
                              const int64_t *row_index, int64_t rows, float *hi, float *lo, int32_t row_groups,
                              int32_t kb_offset, double *row_sumsq, double *row_sum, void *stream) {
  using namespace plb;
  PLB_REQUIRE(x && hi && lo, PLB_EINVAL, "plb_pack_split: null pointer");
  PLB_REQUIRE(outer > 0 && src_rows > 0 && inner > 0 && rows > 0, PLB_EINVAL, "plb_pack_split: empty operand");
  PLB_REQUIRE(row_index != nullptr || rows <= src_rows, PLB_EINVAL, "plb_pack_split: rows > src_rows without index");
  const int64_t K = outer * inner;
  PLB_REQUIRE(K < (int64_t)1 << 31 && inner < (int64_t)1 << 31, PLB_ESIZE, "plb_pack_split: K too large");
  PLB_REQUIRE(row_groups > 0 && (int64_t)row_groups * 8 >= rows, PLB_EINVAL,
              "plb_pack_split: row_groups must cover rows");
  PLB_REQUIRE(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0, PLB_EALIGN, "plb_pack_split: planes unaligned");
  const int k_blocks = (int)ceil_div(K, kPackK);
  dim3 grid((unsigned)ceil_div(k_blocks, kKbPerBlock), (unsigned)ceil_div(rows, 8));
  PLB_REQUIRE(grid.y <= 65535, PLB_ESIZE, "plb_pack_split: too many rows");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec4 = (inner % 4 == 0) && (((uintptr_t)x & 15) == 0);
  if (vec4)
    pack_split_kernel<true><<<grid, 256, 0, s>>>(x, src_rows, (uint32_t)inner, (uint32_t)K, row_index, (int)rows, hi,
                                                 lo, row_groups, kb_offset, k_blocks, row_sumsq, row_sum);
  else
    pack_split_kernel<false><<<grid, 256, 0, s>>>(x, src_rows, (uint32_t)inner, (uint32_t)K, row_index, (int)rows, hi,
                                                  lo, row_groups, kb_offset, k_blocks, row_sumsq, row_sum);
  return launch_status("pack_split_kernel");
}

extern "C" int plb_pack_split_pair_sums(const float *xa, const float *xb, int64_t outer, int64_t src_rows,
                                        int64_t inner, float *hi_a, float *lo_a, float *hi_b, float *lo_b,
                                        int32_t row_groups, int32_t kb_offset, double *sumsq_a, double *sumsq_b,
                                        double *sum_a, double *sum_b, void *stream) {
  using namespace plb;
  PLB_REQUIRE(xa && xb && hi_a && lo_a && hi_b && lo_b, PLB_EINVAL, "plb_pack_split_pair: null pointer");
  PLB_REQUIRE(outer > 0 && src_rows > 0 && inner > 0, PLB_EINVAL, "plb_pack_split_pair: empty operand");
  PLB_REQUIRE((sumsq_a == nullptr) == (sumsq_b == nullptr) && (sum_a == nullptr) == (sum_b == nullptr), PLB_EINVAL,
              "plb_pack_split_pair: row statistics must be requested for both operands or neither");
  const int64_t K = outer * inner, rows = src_rows;
  PLB_REQUIRE(K < (int64_t)1 << 31 && inner < (int64_t)1 << 31, PLB_ESIZE, "plb_pack_split_pair: K too large");
  PLB_REQUIRE(row_groups > 0 && (int64_t)row_groups * 8 >= rows, PLB_EINVAL,
              "plb_pack_split_pair: row_groups must cover rows");
  PLB_REQUIRE((((uintptr_t)hi_a | (uintptr_t)lo_a | (uintptr_t)hi_b | (uintptr_t)lo_b) & 15) == 0, PLB_EALIGN,
              "plb_pack_split_pair: planes unaligned");
  const int k_blocks = (int)ceil_div(K, kPackK);
  dim3 grid((unsigned)ceil_div(k_blocks, kKbPerBlock), (unsigned)ceil_div(rows, 8), 2);
  PLB_REQUIRE(grid.y <= 65535, PLB_ESIZE, "plb_pack_split_pair: too many rows");
  cudaStream_t s = (cudaStream_t)stream;
  PackPair pp{{xa, xb}, {hi_a, hi_b}, {lo_a, lo_b}, {sumsq_a, sumsq_b}, {sum_a, sum_b}};
  const bool vec4 = (inner % 4 == 0) && ((((uintptr_t)xa | (uintptr_t)xb) & 15) == 0);
  if (vec4)
    pack_split_pair_kernel<true><<<grid, 256, 0, s>>>(pp, src_rows, (uint32_t)inner, (uint32_t)K, (int)rows,
                                                      row_groups, kb_offset, k_blocks);
  else
    pack_split_pair_kernel<false><<<grid, 256, 0, s>>>(pp, src_rows, (uint32_t)inner, (uint32_t)K, (int)rows,
                                                       row_groups, kb_offset, k_blocks);
  return launch_status("pack_split_pair_kernel");
}

extern "C" int plb_pack_split_pair(const float *xa, const float *xb, int64_t outer, int64_t src_rows, int64_t inner,
                                   float *hi_a, float *lo_a, float *hi_b, float *lo_b, int32_t row_groups,
                                   int32_t kb_offset, double *sumsq_a, double *sumsq_b, void *stream) {
  return plb_pack_split_pair_sums(xa, xb, outer, src_rows, inner, hi_a, lo_a, hi_b, lo_b, row_groups, kb_offset,
                                  sumsq_a, sumsq_b, nullptr, nullptr, stream);
}

extern "C" int plb_pack_im2col(const float *x1, const float *x2, int64_t N, int64_t cin_src, int64_t H, int64_t W,
                               const int32_t *chan1, const int32_t *chan2, const float *scale1, const float *scale2,
                               int64_t cmerged, int32_t kh, int32_t kw, int32_t stride_h, int32_t stride_w,
                               int32_t pad_h, int32_t pad_w, int32_t dil_h, int32_t dil_w, int64_t Ho, int64_t Wo,
                               int32_t ones_row, float *hi, float *lo, int32_t row_groups, int32_t kb_offset,
                               void *stream) {
  using namespace plb;
  PLB_REQUIRE(x1 && hi && lo, PLB_EINVAL, "plb_pack_im2col: null pointer");
  PLB_REQUIRE(N > 0 && cin_src > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && cmerged > 0 && kh > 0 && kw > 0,
              PLB_EINVAL, "plb_pack_im2col: empty geometry");
  PLB_REQUIRE(x2 != nullptr || chan2 == nullptr, PLB_EINVAL, "plb_pack_im2col: chan2 given without x2");
  const int64_t rows = cmerged * kh * kw + (ones_row ? 1 : 0);
  const int64_t K = N * Ho * Wo;
  PLB_REQUIRE(K < (int64_t)1 << 31, PLB_ESIZE, "plb_pack_im2col: K too large");
  PLB_REQUIRE(row_groups > 0 && (int64_t)row_groups * 8 >= rows, PLB_EINVAL,
              "plb_pack_im2col: row_groups must cover rows");
  Im2colGeom gm{N, cin_src, H, W, Ho, Wo, cmerged, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, ones_row};
  const int k_blocks = (int)ceil_div(K, kPackK);
  dim3 grid((unsigned)ceil_div(k_blocks, kKbPerBlock), (unsigned)ceil_div(rows, 8));
  PLB_REQUIRE(grid.y <= 65535, PLB_ESIZE, "plb_pack_im2col: too many rows");
  const bool fast = kh == 1 && kw == 1 && stride_h == 1 && stride_w == 1 && pad_h == 0 && pad_w == 0 && Ho == H &&
                    Wo == W && (H * W) % 4 == 0 && ((uintptr_t)x1 % 16) == 0 && ((uintptr_t)x2 % 16) == 0;
  if (fast)
    pack_im2col_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x1, x2, chan1, chan2, scale1, scale2, gm,
                                                                    (int)rows, (uint32_t)K, hi, lo, row_groups,
                                                                    kb_offset, k_blocks);
  else
    pack_im2col_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x1, x2, chan1, chan2, scale1, scale2, gm,
                                                                     (int)rows, (uint32_t)K, hi, lo, row_groups,
                                                                     kb_offset, k_blocks);
  return launch_status("pack_im2col_kernel");
}
