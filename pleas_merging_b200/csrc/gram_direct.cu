// Fused cross-Gram for NARROW taps (C <= 128): reads the two fp32 activations straight from
// their [outer][C][inner] layout — no packed planes — splits every value into tf32 hi + lo on the
// way into shared memory and contracts with the same 3xTF32 tcgen05 pipeline as gemm.cu.
//
// Why a second kernel: a tap with C channels has arithmetic intensity C/4 flop per activation
// byte (activation_matching.py:26-28, 44-46: X Y^T over K = batch x spatial positions), so the
// C = 64 / 128 taps of a ResNet (46 of the 174 taps of a ResNet-50 pair, K up to 401 408) are
// HBM-bound.  Through packed planes they move 20 B per element (4 read + 8 written by pack.cu, 8
// read by gemm.cu); here they move 4.  The converter stage costs shared-memory bandwidth (write
// hi/lo: 2 x 4 B per element on top of the MMA's operand reads), which is why the WIDE taps keep
// the packed-plane kernel: at BN = 256 that stage would make the tensor-bound kernel smem-bound.
//
// One output tile (128 x BN accumulator, rows/cols >= C unused), K split over the CTAs.
// Warp roles (512 threads, 1 CTA/SM):
//   warp 0      TMEM owner + MMA issuer
//   warps 1-8   promotion / epilogue (as gemm3xtf32_v2_kernel)
//   warps 9-10  LOADERS (x / y): cp.async of the raw fp32 k-blocks straight into the MMA stages' hi planes in
//               the K-major core-matrix order (lane 8 jj + r = row r, jj-th 16-byte k-chunk of an
//               8-row panel), completion signalled per stage with cp.async.mbarrier.arrive — it runs
//               up to kStages k-blocks (96+ KB per SM) ahead and never executes a fence;
//   warps 11-15 CONVERTERS: read a landed hi plane, write lo = x - trunc_tf32(x) into the stage's lo
//               plane, keep the rows' running sums of squares, proxy-fence, signal the MMA warp.
// The hi operand is the RAW fp32 word: kind::tf32 ignores the 13 low mantissa bits, i.e. the tensor
// core multiplies trunc_tf32(x), and lo is taken against exactly that value, so hi + lo = x.
// Measured (ncu profiles/gram_direct_c64_r01_raw.csv, C = 64, K = 401 408): 96.5 us, DRAM read =
// algorithmic bytes, 2.1 TB/s — limited by how many cp.async two loader warps can keep outstanding
// (converters wait for data 46 % of the time; converters that issue their own cp.async are instead
// gated by their proxy fence, whose MEMBAR.ALL.CTA waits for the thread's outstanding copies).
// Next: TMA 2-D tensor-map loads (legal here because inner % 16 == 0) into 128B-swizzled stages.
#include "common.cuh"

namespace plb {

struct DirectProblem {
  const float *x, *y;   // [outer][C][inner] fp32, 16-byte aligned, inner % 16 == 0
  float *partial;       // [splits][128][BN]
  double *qa, *qb;      // row sums of squares (fp64 atomics) or nullptr
  int C, inner, k_blocks, splits;
};

template <int BN>
struct DirectCfg {
  static constexpr int kPlaneBytes = BN * kPackK * 4;      // one operand plane of one k-block (BN rows)
  static constexpr int kStageBytes = 4 * kPlaneBytes;      // [X raw/hi][X lo][Y raw/hi][Y lo]
  static constexpr int kConvWarps = 5;                     // warps 11-15 (BN = 128: two teams of two, one idle)
  static constexpr int kStages = BN == 128 ? 6 : 10;       // 192 KB / 160 KB of operand stages
  static constexpr int kTeam = BN == 128 ? 2 : 1;          // converter warps per k-block
  static constexpr int kTeams = kConvWarps / kTeam;        // k-block kb is converted by team kb % kTeams
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  static constexpr int kThreads = 512;
  static constexpr int kEpiWarps = 8;
  static_assert(kStages % kTeams == 0, "a stage is always converted by the same team");
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// the mbarrier receives one arrival from this thread once all its prior cp.async have completed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int BN>
__global__ void __launch_bounds__(512, 1) gram_direct_kernel(DirectProblem p, int chain_kb) {
  using Cfg = DirectCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_raw[Cfg::kStages];    // loader -> converters: the raw k-block has landed
  __shared__ uint64_t bar_full[Cfg::kStages];   // converters -> MMA: lo planes written
  __shared__ uint64_t bar_empty[Cfg::kStages];  // MMA -> loader: the stage's MMAs have retired
  __shared__ uint64_t bar_acc_full[2];
  __shared__ uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int kb0 = (int)((int64_t)p.k_blocks * split / p.splits);
  const int nkb = (int)((int64_t)p.k_blocks * (split + 1) / p.splits) - kb0;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::kStages; ++s) {
        mbar_init(&bar_raw[s], 64);  // both loader warps, one arrival per lane
        mbar_init(&bar_full[s], Cfg::kTeam);
        mbar_init(&bar_empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bar_acc_full[b], 1);
        mbar_init(&bar_acc_empty[b], Cfg::kEpiWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_s, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 9 || warp == 10) {
    // ------------------------------------------------------------------ loaders (warp 9: x, warp 10: y)
    const int r = lane & 7, jj = lane >> 3;
    const int groups = p.C >> 3;                 // 8-row groups per operand
    const int64_t img = (int64_t)p.C * p.inner;  // floats per outer index
    const int64_t gstride = (int64_t)8 * p.inner;
    const float *src0 = (warp == 9 ? p.x : p.y) + (int64_t)r * p.inner + jj * 4;
    const uint32_t plane = warp == 9 ? 0u : 2u * Cfg::kPlaneBytes;
    int64_t o = ((int64_t)kb0 * kPackK) / p.inner;
    int64_t i = ((int64_t)kb0 * kPackK) - o * p.inner;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % Cfg::kStages;
      mbar_wait(&bar_empty[s], ((uint32_t)(kb / Cfg::kStages) & 1u) ^ 1u);
      uint8_t *dst = smem + (size_t)s * Cfg::kStageBytes + plane + (uint32_t)lane * 16u;  // (jj * 8 + r) * 16 == lane * 16
      const float *src = src0 + o * img + i;
#pragma unroll 4
      for (int g = 0; g < groups; ++g) cp_async16(dst + (uint32_t)g * 512u, src + g * gstride);
      cp_async_arrive_noinc(&bar_raw[s]);
      i += kPackK;
      if (i >= p.inner) {
        i = 0;
        ++o;
      }
    }
  } else if (warp >= 11) {
    // ------------------------------------------------------------------ converters
    constexpr int P = 16;                        // panels per warp per k-block (2 * BN / 8 / kTeam)
    const int cw = warp - 11;
    const int team = cw / Cfg::kTeam, tw = cw % Cfg::kTeam;
    const bool idle = team >= Cfg::kTeams;       // BN = 128: the fifth converter warp has no partner
    const int groups = p.C >> 3;
    float s2[P];
#pragma unroll
    for (int q = 0; q < P; ++q) s2[q] = 0.f;
    // panel q of this warp: pidx = tw + kTeam * q over [x groups | y groups]
    for (int kb = idle ? nkb : team; kb < nkb; kb += Cfg::kTeams) {
      const int s = kb % Cfg::kStages;
      const uint32_t ph = (uint32_t)(kb / Cfg::kStages) & 1u;
      mbar_wait(&bar_raw[s], ph);
      uint8_t *st = smem + (size_t)s * Cfg::kStageBytes + (uint32_t)lane * 16u;
#pragma unroll
      for (int q = 0; q < P; ++q) {
        const int pidx = tw + Cfg::kTeam * q;
        if (pidx < 2 * groups) {
          const int opnd = pidx >= groups ? 1 : 0;
          uint8_t *hp = st + (opnd ? 2 * Cfg::kPlaneBytes : 0) + (uint32_t)(pidx - opnd * groups) * 512u;
          const float4 v = *reinterpret_cast<const float4 *>(hp);
          float4 l;  // x - trunc_tf32(x): exact, at most 13 significant bits
          l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          *reinterpret_cast<float4 *>(hp + Cfg::kPlaneBytes) = l;
          s2[q] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2[q]))));
        }
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to tcgen05.mma (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s]);
    }
    // row sums of squares: the 4 jj lanes of a row, then one fp64 atomic per row
    if (p.qa != nullptr && !idle) {
#pragma unroll
      for (int q = 0; q < P; ++q) {
        float v = s2[q];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        const int pidx = tw + Cfg::kTeam * q;
        if (pidx < 2 * groups && lane < 8) {
          const int opnd = pidx >= groups ? 1 : 0;
          atomicAdd((opnd ? p.qb : p.qa) + (pidx - opnd * groups) * 8 + lane, (double)v);
        }
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(128, BN);
    uint32_t it = 0, chain = 0;
    for (int i0 = 0; i0 < nkb; i0 += chain_kb, ++chain) {
      const uint32_t buf = chain & 1u;
      mbar_wait(&bar_acc_empty[buf], ((chain >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BN;
      const int i1 = min(nkb, i0 + chain_kb);
      for (int i = i0; i < i1; ++i, ++it) {
        const int s = it % Cfg::kStages;
        mbar_wait(&bar_full[s], (it / Cfg::kStages) & 1u);
        tc_fence_after();
        if (elect_one()) {  // elect.sync: the compiler keeps descriptors in uniform registers (no per-MMA R2UR loop)
          const uint32_t st = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
#pragma unroll
          for (int ks = 0; ks < kPackK / 8; ++ks) {
            const uint32_t koff = ks * 256;
            // A descriptors span 128 rows (8 KB): with BN = 64 rows 64..127 alias the next plane — they
            // only feed accumulator rows nobody reads
            const uint64_t a_hi = umma_desc_kmajor(st + koff, 128, 512);
            const uint64_t a_lo = umma_desc_kmajor(st + Cfg::kPlaneBytes + koff, 128, 512);
            const uint64_t b_hi = umma_desc_kmajor(st + 2 * Cfg::kPlaneBytes + koff, 128, 512);
            const uint64_t b_lo = umma_desc_kmajor(st + 3 * Cfg::kPlaneBytes + koff, 128, 512);
            umma_tf32(d_tmem, a_lo, b_hi, idesc, (i > i0 || ks > 0) ? 1u : 0u);
            umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
            umma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
          }
          umma_commit(&bar_empty[s]);
          if (i == i1 - 1) umma_commit(&bar_acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ promotion / epilogue
    constexpr int COLS = BN / 2;
    const int q = warp & 3;
    const int half = (warp - 1) >> 2;
    uint32_t chain = 0;
    float acc[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc[c] = 0.f;
    for (int i0 = 0; i0 < nkb; i0 += chain_kb, ++chain) {
      const uint32_t buf = chain & 1u;
      mbar_wait(&bar_acc_full[buf], (chain >> 1) & 1u);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * COLS;
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t0 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[c * 32 + e] += __uint_as_float(v[e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[buf]);
    }
    const int64_t row = q * 32 + lane;
    float4 *d4 = reinterpret_cast<float4 *>(p.partial + ((int64_t)split * 128 + row) * BN + half * COLS);
#pragma unroll
    for (int e = 0; e < COLS / 4; ++e) d4[e] = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 2 * BN);
}

template <int BN>
static int launch_direct(const DirectProblem &p, int chain_kb, cudaStream_t stream) {
  if (int rc = ensure_dynamic_smem((const void *)gram_direct_kernel<BN>, DirectCfg<BN>::kSmemBytes,
                                   "gram_direct_kernel"))
    return rc;
  gram_direct_kernel<BN><<<p.splits, DirectCfg<BN>::kThreads, DirectCfg<BN>::kSmemBytes, stream>>>(p, chain_kb);
  return launch_status("gram_direct_kernel");
}

}  // namespace plb

extern "C" int plb_gram_direct(const float *x, const float *y, int64_t outer, int64_t C, int64_t inner,
                               float *partial, int32_t splits, int32_t chain_kb, double *row_sumsq_x,
                               double *row_sumsq_y, void *stream) {
  using namespace plb;
  PLB_REQUIRE(x && y && partial, PLB_EINVAL, "plb_gram_direct: null pointer");
  PLB_REQUIRE(outer > 0 && inner > 0 && C > 0, PLB_EINVAL, "plb_gram_direct: empty operand");
  PLB_REQUIRE(C <= 128 && C % 8 == 0, PLB_ESIZE, "plb_gram_direct: C must be a multiple of 8, at most 128");
  PLB_REQUIRE(inner % 16 == 0, PLB_ESIZE, "plb_gram_direct: inner must be a multiple of 16 (use the packed path)");
  PLB_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)partial) & 15) == 0, PLB_EALIGN,
              "plb_gram_direct: pointers must be 16-byte aligned");
  PLB_REQUIRE((row_sumsq_x == nullptr) == (row_sumsq_y == nullptr), PLB_EINVAL,
              "plb_gram_direct: row statistics for both operands or neither");
  const int64_t kb = outer * inner / kPackK;
  PLB_REQUIRE(kb < ((int64_t)1 << 31), PLB_ESIZE, "plb_gram_direct: K too large");
  PLB_REQUIRE(splits > 0 && splits <= kb && chain_kb > 0, PLB_EINVAL, "plb_gram_direct: bad splits / chain");
  DirectProblem p{x, y, partial, row_sumsq_x, row_sumsq_y, (int)C, (int)inner, (int)kb, splits};
  cudaStream_t s = (cudaStream_t)stream;
  if (C <= 64) return launch_direct<64>(p, chain_kb, s);
  return launch_direct<128>(p, chain_kb, s);
}
