// 3xTF32 tcgen05 GEMM over packed hi/lo planes:  partial[s] = A[:, Ks] * B[:, Ks]^T.
//
// This is the dense contraction behind every GEMM-shaped step of the merge path:
//   * activation statistics  G = X_a X_b^T        (activation_matching.py:26-28, 44-46)
//   * weight-matching cost   A = sum_ax W_a W_b^T (weight_matching.py:65-75)
//   * PLeaS normal equations U^T U, U^T Y         (closed form of pleas_merging.py:281-291)
//
// Operands arrive as packed planes (pack.cu) whose (128·t rows x 16 k) tiles are contiguous,
// so the producer needs no tensor map: one cp.async.bulk per plane per stage lands a tile in
// shared memory already in the K-major no-swizzle core-matrix order tcgen05.mma reads.
//
// Two kernels share this file.  gemm3xtf32_v2_kernel (further down) is the product path: persistent
// CTAs, ping-pong TMEM accumulators, in-kernel promotion.  gemm3xtf32_kernel (v1, kept for A/B runs
// and as the chain-length experiment's subject) is described first.
//
// v1 warp roles (192 threads, 2 CTAs/SM):  warp 0 = bulk-copy producer, warp 1 = TMEM owner +
// single-thread MMA issuer (3 MMAs per 8-wide k-step: lo·hi, hi·lo, hi·hi, fp32 accumulate in
// TMEM), warps 2-5 = epilogue (tcgen05.ld -> global partial tile).  smem ring of kStages
// stages; mbarrier full/empty per stage; tcgen05.commit releases stages and publishes the
// accumulator.  Each CTA owns one SHORT k-chain (<= ops.MAX_CHAIN_KB k-blocks): the tensor core
// accumulates in fp32 with truncation, so long chains drift (-1e-7 relative per k-block,
// profiles/experiments/exp_chain_length.py); chains are summed in fp64 by finalize.cu.
#include "common.cuh"

namespace plb {

template <int BN>
struct GemmCfg {
  static constexpr int kATileBytes = 128 * kPackK * 4;  // 8 KB
  static constexpr int kBTileBytes = BN * kPackK * 4;
  static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
  // 96 KB of stages per CTA so two CTAs share an SM: one CTA's epilogue (TMEM drain + partial
  // tile store) overlaps the other's mainloop.  Chains are short by design (see ops.py:
  // the tensor core's fp32 accumulator truncates, ~1e-7 relative bias per k-block).
  static constexpr int kStages = BN == 256 ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + alignment slack
  static constexpr int kThreads = 192;
};

struct WorkItem {
  const PlbGemmProblem *p;
  int m_tile, n_tile, split, kb0, nkb;
};


// Work items of a problem are numbered (tile, split) with the split fastest.  Symmetric problems
// (A and B are the same operand, e.g. U^T U) only enumerate the tiles that touch the lower triangle
// — row-tile mt has min(n_tiles, (128 mt + 127) / BN + 1) of them — so the persistent CTAs, which
// take items round-robin, stay balanced; finalize.cu never reads the tiles that are not produced.
template <int BN>
__device__ __forceinline__ WorkItem decode_work(const PlbGemmProblem *probs, int nprob, int cta) {
  int lo = 0, hi = nprob - 1;
  while (lo < hi) {  // last problem whose cta_begin <= cta
    int mid = (lo + hi + 1) >> 1;
    if (probs[mid].cta_begin <= cta) lo = mid; else hi = mid - 1;
  }
  WorkItem w;
  w.p = probs + lo;
  int local = cta - w.p->cta_begin;
  int splits = w.p->splits;
  w.split = local % splits;
  int t = local / splits;
  if (w.p->symmetric) {
    int mt = 0;
    for (;; ++mt) {
      const int cnt = min(w.p->n_tiles, (128 * mt + 127) / BN + 1);
      if (t < cnt) break;
      t -= cnt;
    }
    w.m_tile = mt;
    w.n_tile = t;
  } else {
    w.n_tile = t % w.p->n_tiles;
    w.m_tile = t / w.p->n_tiles;
  }
  int64_t kb = w.p->k_blocks;
  w.kb0 = (int)(kb * w.split / splits);
  w.nkb = (int)(kb * (w.split + 1) / splits) - w.kb0;
  return w;
}

template <int BN>
__global__ void __launch_bounds__(192, 2) gemm3xtf32_kernel(const PlbGemmProblem *__restrict__ probs, int nprob) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[Cfg::kStages];
  __shared__ uint64_t bar_empty[Cfg::kStages];
  __shared__ uint64_t bar_acc;
  __shared__ uint32_t tmem_base_s;

  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WorkItem w = decode_work<BN>(probs, nprob, blockIdx.x);
  const PlbGemmProblem *p = w.p;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_acc, 1);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc(&tmem_base_s, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      const int ga = p->a_row_groups, gb = p->b_row_groups;
      const int g0a = w.m_tile * 16;
      const int g0b = w.n_tile * (BN / 8);
      const int gcb = min(BN / 8, gb - g0b);
      const uint32_t b_bytes = (uint32_t)gcb * kPanelFloats * 4;
      const uint32_t a_bytes = (uint32_t)min(16, ga - g0a) * kPanelFloats * 4;
      for (int i = 0; i < w.nkb; ++i) {
        const int s = i % Cfg::kStages;
        const uint32_t phase = (uint32_t)(i / Cfg::kStages) & 1u;
        mbar_wait(&bar_empty[s], phase ^ 1u);
        uint8_t *st = smem + (size_t)s * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&bar_full[s], 2u * a_bytes + 2u * b_bytes);
        const int64_t kb = w.kb0 + i;
        const int64_t oa = panel_offset(kb, g0a, ga), ob = panel_offset(kb, g0b, gb);
        bulk_g2s(st, p->a_hi + oa, a_bytes, &bar_full[s]);
        bulk_g2s(st + Cfg::kATileBytes, p->a_lo + oa, a_bytes, &bar_full[s]);
        bulk_g2s(st + 2 * Cfg::kATileBytes, p->b_hi + ob, b_bytes, &bar_full[s]);
        bulk_g2s(st + 2 * Cfg::kATileBytes + Cfg::kBTileBytes, p->b_lo + ob, b_bytes, &bar_full[s]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(128, BN);
    for (int i = 0; i < w.nkb; ++i) {
      const int s = i % Cfg::kStages;
      const uint32_t phase = (uint32_t)(i / Cfg::kStages) & 1u;
      mbar_wait(&bar_full[s], phase);
      tc_fence_after();
      if (elect_one()) {  // elect.sync: the compiler keeps descriptors in uniform registers (no per-MMA R2UR loop)
        const uint32_t st = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
#pragma unroll
        for (int ks = 0; ks < kPackK / 8; ++ks) {
          // one 8-wide k-step = two 16-byte chunks (LBO 128 B apart); 8-row groups 512 B apart
          const uint32_t koff = ks * 256;
          const uint64_t a_hi = umma_desc_kmajor(st + koff, 128, 512);
          const uint64_t a_lo = umma_desc_kmajor(st + Cfg::kATileBytes + koff, 128, 512);
          const uint64_t b_hi = umma_desc_kmajor(st + 2 * Cfg::kATileBytes + koff, 128, 512);
          const uint64_t b_lo = umma_desc_kmajor(st + 2 * Cfg::kATileBytes + Cfg::kBTileBytes + koff, 128, 512);
          umma_tf32(tmem_base, a_lo, b_hi, idesc, (i > 0 || ks > 0) ? 1u : 0u);
          umma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(&bar_empty[s]);               // stage reusable once these MMAs retire
        if (i == w.nkb - 1) umma_commit(&bar_acc);  // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int64_t ld_n = (int64_t)p->n_tiles * BN;
    const int64_t ld_m = (int64_t)p->m_tiles * 128;
    const int64_t row = (int64_t)w.m_tile * 128 + q * 32 + lane;
    float *dst = p->partial + ((int64_t)w.split * ld_m + row) * ld_n + (int64_t)w.n_tile * BN;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      float4 *d4 = reinterpret_cast<float4 *>(dst + c * 32);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        d4[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                            __uint_as_float(v[4 * e + 3]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}


// ------------------------------------------------------------------------------------------
// v2: persistent CTAs with in-kernel promotion.
//
// The tensor core's fp32 accumulator truncates (-1e-7 relative per chained k-block), so a chain
// into one TMEM accumulator is capped at kChainKb k-blocks.  Instead of cutting K into many
// CTAs (v1: one partial tile per chain, summed by finalize.cu), the MMA warp ping-pongs between
// two TMEM accumulators and eight epilogue warps drain every finished chain into fp32 REGISTER
// accumulators with round-to-nearest adds while the next chain is being multiplied.  A CTA
// therefore owns a long K range (splits exist only to fill the 148 SMs), writes its tile once,
// and — being persistent — overlaps that write with the first chains of its next work item.
//
// 320 threads, 1 CTA/SM: warp 0 producer, warp 1 MMA issuer + TMEM owner (2*BN columns),
// warps 2-9 epilogue (two per TMEM lane quadrant, BN/2 columns each).
// ------------------------------------------------------------------------------------------
template <int BN>
struct GemmCfg2 {
  static constexpr int kATileBytes = 128 * kPackK * 4;
  static constexpr int kBTileBytes = BN * kPackK * 4;
  static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);  // 192 KB
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  static constexpr int kThreads = 320;
  static constexpr int kEpiWarps = 8;
};

template <int BN>
__global__ void __launch_bounds__(320, 1) gemm3xtf32_v2_kernel(const PlbGemmProblem *__restrict__ probs, int nprob,
                                                               int total_items, int chain_kb) {
  using Cfg = GemmCfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[Cfg::kStages];
  __shared__ uint64_t bar_empty[Cfg::kStages];
  __shared__ uint64_t bar_acc_full[2];
  __shared__ uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_acc_full[b], 1);
      mbar_init(&bar_acc_empty[b], Cfg::kEpiWarps);
    }
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc(&tmem_base_s, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      uint32_t it = 0;  // global stage counter across work items
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const WorkItem w = decode_work<BN>(probs, nprob, item);
        const PlbGemmProblem *p = w.p;
        const int ga = p->a_row_groups, gb = p->b_row_groups;
        const int g0a = w.m_tile * 16, g0b = w.n_tile * (BN / 8);
        const uint32_t b_bytes = (uint32_t)min(BN / 8, gb - g0b) * kPanelFloats * 4;
        // only the row groups the operand has: rows past them keep stale shared memory and feed
        // output rows nobody reads (a fixed 16-group copy would re-read the next k-block's panel
        // from DRAM: +50 % traffic on the HBM-bound C=64 taps, ncu profiles/gemm_c64_r01_raw.csv)
        const uint32_t a_bytes = (uint32_t)min(16, ga - g0a) * kPanelFloats * 4;
        for (int i = 0; i < w.nkb; ++i, ++it) {
          const int s = it % Cfg::kStages;
          mbar_wait(&bar_empty[s], ((it / Cfg::kStages) & 1u) ^ 1u);
          uint8_t *st = smem + (size_t)s * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&bar_full[s], 2u * a_bytes + 2u * b_bytes);
          const int64_t kb = w.kb0 + i;
          const int64_t oa = panel_offset(kb, g0a, ga), ob = panel_offset(kb, g0b, gb);
          bulk_g2s(st, p->a_hi + oa, a_bytes, &bar_full[s]);
          bulk_g2s(st + Cfg::kATileBytes, p->a_lo + oa, a_bytes, &bar_full[s]);
          bulk_g2s(st + 2 * Cfg::kATileBytes, p->b_hi + ob, b_bytes, &bar_full[s]);
          bulk_g2s(st + 2 * Cfg::kATileBytes + Cfg::kBTileBytes, p->b_lo + ob, b_bytes, &bar_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_tf32(128, BN);
    uint32_t it = 0, chain = 0;  // global stage / chain counters
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const WorkItem w = decode_work<BN>(probs, nprob, item);
      for (int i0 = 0; i0 < w.nkb; i0 += chain_kb, ++chain) {
        const uint32_t buf = chain & 1u;
        mbar_wait(&bar_acc_empty[buf], ((chain >> 1) & 1u) ^ 1u);  // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        const int i1 = min(w.nkb, i0 + chain_kb);
        for (int i = i0; i < i1; ++i, ++it) {
          const int s = it % Cfg::kStages;
          mbar_wait(&bar_full[s], (it / Cfg::kStages) & 1u);
          tc_fence_after();
          if (elect_one()) {  // elect.sync: the compiler keeps descriptors in uniform registers (no per-MMA R2UR loop)
            const uint32_t st = smem_u32(smem + (size_t)s * Cfg::kStageBytes);
#pragma unroll
            for (int ks = 0; ks < kPackK / 8; ++ks) {
              const uint32_t koff = ks * 256;
              const uint64_t a_hi = umma_desc_kmajor(st + koff, 128, 512);
              const uint64_t a_lo = umma_desc_kmajor(st + Cfg::kATileBytes + koff, 128, 512);
              const uint64_t b_hi = umma_desc_kmajor(st + 2 * Cfg::kATileBytes + koff, 128, 512);
              const uint64_t b_lo = umma_desc_kmajor(st + 2 * Cfg::kATileBytes + Cfg::kBTileBytes + koff, 128, 512);
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
    }
  } else {
    // ------------------------------------------------------------------ epilogue / promotion
    constexpr int COLS = BN / 2;
    const int q = warp & 3;              // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;    // which half of the BN columns
    uint32_t chain = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const WorkItem w = decode_work<BN>(probs, nprob, item);
      const PlbGemmProblem *p = w.p;
      float acc[COLS];
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[c] = 0.f;
      for (int i0 = 0; i0 < w.nkb; i0 += chain_kb, ++chain) {
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
      const int64_t ld_n = (int64_t)p->n_tiles * BN, ld_m = (int64_t)p->m_tiles * 128;
      const int64_t row = (int64_t)w.m_tile * 128 + q * 32 + lane;
      float4 *d4 = reinterpret_cast<float4 *>(p->partial + ((int64_t)w.split * ld_m + row) * ld_n +
                                               (int64_t)w.n_tile * BN + half * COLS);
#pragma unroll
      for (int e = 0; e < COLS / 4; ++e) d4[e] = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// SIMT fp32 reference on the same planes (x = hi + lo), same work decomposition and output
// layout.  Cross-check / debug only — never the product path.
template <int BN>
__global__ void __launch_bounds__(256) gemm_simt_ref_kernel(const PlbGemmProblem *__restrict__ probs, int nprob) {
  __shared__ float sa[kPackK][128 + 1];
  __shared__ float sb[kPackK][BN + 1];
  const WorkItem w = decode_work<BN>(probs, nprob, blockIdx.x);
  const PlbGemmProblem *p = w.p;
  const int tid = threadIdx.x;
  const int row = tid & 127, chalf = tid >> 7;
  float acc[BN / 2];
#pragma unroll
  for (int c = 0; c < BN / 2; ++c) acc[c] = 0.f;
  const int g0a = w.m_tile * 16, g0b = w.n_tile * (BN / 8);
  for (int i = 0; i < w.nkb; ++i) {
    const int64_t kb = w.kb0 + i;
    for (int idx = tid; idx < 128 * kPackK; idx += 256) {
      int r = idx & 127, k = idx >> 7;
      float v = 0.f;
      if (g0a + (r >> 3) < p->a_row_groups) {
        int64_t off = panel_offset(kb, g0a + (r >> 3), p->a_row_groups) + ((k >> 2) * 8 + (r & 7)) * 4 + (k & 3);
        v = p->a_hi[off] + p->a_lo[off];
      }
      sa[k][r] = v;
    }
    for (int idx = tid; idx < BN * kPackK; idx += 256) {
      int r = idx % BN, k = idx / BN;
      float v = 0.f;
      if (g0b + (r >> 3) < p->b_row_groups) {
        int64_t off = panel_offset(kb, g0b + (r >> 3), p->b_row_groups) + ((k >> 2) * 8 + (r & 7)) * 4 + (k & 3);
        v = p->b_hi[off] + p->b_lo[off];
      }
      sb[k][r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPackK; ++k) {
      const float a = sa[k][row];
#pragma unroll
      for (int c = 0; c < BN / 2; ++c) acc[c] = fmaf(a, sb[k][2 * c + chalf], acc[c]);
    }
    __syncthreads();
  }
  const int64_t ld_n = (int64_t)p->n_tiles * BN, ld_m = (int64_t)p->m_tiles * 128;
  float *dst = p->partial + ((int64_t)w.split * ld_m + (int64_t)w.m_tile * 128 + row) * ld_n + (int64_t)w.n_tile * BN;
#pragma unroll
  for (int c = 0; c < BN / 2; ++c) dst[2 * c + chalf] = acc[c];
}

template <int BN>
static int launch_gemm_v2(const PlbGemmProblem *probs, int nprob, int total_items, int chain_kb, cudaStream_t stream) {
  if (int rc = ensure_dynamic_smem((const void *)gemm3xtf32_v2_kernel<BN>, GemmCfg2<BN>::kSmemBytes,
                                   "gemm3xtf32_v2_kernel"))
    return rc;
  const int num_sms = device_sm_count();
  const int grid = total_items < num_sms ? total_items : num_sms;
  gemm3xtf32_v2_kernel<BN><<<grid, GemmCfg2<BN>::kThreads, GemmCfg2<BN>::kSmemBytes, stream>>>(probs, nprob,
                                                                                              total_items, chain_kb);
  return launch_status("gemm3xtf32_v2_kernel");
}

template <int BN>
static int launch_gemm(const PlbGemmProblem *probs, int nprob, int total_ctas, int impl, cudaStream_t stream) {
  if (impl >= 16) return launch_gemm_v2<BN>(probs, nprob, total_ctas, impl >> 4, stream);
  if (impl == 1) {
    gemm_simt_ref_kernel<BN><<<total_ctas, 256, 0, stream>>>(probs, nprob);
    return launch_status("gemm_simt_ref_kernel");
  }
  if (int rc = ensure_dynamic_smem((const void *)gemm3xtf32_kernel<BN>, GemmCfg<BN>::kSmemBytes, "gemm3xtf32_kernel"))
    return rc;
  gemm3xtf32_kernel<BN><<<total_ctas, GemmCfg<BN>::kThreads, GemmCfg<BN>::kSmemBytes, stream>>>(probs, nprob);
  return launch_status("gemm3xtf32_kernel");
}

}  // namespace plb

extern "C" int plb_gemm_grouped(const PlbGemmProblem *problems_dev, int32_t n_problems, int32_t total_ctas,
                                int32_t bn, int32_t impl, void *stream) {
  using namespace plb;
  PLB_REQUIRE(problems_dev != nullptr && n_problems > 0 && total_ctas > 0, PLB_EINVAL,
              "plb_gemm_grouped: empty problem table");
  PLB_REQUIRE(impl == 0 || impl == 1 || (impl >= 16 && (impl & 15) == 0), PLB_EINVAL,
              "plb_gemm_grouped: impl must be 0 (tcgen05 v1), 1 (simt ref) or 16*chain_kb (persistent tcgen05)");
  cudaStream_t s = (cudaStream_t)stream;
  switch (bn) {
    case 64: return launch_gemm<64>(problems_dev, n_problems, total_ctas, impl, s);
    case 128: return launch_gemm<128>(problems_dev, n_problems, total_ctas, impl, s);
    case 256: return launch_gemm<256>(problems_dev, n_problems, total_ctas, impl, s);
  }
  set_error("plb_gemm_grouped: bn must be 64, 128 or 256 (got %d)", bn);
  return PLB_EINVAL;
}
