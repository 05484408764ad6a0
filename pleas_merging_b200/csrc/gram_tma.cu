// TMA-fed fused cross-Gram, straight from the fp32 activations:  partial[s] = X[:, Ks] Y[:, Ks]^T.
//
// The statistic behind cross_features_inner_product / cross_features_cdist
// (pleas/methods/activation_matching.py:14-46) is a contraction of two [C, K] views of NCHW
// activations over K = batch x spatial positions.  An [outer][C][inner] fp32 tensor whose inner
// extent is a multiple of 4 elements is a legal 3-D tensor map (inner, C, outer), so the operands
// need no packing pass at all: TMA lands (rows x 32 k) boxes in shared memory in the 128-byte
// swizzled K-major layout tcgen05.mma reads, converter warps derive the tf32 "lo" plane in place
// (4 bytes of HBM traffic per element instead of the 20 of pack + packed-plane GEMM), and the
// contraction is the same 3xTF32 scheme as gemm.cu:
//     x = hi + lo,  hi = trunc_tf32(x) (the tensor core ignores the 13 low mantissa bits of the raw
//     fp32 word, so the raw tile IS the hi operand),  lo = rna_tf32(x - hi)  (exact difference,
//     rounded to nearest so the core's truncation of lo is exact),
//     X Y^T ~= lo_x hi_y^T + hi_x lo_y^T + hi_x hi_y^T      (fp32 accumulation in TMEM).
// The tensor core's fp32 accumulator truncates (gemm.cu), so chains are short: the MMA warp
// ping-pongs between two TMEM accumulators and eight promotion warps drain every finished chain
// into fp32 registers with round-to-nearest adds.
//
// Two shapes of one kernel:
//   CG = 1, tile 128 x {64,128}   narrow taps (C <= 128; HBM-bound, one output tile, K split
//                                 over the SMs)
//   CG = 2, tile 256 x 256        wide taps: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns a
//                                 256 x 256 tile; each CTA stages 128 rows of X and 128 rows of Y,
//                                 i.e. half the B operand per CTA — what keeps the in-kernel hi/lo
//                                 split within the SM's shared-memory bandwidth.
// Persistent: work items (tile, K split) are dealt round-robin to the clusters.
//
// Warp roles (384 threads, 1 CTA/SM):
//   warp 0      TMA producer (one elected lane): two 3-D box loads per stage
//   warp 1      TMEM owner; in the leader CTA the MMA issuer
//   warps 2-3   converters: raw tile -> lo tile, row sums of squares, proxy fence, signal the
//               LEADER's barrier (remote mbarrier arrive for the peer CTA)
//   warps 4-11  promotion / epilogue (two per TMEM lane quadrant, half the columns each)
#include "tma.cuh"

namespace plb {

// [outer][C][inner] fp32 as a 3-D tensor map (inner, C, outer), boxes of (box_k k, box_rows, 1) whose rows are
// one swizzle row: 32 k = 128-byte swizzle, 16 k = 64-byte swizzle; out-of-range rows / k read as zeros.
static int make_operand_map(CUtensorMap *m, const float *base, int64_t outer, int64_t C, int64_t inner,
                            int box_rows, int box_k) {
  EncodeTiledFn enc = encode_tiled();
  PLB_REQUIRE(enc != nullptr, PLB_EINVAL, "plb_gram_tma: cuTensorMapEncodeTiled is not available from this driver");
  static int exp_flags = -1;  // experiments only (PLB_TMA_EXP): bit 1 = no L2 promotion, bit 2 = 128 B promotion
  if (exp_flags < 0) {
    const char *e = getenv("PLB_TMA_EXP");
    exp_flags = e ? atoi(e) : 0;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)C, (cuuint64_t)outer};
  cuuint64_t gstride[2] = {(cuuint64_t)inner * 4, (cuuint64_t)inner * (cuuint64_t)C * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_k == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   (exp_flags & 2) ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                   : ((exp_flags & 4) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PLB_REQUIRE(r == CUDA_SUCCESS, PLB_EINVAL, "plb_gram_tma: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PLB_OK;
}

struct TmaGramParams {
  float *partial;    // [splits][ld_m][ld_n]
  double *qa, *qb;   // row sums of squares (fp64 atomics) or nullptr
  double *sa, *sb;   // row sums (correlation statistic; SUMS instantiation only)
  int C, inner, boxes_per_image, total_boxes;
  int m_tiles, n_tiles, splits, total_items;
  int ld_m, ld_n, chain_boxes;
  unsigned long long *trace;  // experiments only (plb_debug_set_trace): per-box clock64 stamps of CTA 0, or nullptr
  int debug;  // experiments only (PLB_TMA_DEBUG): 1 = converters skip their work, 2 = hi x hi MMA only (WRONG results)
};

template <int CG, int TM, int TN, int KBOX>
struct TmaCfg {
  static constexpr int kBoxK = KBOX;                    // k per TMA box = one swizzle row of fp32 (32: 128 B, 16: 64 B)
  static constexpr int kRowBytes = KBOX * 4;
  static constexpr int kCpr = kRowBytes / 16;           // 16-byte chunks per tile row
  static constexpr int kXBytes = TM * kRowBytes, kYBytes = TN * kRowBytes;
  static constexpr int kRawBytes = kXBytes + kYBytes;   // what TMA lands per stage and CTA
  static constexpr int kStageBytes = 2 * kRawBytes;     // [X raw = hi][Y raw = hi][X lo][Y lo]
  static constexpr int kStages = (192 * 1024) / kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  static constexpr int kThreads = 384;
  static constexpr int kConvWarps = 2, kEpiWarps = 8;
  static constexpr int kMmaM = 128 * CG, kMmaN = TN * CG;
  static constexpr int kTmemCols = 2 * kMmaN;
  static constexpr int kCols = kMmaN / 2;               // accumulator columns per promotion warp
  static constexpr int kChunks = (TM + TN) * kCpr;      // 16-byte chunks of the raw region
  static constexpr int kPer = kChunks / (kConvWarps * 32);
  static_assert(kStages >= 2 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "geometry");
  static_assert(kChunks % (kConvWarps * 32) == 0, "converter mapping");
};

template <int CG, int TM, int TN, int KBOX, bool SUMS>
__global__ void __launch_bounds__(384, 1) gram_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                          const __grid_constant__ CUtensorMap tmy, TmaGramParams p) {
  using Cfg = TmaCfg<CG, TM, TN, KBOX>;
  constexpr int kBoxK = KBOX;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_raw[Cfg::kStages];    // TMA -> converters (this CTA)
  __shared__ uint64_t bar_conv[Cfg::kStages];   // converters of both CTAs -> MMA issuer (leader's copy is used)
  __shared__ uint64_t bar_empty[Cfg::kStages];  // MMA -> producer (multicast to both CTAs)
  __shared__ uint64_t bar_acc_full[2];          // MMA -> promotion warps (multicast)
  __shared__ uint64_t bar_acc_empty[2];         // promotion warps of both CTAs -> MMA issuer (leader's copy)
  __shared__ uint32_t tmem_base_s;

  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[1016] = clock64();
    p.trace[1022] = gt;
  }
  const int cluster_id = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_clusters = (int)gridDim.x / CG;
  const int tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&bar_raw[s], 1);
      mbar_init(&bar_conv[s], CG * Cfg::kConvWarps);
      mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_acc_full[b], 1);
      mbar_init(&bar_acc_empty[b], CG * Cfg::kEpiWarps);
    }
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc_cg<CG>(&tmem_base_s, Cfg::kTmemCols);
    tmem_relinquish_cg<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[1017] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t s = 0, ph = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      const int tile = item % tiles, split = item / tiles;
      const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
      const int box0 = (int)((int64_t)p.total_boxes * split / p.splits);
      const int nbox = (int)((int64_t)p.total_boxes * (split + 1) / p.splits) - box0;
      const int row_x = mt * (TM * CG) + (int)rank * TM, row_y = nt * (TN * CG) + (int)rank * TN;
      int img = box0 / p.boxes_per_image, kb = box0 - img * p.boxes_per_image;
      for (int i = 0; i < nbox; ++i) {
        mbar_wait_wd(&bar_empty[s], ph ^ 1u);
        if (p.trace && blockIdx.x == 0 && lane == 0 && i < 256) p.trace[i * 4 + 0] = clock64();
        if (elect_one()) {
          uint8_t *st = smem + (size_t)s * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&bar_raw[s], Cfg::kRawBytes);
          tma_load_3d(st, &tmx, kb * kBoxK, row_x, img, &bar_raw[s]);
          tma_load_3d(st + Cfg::kXBytes, &tmy, kb * kBoxK, row_y, img, &bar_raw[s]);
        }
        __syncwarp();
        if (++kb == p.boxes_per_image) {
          kb = 0;
          ++img;
        }
        if (++s == Cfg::kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(Cfg::kMmaM, Cfg::kMmaN);
      const uint32_t smem_base = smem_u32(smem);
      // Every clock between the last MMA of a box and the first of the next is an idle tensor core, and a barrier
      // probe costs 100-250 clocks on an SM whose issue slots are contended by the converter / promotion warps
      // (conv.cu, profiles/r02_notes.md): the NEXT box's barrier (and the next chain's accumulator) is probed with a
      // non-blocking test_wait BEFORE this box's MMAs are issued, so its latency hides behind them.
      uint32_t s = 0, ph = 0, chain = 0;
      bool ready = false, acc_ok = false;
      for (int item = cluster_id; item < p.total_items; item += n_clusters) {
        const int split = item / tiles;
        const int box0 = (int)((int64_t)p.total_boxes * split / p.splits);
        const int nbox = (int)((int64_t)p.total_boxes * (split + 1) / p.splits) - box0;
        int kb = box0 % p.boxes_per_image;
        for (int b0 = 0; b0 < nbox; b0 += p.chain_boxes, ++chain) {
          const uint32_t buf = chain & 1u;
          if (!acc_ok) mbar_wait_wd(&bar_acc_empty[buf], ((chain >> 1) & 1u) ^ 1u);  // both CTAs drained this buffer
          acc_ok = false;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * Cfg::kMmaN;
          const int b1 = min(nbox, b0 + p.chain_boxes);
          for (int b = b0; b < b1; ++b) {
            if (!ready) mbar_wait_wd(&bar_conv[s], ph);
            tc_fence_after();
            if (p.trace && blockIdx.x == 0 && lane == 0 && b < 256) p.trace[b * 4 + 3] = clock64();
            const uint32_t s1 = (s + 1 == Cfg::kStages) ? 0u : s + 1, ph1 = (s + 1 == Cfg::kStages) ? ph ^ 1u : ph;
            const bool next_ready = mbar_test_wait(&bar_conv[s1], ph1);
            bool next_acc = false;
            if (b == b1 - 1)
              next_acc = mbar_test_wait(&bar_acc_empty[(chain + 1) & 1u], (((chain + 1) >> 1) & 1u) ^ 1u);
            const int valid = min(kBoxK, p.inner - kb * kBoxK);  // k beyond the image's extent is TMA zero fill
            const int nk8 = (valid + 7) >> 3;
            if (elect_one()) {
              const uint32_t st = smem_base + s * Cfg::kStageBytes;
#pragma unroll
              for (int j = 0; j < kBoxK / 8; ++j) {
                if (j < nk8 && p.debug != 3) {
                  const uint64_t a_hi = umma_desc_sw<Cfg::kRowBytes>(st + 32 * j);
                  const uint64_t b_hi = umma_desc_sw<Cfg::kRowBytes>(st + Cfg::kXBytes + 32 * j);
                  const uint64_t a_lo = umma_desc_sw<Cfg::kRowBytes>(st + Cfg::kRawBytes + 32 * j);
                  const uint64_t b_lo = umma_desc_sw<Cfg::kRowBytes>(st + Cfg::kRawBytes + Cfg::kXBytes + 32 * j);
                  if (p.debug != 2) {
                    umma_tf32_cg<CG>(d_tmem, a_lo, b_hi, idesc, (b > b0 || j > 0) ? 1u : 0u);
                    umma_tf32_cg<CG>(d_tmem, a_hi, b_lo, idesc, 1u);
                    umma_tf32_cg<CG>(d_tmem, a_hi, b_hi, idesc, 1u);
                  } else {
                    umma_tf32_cg<CG>(d_tmem, a_hi, b_hi, idesc, (b > b0 || j > 0) ? 1u : 0u);
                  }
                }
              }
              umma_commit_cg<CG>(&bar_empty[s]);
              if (b == b1 - 1) umma_commit_cg<CG>(&bar_acc_full[buf]);
            }
            ready = __all_sync(0xffffffffu, next_ready);
            if (b == b1 - 1) acc_ok = __all_sync(0xffffffffu, next_acc);
            if (++kb == p.boxes_per_image) kb = 0;
            if (++s == Cfg::kStages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp < 2 + Cfg::kConvWarps) {
    // ------------------------------------------------------------------ converters
    const int ct = (warp - 2) * 32 + lane;  // chunk q of this thread: (q * 64 + ct) * 16 bytes into the raw region
    uint32_t s = 0, ph = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      const int tile = item % tiles, split = item / tiles;
      const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
      const int box0 = (int)((int64_t)p.total_boxes * split / p.splits);
      const int nbox = (int)((int64_t)p.total_boxes * (split + 1) / p.splits) - box0;
      float s2[Cfg::kPer], s1[SUMS ? Cfg::kPer : 1];
#pragma unroll
      for (int q = 0; q < Cfg::kPer; ++q) s2[q] = 0.f;
#pragma unroll
      for (int q = 0; q < (SUMS ? Cfg::kPer : 1); ++q) s1[q] = 0.f;
      for (int i = 0; i < nbox; ++i) {
        mbar_wait_wd(&bar_raw[s], ph);
        if (p.trace && blockIdx.x == 0 && ct == 0 && i < 256) p.trace[i * 4 + 1] = clock64();
        uint8_t *st = smem + (size_t)s * Cfg::kStageBytes + (uint32_t)ct * 16u;
        if (p.debug != 1 && p.debug != 3)
#pragma unroll
        for (int q = 0; q < Cfg::kPer; ++q) {
          const float4 v = *reinterpret_cast<const float4 *>(st + q * 1024);
          // x - trunc_tf32(x) is exact (<= 13 significant bits).  Adding half a tf32 ulp (0x1000) to its bit
          // pattern and letting the tensor core drop the 13 low bits rounds it to nearest (ties away): the
          // final mask is the hardware's, so the split costs two integer ops and one subtraction per value.
          float4 l;
          l.x = __uint_as_float(__float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u)) + 0x1000u);
          l.y = __uint_as_float(__float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)) + 0x1000u);
          l.z = __uint_as_float(__float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u)) + 0x1000u);
          l.w = __uint_as_float(__float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u)) + 0x1000u);
          *reinterpret_cast<float4 *>(st + Cfg::kRawBytes + q * 1024) = l;
          s2[q] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2[q]))));
          if (SUMS) s1[q] += (v.x + v.y) + (v.z + v.w);
        }
        fence_proxy_async();  // generic-proxy stores -> visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive_leader<CG>(&bar_conv[s]);
        if (p.trace && blockIdx.x == 0 && ct == 0 && i < 256) p.trace[i * 4 + 2] = clock64();
        if (++s == Cfg::kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
      // Row sums of squares: chunk (q * 64 + ct) lies in tile row (q * 64 + ct) / kCpr; the kCpr lanes of a row
      // combine, then one fp64 atomic per row.  X rows are counted by the tiles of column 0, Y rows by those of row 0.
      if (p.qa != nullptr) {
        const int row_x = mt * (TM * CG) + (int)rank * TM, row_y = nt * (TN * CG) + (int)rank * TN;
#pragma unroll
        for (int q = 0; q < Cfg::kPer; ++q) {
          float v = s2[q], w = SUMS ? s1[q] : 0.f;
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (Cfg::kCpr == 8) v += __shfl_xor_sync(0xffffffffu, v, 4);
          if (SUMS) {
            w += __shfl_xor_sync(0xffffffffu, w, 1);
            w += __shfl_xor_sync(0xffffffffu, w, 2);
            if (Cfg::kCpr == 8) w += __shfl_xor_sync(0xffffffffu, w, 4);
          }
          const int r = (q * 64 + ct) / Cfg::kCpr;
          if ((ct & (Cfg::kCpr - 1)) == 0) {
            if (r < TM) {
              if (nt == 0 && row_x + r < p.C) {
                atomicAdd(p.qa + row_x + r, (double)v);
                if (SUMS) atomicAdd(p.sa + row_x + r, (double)w);
              }
            } else {
              if (mt == 0 && row_y + (r - TM) < p.C) {
                atomicAdd(p.qb + row_y + (r - TM), (double)v);
                if (SUMS) atomicAdd(p.sb + row_y + (r - TM), (double)w);
              }
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ promotion / epilogue
    constexpr int COLS = Cfg::kCols;
    const int q = warp & 3;             // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;   // which half of the accumulator columns
    uint32_t chain = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      const int tile = item % tiles, split = item / tiles;
      const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
      const int box0 = (int)((int64_t)p.total_boxes * split / p.splits);
      const int nbox = (int)((int64_t)p.total_boxes * (split + 1) / p.splits) - box0;
      float acc[COLS];
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[c] = 0.f;
      for (int b0 = 0; b0 < nbox; b0 += p.chain_boxes, ++chain) {
        const uint32_t buf = chain & 1u;
        mbar_wait_wd(&bar_acc_full[buf], (chain >> 1) & 1u);
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::kMmaN + half * COLS;
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
        if (lane == 0) mbar_arrive_leader<CG>(&bar_acc_empty[buf]);
      }
      const int64_t row = (int64_t)mt * (128 * CG) + rank * 128 + q * 32 + lane;
      float4 *d4 = reinterpret_cast<float4 *>(p.partial + ((int64_t)split * p.ld_m + row) * p.ld_n +
                                               (int64_t)nt * Cfg::kMmaN + half * COLS);
      if (p.trace && blockIdx.x == 0 && warp == 4 && lane == 0) p.trace[1018] = clock64();  // last chain drained
#pragma unroll
      for (int e = 0; e < COLS / 4; ++e) d4[e] = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
      if (p.trace && blockIdx.x == 0 && warp == 4 && lane == 0) p.trace[1019] = clock64();  // partial tile stored
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc_cg<CG>(tmem_base, Cfg::kTmemCols);
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 32) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[1020] = clock64();
    p.trace[1023] = gt;
  }
}

static unsigned long long *g_trace = nullptr;

struct TmaGeometry {
  int cg, tm, tn, m_tiles, n_tiles, ld_m, ld_n;
};

static TmaGeometry tma_geometry(int64_t C) {
  TmaGeometry g;
  if (C <= 64) {
    g = {1, 64, 64, 1, 1, 128, 64};
  } else if (C <= 128) {
    g = {1, 128, 128, 1, 1, 128, 128};
  } else {
    const int t = (int)ceil_div(C, 256);
    g = {2, 128, 128, t, t, t * 256, t * 256};
  }
  return g;
}

template <int CG, int TM, int TN, int KBOX, bool SUMS>
static int launch_tma(const CUtensorMap &tmx, const CUtensorMap &tmy, const TmaGramParams &p, cudaStream_t stream) {
  using Cfg = TmaCfg<CG, TM, TN, KBOX>;
  if (int rc = ensure_dynamic_smem((const void *)gram_tma_kernel<CG, TM, TN, KBOX, SUMS>, Cfg::kSmemBytes,
                                   "gram_tma_kernel"))
    return rc;
  const int clusters = min(p.total_items, device_sm_count() / CG);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CG));
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gram_tma_kernel<CG, TM, TN, KBOX, SUMS>, tmx, tmy, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("gram_tma_kernel<%d,%d,%d>: %s", CG, TM, TN, cudaGetErrorString(e));
    return (int)e;
  }
  return launch_status("gram_tma_kernel");
}

}  // namespace plb

// experiments only: CTA 0 of every later plb_gram_tma launch writes 4 clock64 stamps per box (first 256 boxes
// of its first work item) into dev_buf: [issue ok, raw landed, converted, MMA start]; nullptr switches it off
extern "C" int plb_debug_set_trace(unsigned long long *dev_buf) {
  plb::g_trace = dev_buf;
  return PLB_OK;
}

extern "C" int plb_gram_tma_geometry(int64_t C, int32_t *cta_group, int32_t *m_tiles, int32_t *n_tiles,
                                     int32_t *ld_m, int32_t *ld_n) {
  using namespace plb;
  PLB_REQUIRE(C > 0, PLB_EINVAL, "plb_gram_tma_geometry: C must be positive");
  const TmaGeometry g = tma_geometry(C);
  if (cta_group) *cta_group = g.cg;
  if (m_tiles) *m_tiles = g.m_tiles;
  if (n_tiles) *n_tiles = g.n_tiles;
  if (ld_m) *ld_m = g.ld_m;
  if (ld_n) *ld_n = g.ld_n;
  return PLB_OK;
}

extern "C" int plb_gram_tma(const float *x, const float *y, int64_t outer, int64_t C, int64_t inner, float *partial,
                            int32_t splits, int32_t chain_kb, double *row_sumsq_x, double *row_sumsq_y,
                            double *row_sum_x, double *row_sum_y, void *stream) {
  using namespace plb;
  PLB_REQUIRE(x && y && partial, PLB_EINVAL, "plb_gram_tma: null pointer");
  PLB_REQUIRE(outer > 0 && inner > 0 && C > 0, PLB_EINVAL, "plb_gram_tma: empty operand");
  PLB_REQUIRE(inner % 4 == 0, PLB_ESIZE,
              "plb_gram_tma: inner must be a multiple of 4 (16-byte tensor-map strides; use the packed path)");
  PLB_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)partial) & 15) == 0, PLB_EALIGN,
              "plb_gram_tma: pointers must be 16-byte aligned");
  PLB_REQUIRE((row_sumsq_x == nullptr) == (row_sumsq_y == nullptr), PLB_EINVAL,
              "plb_gram_tma: row statistics for both operands or neither");
  PLB_REQUIRE((row_sum_x == nullptr) == (row_sum_y == nullptr) && (row_sum_x == nullptr || row_sumsq_x != nullptr),
              PLB_EINVAL, "plb_gram_tma: row sums need both operands and the sums of squares");
  PLB_REQUIRE(outer < (1ll << 31) && C < (1ll << 31) && inner < (1ll << 31), PLB_ESIZE, "plb_gram_tma: extent too large");
  const TmaGeometry g = tma_geometry(C);
  // k per TMA box / pipeline stage: 32 (128-byte swizzle rows, the default) or 16 (64-byte swizzle: twice the
  // stages in the same shared memory; measured 5-6 % slower on the wide taps and 10-20 % on the narrow ones,
  // profiles/r02_notes.md).  PLB_TMA_BOXK_WIDE / PLB_TMA_BOXK_NARROW select it for A/B runs.
  static int boxk_wide = 0, boxk_narrow = 0;
  if (boxk_wide == 0) {
    const char *w = getenv("PLB_TMA_BOXK_WIDE"), *n = getenv("PLB_TMA_BOXK_NARROW");
    boxk_wide = (w && atoi(w) == 16) ? 16 : 32;
    boxk_narrow = (n && atoi(n) == 16) ? 16 : 32;
  }
  const int box_k = g.cg == 2 ? boxk_wide : boxk_narrow;
  const int64_t bpi = ceil_div(inner, box_k);
  const int64_t total_boxes = outer * bpi;
  PLB_REQUIRE(total_boxes < (1ll << 31), PLB_ESIZE, "plb_gram_tma: K too large");
  PLB_REQUIRE(splits > 0 && splits <= total_boxes && chain_kb > 0, PLB_EINVAL, "plb_gram_tma: bad splits / chain");
  CUtensorMap tmx, tmy;
  if (int rc = make_operand_map(&tmx, x, outer, C, inner, g.tm, box_k)) return rc;
  if (int rc = make_operand_map(&tmy, y, outer, C, inner, g.tn, box_k)) return rc;
  TmaGramParams p;
  p.partial = partial;
  p.qa = row_sumsq_x;
  p.qb = row_sumsq_y;
  p.sa = row_sum_x;
  p.sb = row_sum_y;
  p.C = (int)C;
  p.inner = (int)inner;
  p.boxes_per_image = (int)bpi;
  p.total_boxes = (int)total_boxes;
  p.m_tiles = g.m_tiles;
  p.n_tiles = g.n_tiles;
  p.splits = splits;
  p.total_items = g.m_tiles * g.n_tiles * splits;
  p.ld_m = g.ld_m;
  p.ld_n = g.ld_n;
  p.chain_boxes = chain_kb * 16 / box_k > 0 ? chain_kb * 16 / box_k : 1;  // chain_kb counts 16-wide k-blocks
  static int debug = -1;
  if (debug < 0) {
    const char *d = getenv("PLB_TMA_DEBUG");
    debug = d ? atoi(d) : 0;
  }
  p.debug = debug;
  p.trace = g_trace;
  cudaStream_t s = (cudaStream_t)stream;
  const bool sums = row_sum_x != nullptr;
#define PLB_TMA_DISPATCH(CGV, TMV, TNV)                                                                   \
  do {                                                                                                      \
    if (box_k == 32)                                                                                        \
      return sums ? launch_tma<CGV, TMV, TNV, 32, true>(tmx, tmy, p, s)                                     \
                  : launch_tma<CGV, TMV, TNV, 32, false>(tmx, tmy, p, s);                                   \
    return sums ? launch_tma<CGV, TMV, TNV, 16, true>(tmx, tmy, p, s)                                       \
                : launch_tma<CGV, TMV, TNV, 16, false>(tmx, tmy, p, s);                                     \
  } while (0)
  if (g.cg == 2) PLB_TMA_DISPATCH(2, 128, 128);
  if (g.tm == 64) PLB_TMA_DISPATCH(1, 64, 64);
  PLB_TMA_DISPATCH(1, 128, 128);
#undef PLB_TMA_DISPATCH
}
