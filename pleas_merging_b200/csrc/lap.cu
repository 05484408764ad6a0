// Batched linear sum assignment on the GPU — replaces scipy_solve_lsa
// (pleas/core/solvers.py:18-33), i.e. SciPy's Crouse shortest-augmenting-path solver run in
// float64 on the float32 cost matrix (negated for maximize), behind a D2H copy and sync.
//
// One CTA per problem, all problems of a call in one launch (activation matching solves its
// 37 / 71 groups together; weight matching one launch per level of a sweep's conflict DAG).  The
// algorithm is SciPy's, step for step, so the answer is the same assignment even when optima tie:
//   * rows are inserted in order; each insertion is a Dijkstra search over reduced costs
//     r = ((min + c[i,j]) - u[i]) - v[j] in fp64 with the same association order;
//   * the not-yet-scanned columns live in a list `todo` filled in reverse and compacted by
//     moving its last element into the freed slot;
//   * among columns at the minimum tentative distance an unassigned one wins (the LAST such
//     list position), otherwise the FIRST list position.
// What is parallel: the per-iteration relaxation + arg-min over the todo list, the dual updates
// and the initialisation.  What is sequential: the Dijkstra iterations themselves (38 k of them
// for n = 2048 on N(0,1) costs, 125 k on a ResNet-50 cost matrix) and the final path flip, so
// the latency of ONE iteration is the whole cost.  Two kernels: lap_kernel (v1: one column per
// thread, warp-shuffle (value, rank) min, two barriers per step; kept for A/B runs) and
// lap_kernel_v2 (product path, further down).  All solver state lives in shared memory
// (45 B per column: n <= 4096 fits the 227 KB CTA limit); only the cost row is read from
// global/L2 per iteration.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace plb {

struct LapSmem {  // carved from dynamic shared memory for a given n
  double *u, *v, *dist;
  int *pred, *row4col, *col4row, *todo, *rows_seen, *cols_seen;
};

__device__ __forceinline__ LapSmem carve(uint8_t *base, int n) {
  LapSmem s;
  s.u = (double *)base;
  s.v = s.u + n;
  s.dist = s.v + n;
  s.pred = (int *)(s.dist + n);
  s.row4col = s.pred + n;
  s.col4row = s.row4col + n;
  s.todo = s.col4row + n;
  s.rows_seen = s.todo + n;
  s.cols_seen = s.rows_seen + n;
  return s;
}
static inline size_t lap_smem_bytes(int n) { return (size_t)n * (3 * 8 + 6 * 4) + 64; }

struct Cand {
  double d;
  int rank;  // smaller wins among equal d: sinks get [0,n) by descending list position, others [n,2n) ascending
};
__device__ __forceinline__ bool better(const Cand &a, const Cand &b) {
  return a.d < b.d || (a.d == b.d && a.rank < b.rank);
}
__device__ __forceinline__ Cand shfl_xor(const Cand &c, int m) {
  Cand o;
  o.d = __shfl_xor_sync(0xffffffffu, c.d, m);
  o.rank = __shfl_xor_sync(0xffffffffu, c.rank, m);
  return o;
}

__global__ void __launch_bounds__(1024, 1) lap_kernel(const float *const *__restrict__ costs,
                                                      const int32_t *__restrict__ ns,
                                                      const int32_t *__restrict__ lds, int64_t *const *__restrict__ outs,
                                                      double *__restrict__ objective, int32_t *__restrict__ status,
                                                      int maximize) {
  extern __shared__ __align__(16) uint8_t lap_smem[];
  __shared__ Cand warp_best[32];
  __shared__ double sh_min;
  __shared__ int sh_row, sh_sink, sh_ntodo, sh_nrows, sh_ncols, sh_bad;

  const int prob = blockIdx.x;
  const int n = ns[prob];
  const int64_t ld = lds[prob];
  const float *__restrict__ C = costs[prob];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = nthr >> 5;
  const double sgn = maximize ? -1.0 : 1.0;
  LapSmem s = carve(lap_smem, n);

  if (tid == 0) sh_bad = 0;
  for (int k = tid; k < n; k += nthr) {
    s.u[k] = 0.0;
    s.v[k] = 0.0;
    s.pred[k] = -1;
    s.row4col[k] = -1;
    s.col4row[k] = -1;
  }
  __syncthreads();
  // NaN / -inf screen (SciPy returns an error for those)
  {
    int bad = 0;
    for (int i = 0; i < n; ++i) {
      const float *__restrict__ crow = C + (int64_t)i * ld;
      for (int j = tid; j < n; j += nthr) {
        const double x = sgn * (double)__ldg(crow + j);
        if (x != x || x == -INFINITY) bad = 1;
      }
    }
    if (bad) sh_bad = 1;
  }
  __syncthreads();
  if (sh_bad) {
    if (tid == 0) {
      status[prob] = 2;
      objective[prob] = 0.0;
    }
    return;
  }

  int result = 0;
  for (int cur = 0; cur < n; ++cur) {
    for (int k = tid; k < n; k += nthr) {
      s.dist[k] = INFINITY;
      s.todo[k] = n - 1 - k;
    }
    if (tid == 0) {
      sh_min = 0.0;
      sh_row = cur;
      sh_sink = -1;
      sh_ntodo = n;
      sh_nrows = 0;
      sh_ncols = 0;
    }
    __syncthreads();

    while (true) {
      const int row = sh_row;
      const int ntodo = sh_ntodo;
      const double min_val = sh_min;
      const double u_row = s.u[row];
      const float *__restrict__ crow = C + (int64_t)row * ld;
      Cand best;
      best.d = INFINITY;
      best.rank = 0x7fffffff;
      for (int t = tid; t < ntodo; t += nthr) {
        const int j = s.todo[t];
        const double c = sgn * (double)__ldg(crow + j);
        const double r = __dsub_rn(__dsub_rn(__dadd_rn(min_val, c), u_row), s.v[j]);
        double dj = s.dist[j];
        if (r < dj) {
          s.pred[j] = row;
          s.dist[j] = r;
          dj = r;
        }
        Cand c2;
        c2.d = dj;
        c2.rank = (s.row4col[j] < 0) ? (n - 1 - t) : (n + t);
        if (better(c2, best)) best = c2;
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        const Cand o = shfl_xor(best, m);
        if (better(o, best)) best = o;
      }
      if (lane == 0) warp_best[warp] = best;
      __syncthreads();
      if (warp == 0) {
        Cand b = (lane < nwarps) ? warp_best[lane] : Cand{INFINITY, 0x7fffffff};
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
          const Cand o = shfl_xor(b, m);
          if (better(o, b)) b = o;
        }
        if (lane == 0) {
          if (b.d == INFINITY) {
            sh_sink = -2;  // infeasible
          } else {
            const int t = (b.rank < n) ? (n - 1 - b.rank) : (b.rank - n);
            const int j = s.todo[t];
            s.rows_seen[sh_nrows++] = row;
            s.cols_seen[sh_ncols++] = j;
            sh_min = b.d;
            if (s.row4col[j] < 0) sh_sink = j; else sh_row = s.row4col[j];
            s.todo[t] = s.todo[ntodo - 1];
            sh_ntodo = ntodo - 1;
          }
        }
      }
      __syncthreads();
      if (sh_sink != -1) break;
    }
    if (sh_sink == -2) {
      result = 1;
      break;
    }

    // dual update (rows_seen[0] == cur; row k>0 was reached through column cols_seen[k-1])
    const double min_val = sh_min;
    const int nrows = sh_nrows, ncols = sh_ncols;
    for (int k = tid; k < nrows; k += nthr) {
      const int i = s.rows_seen[k];
      if (k == 0) s.u[i] = __dadd_rn(s.u[i], min_val);
      else s.u[i] = __dadd_rn(s.u[i], __dsub_rn(min_val, s.dist[s.col4row[i]]));
    }
    for (int k = tid; k < ncols; k += nthr) {
      const int j = s.cols_seen[k];
      s.v[j] = __dsub_rn(s.v[j], __dsub_rn(min_val, s.dist[j]));
    }
    __syncthreads();
    if (tid == 0) {  // flip the augmenting path
      int j = sh_sink;
      while (true) {
        const int i = s.pred[j];
        s.row4col[j] = i;
        const int prev = s.col4row[i];
        s.col4row[i] = j;
        j = prev;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }

  if (result != 0) {
    if (tid == 0) {
      status[prob] = result;
      objective[prob] = 0.0;
    }
    return;
  }
  // outputs: int64 column per row + fp64 objective on the caller's (un-negated) costs
  int64_t *out = outs[prob];
  double part = 0.0;
  for (int i = tid; i < n; i += nthr) {
    const int j = s.col4row[i];
    out[i] = (int64_t)j;
    part += (double)C[(int64_t)i * ld + j];
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
  if (lane == 0) warp_best[warp].d = part;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int wq = 0; wq < nwarps; ++wq) tot += warp_best[wq].d;
    objective[prob] = tot;
    status[prob] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// v2: the same algorithm (same selection rules, same fp64 association order => the same
// assignment, ties included) with a shorter critical path per Dijkstra step.  The solve is a
// chain of ~60 n dependent steps on the hard instances of the merge path (ResNet-50, n = 2048:
// 125 k steps), so the step LATENCY is the whole cost.  Measured on the v1 kernel: 0.9 us per
// step with 8-16 warps, 2.1 us with 32.  Changes:
//   * CPT columns per thread, compile-time, processed in phases (all list reads, then all cost
//     loads, then all relaxations) so the loads of a thread overlap: few warps, no 32-warp
//     barrier, no 32-way second reduction stage;
//   * arg-min by redux.sync on an order-preserving integer image of (distance, rank): three
//     single-instruction reductions instead of a 5-level 3-register shuffle tree;
//   * one barrier per step: warp winners (with their column) go to a double-buffered shared
//     slot, every thread reduces them redundantly and tracks (row, list length, minimum,
//     seen counts) in registers; the owner of the winning list slot compacts the list, so no
//     thread ever reads a slot another thread is rewriting.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t order_key(double d) {
  const uint64_t b = (uint64_t)__double_as_longlong(d);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(uint64_t k) {
  const uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct Win {
  uint64_t key;
  uint32_t rank;
  bool mine;  // this lane holds the winning candidate
};
// lexicographic min of (key, rank) over the lanes of a warp; rank is unique per candidate
__device__ __forceinline__ Win warp_argmin(uint64_t key, uint32_t rank) {
  const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
  const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
  const uint32_t mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  const bool tie = hi == mhi && lo == mlo;
  const uint32_t mrk = __reduce_min_sync(0xffffffffu, tie ? rank : 0xffffffffu);
  Win w;
  w.key = ((uint64_t)mhi << 32) | mlo;
  w.rank = mrk;
  w.mine = tie && rank == mrk;
  return w;
}

template <int CPT>
__global__ void __launch_bounds__(1024, 1) lap_kernel_v2(const float *const *__restrict__ costs,
                                                        const int32_t *__restrict__ ns,
                                                        const int32_t *__restrict__ lds,
                                                        int64_t *const *__restrict__ outs,
                                                        double *__restrict__ objective, int32_t *__restrict__ status,
                                                        int maximize, const double *const *__restrict__ v_in,
                                                        double v_scale, double *const *__restrict__ v_out) {
  extern __shared__ __align__(16) uint8_t lap_smem[];
  __shared__ uint64_t w_key[2][32];
  __shared__ uint32_t w_rank[2][32];
  __shared__ int w_col[2][32];
  __shared__ double red[32];
  __shared__ int sh_bad;

  const int prob = blockIdx.x;
  const int n = ns[prob];
  const int64_t ld = lds[prob];
  const float *__restrict__ C = costs[prob];
  // The block is sized for the LARGEST problem of the batch; a smaller problem keeps only the warps its own
  // columns need (CPT columns per thread) and retires the rest before the first barrier: every Dijkstra step
  // pays one block barrier and a cross-warp reduction, so a 64-unit group behind a 2048-unit one would
  // otherwise synchronise 32 warps per step instead of 1.
  const int need = max(32, (((n + CPT - 1) / CPT + 31) / 32) * 32);
  const int tid = threadIdx.x, nthr = min((int)blockDim.x, need), lane = tid & 31, warp = tid >> 5;
  if (tid >= nthr) return;
  const int nwarps = nthr >> 5;
  const double sgn = maximize ? -1.0 : 1.0;
  LapSmem s = carve(lap_smem, n);

  if (tid == 0) sh_bad = 0;
  for (int k = tid; k < n; k += nthr) {
    s.u[k] = 0.0;
    s.v[k] = 0.0;
    s.pred[k] = -1;
    s.row4col[k] = -1;
    s.col4row[k] = -1;
  }
  __syncthreads();
  {  // NaN / -inf screen (SciPy returns an error for those)
    int bad = 0;
    for (int i = 0; i < n; ++i) {
      const float *__restrict__ crow = C + (int64_t)i * ld;
      for (int j = tid; j < n; j += nthr) {
        const double x = sgn * (double)__ldg(crow + j);
        if (x != x || x == -INFINITY) bad = 1;
      }
    }
    if (bad) sh_bad = 1;
  }
  __syncthreads();
  if (sh_bad) {
    if (tid == 0) {
      status[prob] = 2;
      objective[prob] = 0.0;
    }
    return;
  }

  // Warm start (optional): column duals of a related problem (weight matching re-solves every group once per
  // sweep on slowly changing costs).  Any v gives feasible duals with u_i = min_j (c_ij - v_j); a row keeps its
  // arg-min column when no smaller row claimed it (reduced cost 0 on the matched edge), the other rows stay free
  // and are inserted by the search below, which only needs dual feasibility and complementary slackness.  The
  // result is an optimal assignment — identical to the cold start's whenever the optimum is unique.
  const double *vin = (v_in != nullptr) ? v_in[prob] : nullptr;
  if (vin != nullptr) {
    for (int k = tid; k < n; k += nthr) s.v[k] = v_scale * vin[k];
    __syncthreads();
    for (int i = warp; i < n; i += nwarps) {
      const float *__restrict__ crow = C + (int64_t)i * ld;
      double best = INFINITY;
      int bj = n;
      for (int j = lane; j < n; j += 32) {
        const double r = __dsub_rn(sgn * (double)__ldg(crow + j), s.v[j]);
        if (r < best) {
          best = r;
          bj = j;
        }
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, m);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, m);
        if (ob < best || (ob == best && oj < bj)) {
          best = ob;
          bj = oj;
        }
      }
      if (lane == 0) {
        s.u[i] = best;
        s.pred[i] = bj;
      }
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i < n; ++i) {
        const int j = s.pred[i];
        if (j < n && s.row4col[j] < 0) {
          s.row4col[j] = i;
          s.col4row[i] = j;
        }
      }
    }
    __syncthreads();
    for (int k = tid; k < n; k += nthr) s.pred[k] = -1;
    __syncthreads();
  }

  int result = 0;
  uint32_t step = 0;  // parity selects the warp-winner buffer
  for (int cur = 0; cur < n; ++cur) {
    if (s.col4row[cur] >= 0) continue;  // matched by the warm start (uniform: shared memory, read after a barrier)
    for (int k = tid; k < n; k += nthr) {
      s.dist[k] = INFINITY;
      s.todo[k] = n - 1 - k;
    }
    __syncthreads();
    int row = cur, ntodo = n, nrows = 0, ncols = 0, sink = -1;
    double min_val = 0.0;

    while (true) {
      const double u_row = s.u[row];
      const float *__restrict__ crow = C + (int64_t)row * ld;
      int jj[CPT];
      float cf[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int t = tid + c * nthr;
        jj[c] = (t < ntodo) ? s.todo[t] : -1;
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c) cf[c] = (jj[c] >= 0) ? __ldg(crow + jj[c]) : 0.f;
      uint64_t bkey = 0xffffffffffffffffull;
      uint32_t brank = 0xffffffffu;
      int bcol = -1;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (jj[c] >= 0) {
          const int j = jj[c], t = tid + c * nthr;
          const double cst = sgn * (double)cf[c];
          const double r = __dsub_rn(__dsub_rn(__dadd_rn(min_val, cst), u_row), s.v[j]);
          double dj = s.dist[j];
          if (r < dj) {
            s.pred[j] = row;
            s.dist[j] = r;
            dj = r;
          }
          const uint64_t key = order_key(dj);
          const uint32_t rank = (s.row4col[j] < 0) ? (uint32_t)(n - 1 - t) : (uint32_t)(n + t);
          if (key < bkey || (key == bkey && rank < brank)) {
            bkey = key;
            brank = rank;
            bcol = j;
          }
        }
      }
      const uint32_t buf = step & 1u;
      ++step;
      {
        const Win w = warp_argmin(bkey, brank);
        if (w.mine && brank != 0xffffffffu) {
          w_key[buf][warp] = w.key;
          w_rank[buf][warp] = w.rank;
          w_col[buf][warp] = bcol;
        } else if (lane == 0 && w.rank == 0xffffffffu) {  // no candidate in this warp
          w_key[buf][warp] = 0xffffffffffffffffull;
          w_rank[buf][warp] = 0xffffffffu;
          w_col[buf][warp] = -1;
        }
      }
      __syncthreads();
      const bool have = lane < nwarps;
      const Win g = warp_argmin(have ? w_key[buf][lane] : 0xffffffffffffffffull, have ? w_rank[buf][lane] : 0xffffffffu);
      if (g.rank == 0xffffffffu || key_value(g.key) == INFINITY) {
        sink = -2;  // infeasible
        break;
      }
      const int wsrc = __ffs(__ballot_sync(0xffffffffu, g.mine)) - 1;
      const int j = w_col[buf][wsrc];
      const int t = (g.rank < (uint32_t)n) ? (n - 1 - (int)g.rank) : ((int)g.rank - n);
      min_val = key_value(g.key);
      if (tid == 0) {
        s.rows_seen[nrows] = row;
        s.cols_seen[ncols] = j;
      }
      ++nrows;
      ++ncols;
#pragma unroll
      for (int c = 0; c < CPT; ++c)  // the slot's own reader compacts it
        if (tid + c * nthr == t) s.todo[t] = s.todo[ntodo - 1];
      --ntodo;
      const int r4c = s.row4col[j];
      if (r4c < 0) {
        sink = j;
        break;
      }
      row = r4c;
    }
    __syncthreads();  // rows_seen / cols_seen / dist / pred of the last step are visible
    if (sink == -2) {
      result = 1;
      break;
    }

    // dual update (rows_seen[0] == cur; row k>0 was reached through column cols_seen[k-1])
    for (int k = tid; k < nrows; k += nthr) {
      const int i = s.rows_seen[k];
      if (k == 0) s.u[i] = __dadd_rn(s.u[i], min_val);
      else s.u[i] = __dadd_rn(s.u[i], __dsub_rn(min_val, s.dist[s.col4row[i]]));
    }
    for (int k = tid; k < ncols; k += nthr) {
      const int j = s.cols_seen[k];
      s.v[j] = __dsub_rn(s.v[j], __dsub_rn(min_val, s.dist[j]));
    }
    __syncthreads();
    if (tid == 0) {  // flip the augmenting path
      int j = sink;
      while (true) {
        const int i = s.pred[j];
        s.row4col[j] = i;
        const int prev = s.col4row[i];
        s.col4row[i] = j;
        j = prev;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }

  if (result != 0) {
    if (tid == 0) {
      status[prob] = result;
      objective[prob] = 0.0;
    }
    return;
  }
  if (v_out != nullptr && v_out[prob] != nullptr)
    for (int k = tid; k < n; k += nthr) v_out[prob][k] = s.v[k];
  int64_t *out = outs[prob];
  double part = 0.0;
  for (int i = tid; i < n; i += nthr) {
    const int j = s.col4row[i];
    out[i] = (int64_t)j;
    part += (double)C[(int64_t)i * ld + j];
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int wq = 0; wq < nwarps; ++wq) tot += red[wq];
    objective[prob] = tot;
    status[prob] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// v3 (product path): the same algorithm and selection rules again, with the per-column solver
// state in REGISTERS.  v2 keeps v, dist, row4col and the todo list in shared memory and gathers
// four of them per column per step (49 KB of shared-memory traffic per step at n = 2048 — the
// step was bound by the shared-memory pipe, not by the L2 latency of the cost row).  Here a
// thread owns fixed columns j = tid + c * nthr: its column dual, tentative distance and
// row4col live in registers, and SciPy's `remaining` list is kept implicitly as the list
// POSITION of each owned column (the list is filled in reverse: column j starts at position
// n-1-j; removing position t moves the column at the last position into t — the owner of that
// column notices `pos == last` and takes t).  The cost row is read coalesced (fixed columns), a
// step touches shared memory only for u[row], the predecessor store and the 32 warp winners.
// The dual update runs from registers (the owner of a scanned column j updates v_j and the
// dual of the row matched to j); only the path flip walks shared memory.
// Shared memory: u, a column-dual scratch for the warm start, pred, row4col, col4row (28 B/col).
// ------------------------------------------------------------------------------------------
#ifdef PLB_LAP_TRACE  // profiles/experiments/lap_trace.cu: per-phase clock sums of problem 0 (first and last thread)
__device__ uint32_t plb_lap_trace[2][10];
#define LAP_STAMP(k, dep)                                                                                      \
  do { /* the volatile store consumes `dep` first: in-order issue makes the clock read wait for the value */   \
    uint32_t _t;                                                                                               \
    asm volatile("st.volatile.shared.u32 [%1], %2;\n\tmov.u32 %0, %%clock;"                                     \
                 : "=r"(_t)                                                                                    \
                 : "r"(sm_base + 1028u + 0u * (uint32_t)(k)), "r"((uint32_t)(dep))                             \
                 : "memory");                                                                                  \
    tr_acc[k] += _t - tr_last;                                                                                 \
    tr_last = _t;                                                                                              \
  } while (0)
#else
#define LAP_STAMP(k, dep) \
  do {                    \
  } while (0)
#endif
static inline size_t lap_v3_smem_bytes(int n) { return 1088 + (size_t)n * (2 * 8 + 3 * 4) + 64; }

// order_key on the two 32-bit halves (b ^ ((b >> 63) | sign bit)): pure integer instructions — the compiler
// turns the 64-bit form into an fp64 -|x| DADD on the long-latency pipe.
__device__ __forceinline__ uint64_t order_key_i(double d) {
  const int hi = __double2hiint(d);
  const uint32_t lo = (uint32_t)__double2loint(d);
  const uint32_t m = (uint32_t)(hi >> 31);
  return ((uint64_t)((uint32_t)hi ^ (m | 0x80000000u)) << 32) | (uint64_t)(lo ^ m);
}
// Shared-memory accesses of the step loop by 32-bit shared address held in a register: with C++ pointers the
// compiler re-derives the shared window base (S2UR SR_CgaCtaId + ULEA) inside the dependent chain of every step.
__device__ __forceinline__ uint64_t lds_u64(uint32_t a) {
  uint64_t v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_slot(uint32_t a, uint64_t key, uint32_t word) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"((uint32_t)key), "r"((uint32_t)(key >> 32)),
               "r"(word), "r"(0u)
               : "memory");
}
__device__ __forceinline__ void lds_slot(uint32_t a, uint64_t &key, uint32_t &word) {
  uint32_t lo, hi;
  [[maybe_unused]] uint32_t pad;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(hi), "=r"(word), "=r"(pad) : "r"(a) : "memory");
  key = ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void sts_u64(uint32_t a, uint64_t v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

template <int CPT>
__global__ void __launch_bounds__(1024, 1) lap_kernel_v3(const float *const *__restrict__ costs,
                                                        const int32_t *__restrict__ ns,
                                                        const int32_t *__restrict__ lds,
                                                        int64_t *const *__restrict__ outs,
                                                        double *__restrict__ objective, int32_t *__restrict__ status,
                                                        int maximize, const double *const *__restrict__ v_in,
                                                        double v_scale, double *const *__restrict__ v_out) {
  extern __shared__ __align__(16) uint8_t lap_smem[];
  // warp-winner slots (double-buffered by step parity) at the front of the dynamic allocation: [2][32] slots of
  // 16 B {u64 distance key, u32 rank << 13 | (row4col + 1), pad} — one 128-bit store / load per step and thread;
  // sink column of the finished search at +1024.  The base goes through an opaque asm so that it stays a register.
  uint32_t sm_base;
  asm volatile("mov.u32 %0, %1;" : "=r"(sm_base) : "r"(smem_u32(lap_smem)));
  __shared__ double red[32];
  __shared__ int sh_bad;

  const int prob = blockIdx.x;
  const int n = ns[prob];
  const int64_t ld = lds[prob];
  const float *__restrict__ C = costs[prob];
  const int need = max(32, (((n + CPT - 1) / CPT + 31) / 32) * 32);
  const int tid = threadIdx.x, nthr = min((int)blockDim.x, need), lane = tid & 31, warp = tid >> 5;
  if (tid >= nthr) return;  // warps this problem's columns do not need (block sized for the batch's largest n)
  const int nwarps = nthr >> 5;
  const double sgn = maximize ? -1.0 : 1.0;
  const uint32_t smask = maximize ? 0x80000000u : 0u;  // costs are negated by flipping the float's sign bit
  double *__restrict__ u = (double *)(lap_smem + 1088);
  double *__restrict__ vtmp = u + n;
  int *__restrict__ pred = (int *)(vtmp + n);
  int *__restrict__ row4col = pred + n;
  int *__restrict__ col4row = row4col + n;

  if (tid == 0) sh_bad = 0;
  for (int k = tid; k < n; k += nthr) {
    u[k] = 0.0;
    pred[k] = -1;
    row4col[k] = -1;
    col4row[k] = -1;
  }
  __syncthreads();
  {  // NaN / -inf screen (SciPy returns an error for those); rows of 4 GB and more are not addressed (32-bit
    // row offsets in the step loop) and reported the same way
    int bad = (ld >= (int64_t)(1 << 30)) ? 1 : 0;
    for (int i = 0; i < n; ++i) {
      const float *__restrict__ crow = C + (int64_t)i * ld;
      for (int j = tid; j < n; j += nthr) {
        const double x = sgn * (double)__ldg(crow + j);
        if (x != x || x == -INFINITY) bad = 1;
      }
    }
    if (bad) sh_bad = 1;
  }
  __syncthreads();
  if (sh_bad) {
    if (tid == 0) {
      status[prob] = 2;
      objective[prob] = 0.0;
    }
    return;
  }

  // Tentative distances are kept as their order-preserving integer image: fp64 compares sit on a long-latency pipe
  // (a DSETP costs as much as a DADD), the integer compares of the relaxation and of the arg-min do not.  The
  // image is exact and monotonic for every value that occurs (no NaN after the screen, and -0.0 cannot arise:
  // min_val starts at +0.0 and x - x rounds to +0.0).
  double v[CPT];
  uint64_t dkey[CPT];
  int pos[CPT], r4c[CPT];  // pos: list position; -1 = scanned in this search; -2 = no such column
  // candidate word rank << 13 | (row4col + 1) = kb + ks * pos: an unassigned column ranks n-1-pos (the LAST list
  // position wins), an assigned one n+pos (the first wins, after every unassigned column)
  uint32_t kb[CPT];
  int ks[CPT];
  const float *Cj[CPT];  // &C[0][j] of the owned columns: the cost address is one IMAD.WIDE per column
#pragma unroll
  for (int c = 0; c < CPT; ++c) v[c] = 0.0;

  // Warm start (optional), as in v2: any column duals v are feasible with u_i = min_j (c_ij - v_j); a row keeps
  // its arg-min column when no smaller row claimed it, the others are inserted by the search below.
  const double *vin = (v_in != nullptr) ? v_in[prob] : nullptr;
  if (vin != nullptr) {
    for (int k = tid; k < n; k += nthr) vtmp[k] = v_scale * vin[k];
    __syncthreads();
    for (int i = warp; i < n; i += nwarps) {
      const float *__restrict__ crow = C + (int64_t)i * ld;
      double best = INFINITY;
      int bj = n;
      for (int j = lane; j < n; j += 32) {
        const double r = __dsub_rn(sgn * (double)__ldg(crow + j), vtmp[j]);
        if (r < best) {
          best = r;
          bj = j;
        }
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, m);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, m);
        if (ob < best || (ob == best && oj < bj)) {
          best = ob;
          bj = oj;
        }
      }
      if (lane == 0) {
        u[i] = best;
        pred[i] = bj;
      }
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i < n; ++i) {
        const int j = pred[i];
        if (j < n && row4col[j] < 0) {
          row4col[j] = i;
          col4row[i] = j;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = tid + c * nthr;
      if (j < n) v[c] = vtmp[j];
    }
  }
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int j = tid + c * nthr;
    r4c[c] = (j < n) ? row4col[j] : -1;
    kb[c] = (r4c[c] < 0) ? ((uint32_t)(n - 1) << 13) : (((uint32_t)n << 13) | (uint32_t)(r4c[c] + 1));
    ks[c] = (r4c[c] < 0) ? -8192 : 8192;
    Cj[c] = C + (j < n ? j : 0);
  }
  const uint32_t ld4 = (uint32_t)ld * 4u;  // bytes per cost row (the host checks ld < 2^30)

  const uint32_t sm_u = sm_base + 1088u, sm_pred = sm_u + 16u * (uint32_t)n;
  uint32_t sm_wslot = sm_base + 16u * (uint32_t)warp, sm_gslot = sm_base + 16u * (uint32_t)lane;  // ^= 512 per step
  uint32_t pa[CPT <= 2 ? CPT : 1];  // shared address of pred[j] of the owned columns (recomputed when CPT > 2:
#pragma unroll                      // registers)
  for (int c = 0; c < (CPT <= 2 ? CPT : 1); ++c) pa[c] = sm_pred + 4u * (uint32_t)(tid + c * nthr);
  int result = 0;
  uint32_t step = 0;  // parity selects the warp-winner buffer
#ifdef PLB_LAP_TRACE
  uint32_t tr_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tr_last;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(tr_last)::"memory");
#endif
  for (int cur = 0; cur < n; ++cur) {
    if (col4row[cur] >= 0) continue;  // matched by the warm start (uniform: written only between barriers)
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = tid + c * nthr;
      dkey[c] = 0xfff0000000000000ull;  // +inf
      pos[c] = (j < n) ? (n - 1 - j) : -2;
    }
    int row = cur, ntodo = n, sink = -1;
    double min_val = 0.0;
    LAP_STAMP(7, ntodo);
    // The cost row and u[row] of a step are requested as soon as the row is known — for the next step right after
    // the cross-warp arg-min, ahead of the list bookkeeping — and unconditionally (a scanned or padding column
    // reads a valid address and ignores the value): the L2 round trip overlaps the tail of the previous step.
    double u_row = lds_f64(sm_u + 8u * (uint32_t)row);
    float cf[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) cf[c] = __ldg((const float *)((const char *)Cj[c] + (uint64_t)(uint32_t)row * ld4));

    while (true) {
      LAP_STAMP(0, __float_as_uint(cf[CPT - 1]) ^ __float_as_uint(cf[0]));
      // candidate = (distance key, rank << 13 | row4col + 1): rank < 2n <= 8192 is unique, so the third reduction
      // word also carries the row matched to the winning column
      uint64_t bkey = 0xffffffffffffffffull;
      uint32_t bpk = 0xffffffffu;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (pos[c] >= 0) {
          const double cst = (double)__uint_as_float(__float_as_uint(cf[c]) ^ smask);
          const double r = __dsub_rn(__dsub_rn(__dadd_rn(min_val, cst), u_row), v[c]);
          const uint64_t rk = order_key_i(r);
          if (rk < dkey[c]) {
            dkey[c] = rk;
            sts_u32(CPT <= 2 ? pa[CPT <= 2 ? c : 0] : sm_pred + 4u * (uint32_t)(tid + c * nthr), (uint32_t)row);
          }
          const uint32_t pk = kb[c] + (uint32_t)(ks[c] * pos[c]);
          if (c == 0 || dkey[c] < bkey || (dkey[c] == bkey && pk < bpk)) {
            bkey = dkey[c];
            bpk = pk;
          }
        }
      }
      ++step;
      LAP_STAMP(1, (uint32_t)bkey);
      {
        const Win w = warp_argmin(bkey, bpk);
        LAP_STAMP(2, w.rank);
        if (lane == 0) sts_slot(sm_wslot, w.key, w.rank);  // all-ones when the warp has no candidate
      }
      __syncthreads();
      LAP_STAMP(3, 0);
      const bool have = lane < nwarps;
      uint64_t gk = 0xffffffffffffffffull;
      uint32_t gr = 0xffffffffu;
      if (have) lds_slot(sm_gslot, gk, gr);
      sm_wslot ^= 512u;
      sm_gslot ^= 512u;
      LAP_STAMP(4, gr ^ (uint32_t)gk);
      const Win g = warp_argmin(gk, gr);
      LAP_STAMP(5, g.rank);
      if (g.key >= 0xfff0000000000000ull) {  // no candidate, or the nearest column is at +inf: infeasible
        sink = -2;
        break;
      }
      const uint32_t grank = g.rank >> 13;
      const int next_row = (int)(g.rank & 0x1fffu) - 1;
      const uint32_t row_req = (uint32_t)(next_row < 0 ? row : next_row);  // the last step re-requests its own row
      const double u_next = lds_f64(sm_u + 8u * row_req);
      float cfn[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) cfn[c] = __ldg((const float *)((const char *)Cj[c] + (uint64_t)row_req * ld4));
      const int t = (grank < (uint32_t)n) ? (n - 1 - (int)grank) : ((int)grank - n);
      min_val = key_value(g.key);
      const int last = ntodo - 1;
#pragma unroll
      for (int c = 0; c < CPT; ++c)  // scanned; SciPy moves the last list element into the freed slot
        pos[c] = (pos[c] == t) ? -1 : ((pos[c] == last) ? t : pos[c]);
      ntodo = last;
      LAP_STAMP(6, next_row ^ pos[0]);
      if (next_row < 0) {
        sink = 0;
        break;
      }
      row = next_row;
      u_row = u_next;
#pragma unroll
      for (int c = 0; c < CPT; ++c) cf[c] = cfn[c];
    }
    if (sink == -2) {
      result = 1;
      break;
    }

    // dual update from registers: the owner of a scanned column j updates v_j and the dual of the row matched to
    // j (that row was reached through j: col4row[row] == j before the flip); the inserted row gets + min_val.
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (pos[c] == -1 && r4c[c] < 0) sts_u32(sm_base + 1024u, (uint32_t)(tid + c * nthr));  // the sink: the only
      if (pos[c] == -1) {                                                        // scanned unassigned column
        const double dv = __dsub_rn(min_val, key_value(dkey[c]));
        v[c] = __dsub_rn(v[c], dv);
        if (r4c[c] >= 0) u[r4c[c]] = __dadd_rn(u[r4c[c]], dv);
      }
    }
    if (tid == 0) u[cur] = __dadd_rn(u[cur], min_val);
    __syncthreads();  // pred / sh_sink of the last step are visible
    if (tid == 0) {   // flip the augmenting path
      int j = (int)lds_u32(sm_base + 1024u);
      while (true) {
        const int i = pred[j];
        row4col[j] = i;
        const int prev = col4row[i];
        col4row[i] = j;
        j = prev;
        if (i == cur) break;
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = tid + c * nthr;
      if (j < n) {
        r4c[c] = row4col[j];
        kb[c] = (r4c[c] < 0) ? ((uint32_t)(n - 1) << 13) : (((uint32_t)n << 13) | (uint32_t)(r4c[c] + 1));
        ks[c] = (r4c[c] < 0) ? -8192 : 8192;
      }
    }
  }

#ifdef PLB_LAP_TRACE
  if (prob == 0 && (tid == 0 || tid == nthr - 1)) {
    tr_acc[8] = step;
    for (int k = 0; k < 10; ++k) plb_lap_trace[tid == 0 ? 0 : 1][k] = tr_acc[k];
  }
#endif
  if (result != 0) {
    if (tid == 0) {
      status[prob] = result;
      objective[prob] = 0.0;
    }
    return;
  }
  if (v_out != nullptr && v_out[prob] != nullptr) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = tid + c * nthr;
      if (j < n) v_out[prob][j] = v[c];
    }
  }
  int64_t *out = outs[prob];
  double part = 0.0;
  for (int i = tid; i < n; i += nthr) {
    const int j = col4row[i];
    out[i] = (int64_t)j;
    part += (double)C[(int64_t)i * ld + j];
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int wq = 0; wq < nwarps; ++wq) tot += red[wq];
    objective[prob] = tot;
    status[prob] = 0;
  }
}

template <int CPT>
static int launch_lap_v3(const float *const *cost, const int32_t *n, const int32_t *ld, int64_t *const *col4row,
                         double *objective, int32_t *status, int n_problems, int max_n, int maximize,
                         const double *const *v_in, double v_scale, double *const *v_out, cudaStream_t stream) {
  const size_t smem = lap_v3_smem_bytes(max_n);
  if (int rc = ensure_dynamic_smem((const void *)lap_kernel_v3<CPT>, (int)smem, "lap_kernel_v3")) return rc;
  int threads = (int)(ceil_div(max_n, 32 * CPT) * 32);
  threads = threads < 32 ? 32 : threads;  // <= 1024 by the caller's choice of CPT
  lap_kernel_v3<CPT><<<n_problems, threads, smem, stream>>>(cost, n, ld, col4row, objective, status, maximize, v_in,
                                                            v_scale, v_out);
  return launch_status("lap_kernel_v3");
}

template <int CPT>
static int launch_lap_v2(const float *const *cost, const int32_t *n, const int32_t *ld, int64_t *const *col4row,
                         double *objective, int32_t *status, int n_problems, int max_n, int maximize, size_t smem,
                         const double *const *v_in, double v_scale, double *const *v_out, cudaStream_t stream) {
  if (int rc = ensure_dynamic_smem((const void *)lap_kernel_v2<CPT>, (int)smem, "lap_kernel_v2")) return rc;
  int threads = (int)(ceil_div(max_n, 32 * CPT) * 32);
  threads = threads < 32 ? 32 : threads;  // <= 1024 by the caller's choice of CPT
  lap_kernel_v2<CPT><<<n_problems, threads, smem, stream>>>(cost, n, ld, col4row, objective, status, maximize, v_in,
                                                            v_scale, v_out);
  return launch_status("lap_kernel_v2");
}

}  // namespace plb

static int lap_solve_impl(const float *const *cost, const int32_t *n, const int32_t *ld, int64_t *const *col4row,
                          double *objective, int32_t *status, int32_t n_problems, int32_t max_n, int32_t maximize,
                          const double *const *v_in, double v_scale, double *const *v_out, void *stream) {
  using namespace plb;
  PLB_REQUIRE(cost && n && ld && col4row && objective && status, PLB_EINVAL, "plb_lap_solve_batched: null pointer");
  PLB_REQUIRE(n_problems > 0 && max_n > 0, PLB_EINVAL, "plb_lap_solve_batched: empty batch");
  PLB_REQUIRE(max_n <= 4096, PLB_ESIZE, "plb_lap_solve_batched: n > 4096 exceeds the shared-memory working set");
  const size_t smem = lap_smem_bytes(max_n);
  if (int rc = ensure_dynamic_smem((const void *)lap_kernel, (int)smem, "lap_kernel")) return rc;
  // PLB_LAP_IMPL=v1 / v2 select the earlier kernels (v1: one column per thread, two barriers per step; v2: solver
  // state in shared memory) for A/B runs; PLB_LAP_COLS_PER_THREAD overrides the columns per thread.
  static int impl = 0, cols_override = -1;
  if (impl == 0) {
    const char *e = getenv("PLB_LAP_IMPL");
    impl = (e && e[0] == 'v' && e[1] == '1') ? 1 : ((e && e[0] == 'v' && e[1] == '2') ? 2 : 3);
    const char *c = getenv("PLB_LAP_COLS_PER_THREAD");
    cols_override = c ? atoi(c) : 0;
  }
  if (impl == 3) {
    // register-resident column state (product path); columns per thread as for v2, PLB_LAP_COLS_PER_THREAD overrides
    int cpt = cols_override > 0 ? cols_override : (max_n <= 512 ? 1 : (max_n <= 2048 ? 2 : 4));
    while (cpt < 8 && (int64_t)cpt * 1024 < max_n) cpt *= 2;
    cudaStream_t st = (cudaStream_t)stream;
    switch (cpt) {
      case 1: return launch_lap_v3<1>(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, v_in,
                                       v_scale, v_out, st);
      case 2: return launch_lap_v3<2>(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, v_in,
                                       v_scale, v_out, st);
      case 4: return launch_lap_v3<4>(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, v_in,
                                       v_scale, v_out, st);
      default: return launch_lap_v3<8>(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, v_in,
                                       v_scale, v_out, st);
    }
  }
  if (impl == 2) {
    // columns per thread, measured on B200 (profiles/experiments/lap_cpt_sweep.py, N(0,1) costs, ms at
    // 1 / 2 / 4 / 8 columns): n=256 1.7 / 2.0 / 2.7 / 4.2, n=512 5.1 / 5.5 / 6.9 / 11.3, n=1024 14.9 / 13.7 /
    // 15.6 / 24.6, n=2048 - / 52.2 / 51.2 / 71.5, n=4096 - / - / 202 / 233 (v1 kernel: 3.2, 9.2, 22.6, 74.4, 263)
    int cpt = cols_override > 0 ? cols_override : (max_n <= 512 ? 1 : (max_n <= 2048 ? 2 : 4));
    while (cpt < 8 && (int64_t)cpt * 1024 < max_n) cpt *= 2;
    cudaStream_t st = (cudaStream_t)stream;
    switch (cpt) {
      case 1: return launch_lap_v2<1>(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, smem, v_in,
                                       v_scale, v_out, st);
      case 2: return launch_lap_v2<2>(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, smem, v_in,
                                       v_scale, v_out, st);
      case 4: return launch_lap_v2<4>(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, smem, v_in,
                                       v_scale, v_out, st);
      default: return launch_lap_v2<8>(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, smem, v_in,
                                       v_scale, v_out, st);
    }
  }
  PLB_REQUIRE(v_in == nullptr && v_out == nullptr, PLB_EINVAL,
              "plb_lap_solve_batched_warm: the warm start needs the v2 kernel (unset PLB_LAP_IMPL=v1)");
  const int cols_per_thread = cols_override > 0 ? cols_override : 1;
  int threads = (int)(ceil_div(max_n, 32 * cols_per_thread) * 32);
  threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
  lap_kernel<<<n_problems, threads, smem, (cudaStream_t)stream>>>(cost, n, ld, col4row, objective, status, maximize);
  return launch_status("lap_kernel");
}

extern "C" int plb_lap_solve_batched(const float *const *cost, const int32_t *n, const int32_t *ld,
                                     int64_t *const *col4row, double *objective, int32_t *status,
                                     int32_t n_problems, int32_t max_n, int32_t maximize, void *stream) {
  return lap_solve_impl(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, nullptr, 1.0, nullptr,
                        stream);
}

extern "C" int plb_lap_solve_batched_warm(const float *const *cost, const int32_t *n, const int32_t *ld,
                                          int64_t *const *col4row, double *objective, int32_t *status,
                                          int32_t n_problems, int32_t max_n, int32_t maximize,
                                          const double *const *v_in, double v_scale, double *const *v_out,
                                          void *stream) {
  return lap_solve_impl(cost, n, ld, col4row, objective, status, n_problems, max_n, maximize, v_in, v_scale, v_out,
                        stream);
}
