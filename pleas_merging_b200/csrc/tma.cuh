// TMA / cluster / cta_group helpers shared by the TMA-fed kernels (gram_tma.cu, conv.cu).
#pragma once

#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace plb {

// ----------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry-point query: no link-time dependency on
// libcuda (the build box has no driver).
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// ----------------------------------------------------------------------------- PTX (cluster / cta_group::2 forms)
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t *smem_result, uint32_t cols) {
  if (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(cols)
                 : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(cols)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish_cg() {
  if (CG == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t cols) {
  if (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_tf32_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier (same shared-memory offset in every CTA of the pair) when all MMAs issued so far retired
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t *bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at this shared-memory offset in the cluster's rank-0 CTA (the MMA leader).  Default
// semantics (release at CTA scope), as CUTLASS's cluster barriers: the data handed over is read by the tensor
// core through the async proxy (fence.proxy.async / tcgen05.fence precede the arrive), not by the waiting
// thread, so no cluster-scope release is needed — that form compiles to MEMBAR.ALL.GPU + ERRBAR per arrive
// and its acquire counterpart to an L1 invalidate per wait, which made every stage hand-over cost ~1 us.
template <int CG>
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
  if (CG == 1) {
    mbar_arrive(bar);
  } else {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
  }
}
// Watchdog waits: a broken handshake traps (the launch fails loudly) instead of hanging the GPU.
constexpr long long kWatchdogClocks = 1ll << 32;  // ~2 s
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > kWatchdogClocks) __trap();
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major operand tile with 128-byte swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor): rows are 128 bytes,
// 8-row groups 1024 bytes apart (SBO), LBO unused for a K extent inside one swizzle row, version 1,
// layout_type 2 = SWIZZLE_128B.  The tile base is 1024-byte aligned; a k8 step advances the start by 32 bytes.
// ROW_BYTES = 64 is the same with the 64-byte swizzle (layout_type 4): 8-row groups 512 bytes apart.
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;
  return d;
}

}  // namespace plb
