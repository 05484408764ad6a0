// Latency / throughput of the instructions on the LAP step's dependent chain (experiments only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 profiles/experiments/fp64_lat.cu -o /tmp/fp64_lat && /tmp/fp64_lat
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int OP>
__global__ void chain(double *out, uint32_t *clk, double a, double b, float f, int iters) {
  double x = a + threadIdx.x;
  float g = f + threadIdx.x;
  uint32_t k = threadIdx.x;
  uint32_t t0, t1;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(t0)::"memory");
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) x = __dadd_rn(x, b);                                      // DADD chain
      if (OP == 1) { g = (float)((double)g + b); }                          // F2F.F64.F32 + DADD + F2F.F32.F64
      if (OP == 2) k = __reduce_min_sync(0xffffffffu, k + u) + 1;            // CREDUX chain
      if (OP == 3) { x = __dadd_rn(x, b); if (x < a) x = a; }                // DADD + DSETP + select
      if (OP == 4) k = __shfl_xor_sync(0xffffffffu, k, 1) + u;               // SHFL chain
      if (OP == 5) k = __popc(__ballot_sync(0xffffffffu, k & 1)) + k;        // VOTE chain
    }
  }
  asm volatile("mov.u32 %0, %%clock;" : "=r"(t1)::"memory");
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + g + k;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char *name, int threads) {
  double *out;
  uint32_t *clk, h;
  cudaMalloc(&out, 8 * 1024);
  cudaMalloc(&clk, 4);
  const int iters = 1000;
  chain<OP><<<1, threads>>>(out, clk, 1.0, 1e-3, 2.f, iters);
  chain<OP><<<1, threads>>>(out, clk, 1.0, 1e-3, 2.f, iters);
  cudaMemcpy(&h, clk, 4, cudaMemcpyDeviceToHost);
  printf("%-44s %4d threads: %7.1f clk per op-group\n", name, threads, (double)h / (iters * 8));
}

int main() {
  for (int threads : {32, 128, 512, 1024}) {
    run<0>("DADD dependent", threads);
    run<1>("F2F.F64.F32 + DADD + F2F.F32.F64", threads);
    run<2>("CREDUX.MIN + add", threads);
    run<3>("DADD + DSETP + select", threads);
    run<4>("SHFL + add", threads);
    run<5>("VOTE + POPC + add", threads);
  }
  return 0;
}
