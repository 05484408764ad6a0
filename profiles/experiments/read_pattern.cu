// Read-bandwidth of the activation access patterns on B200: x[outer][C][inner] fp32, C = 64 rows,
// each CTA streams a K range.  Pattern A: a warp instruction reads 8 rows x 64 B (the pack / direct
// kernels' panel mapping).  Pattern B: a warp instruction reads 2 rows x 256 B.  Pattern C: 1 row x 512 B.
#include <cstdio>
#include <cuda_runtime.h>
template <int ROWS_PER_INSTR>
__global__ void __launch_bounds__(256) reader(const float4 *__restrict__ x, int C, int inner4, int outer, int kchunks_per_cta,
                                              float *out) {
  // chunk = 16 floats (4 float4) of k for pattern A; generalised: a warp instr covers ROWS_PER_INSTR rows x (32/ROWS_PER_INSTR) float4
  constexpr int F4 = 32 / ROWS_PER_INSTR;       // float4 per row per instruction
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane / F4, c4 = lane % F4;
  const long total_k4 = (long)outer * inner4;   // float4 along k per row (over all images)
  const long k4_begin = (long)blockIdx.x * kchunks_per_cta * 4;
  float acc = 0.f;
  // warps of a CTA take row groups round-robin; each warp walks the CTA's k range
  for (int rg = warp; rg < C / ROWS_PER_INSTR; rg += 8) {
    const int row = rg * ROWS_PER_INSTR + r;
    for (long k4 = k4_begin + c4; k4 < k4_begin + (long)kchunks_per_cta * 4 && k4 < total_k4; k4 += F4) {
      const long o = k4 / inner4, i = k4 - o * inner4;
      const float4 v = __ldg(x + (o * C + row) * inner4 + i);
      acc += v.x + v.y + v.z + v.w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
int main() {
  const int C = 64, inner = 112 * 112, outer = 32;
  const long n = (long)C * inner * outer;
  float4 *x; float *out;
  cudaMalloc(&x, n * 4); cudaMalloc(&out, 4); cudaMemset(x, 0, n * 4);
  const int ctas = 148 * 4;
  const int kchunks = (int)(((long)outer * inner / 16 + ctas - 1) / ctas);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pat = 0; pat < 3; ++pat) {
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
      cudaEventRecord(e0);
      if (pat == 0) reader<8><<<ctas, 256>>>(x, C, inner / 4, outer, kchunks, out);
      if (pat == 1) reader<2><<<ctas, 256>>>(x, C, inner / 4, outer, kchunks, out);
      if (pat == 2) reader<1><<<ctas, 256>>>(x, C, inner / 4, outer, kchunks, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("pattern %c (%d rows x %d B per warp instruction): %.3f ms = %.2f TB/s\n", 'A' + pat, pat == 0 ? 8 : (pat == 1 ? 2 : 1),
           pat == 0 ? 64 : (pat == 1 ? 256 : 512), best, n * 4 / best / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
