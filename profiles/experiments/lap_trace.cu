// Per-phase clock sums of one Dijkstra step of lap_kernel_v3 (experiments only; the product library is built
// without PLB_LAP_TRACE).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPLB_LAP_TRACE \
//        profiles/experiments/lap_trace.cu pleas_merging_b200/csrc/api.cu -o /tmp/lap_trace && /tmp/lap_trace 2048
#include "../../pleas_merging_b200/csrc/lap.cu"

#include <math.h>
#include <vector>

static double lcg_uniform(uint64_t &s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return ((s >> 11) + 0.5) / 9007199254740992.0;
}
static double randn(uint64_t &s) {
  const double a = lcg_uniform(s), b = lcg_uniform(s);
  return sqrt(-2.0 * log(a)) * cos(6.283185307179586 * b);
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2048;
  const char *names[10] = {"cost row arrives (LDS u + LDG)", "relaxation + thread best", "warp arg-min (3 redux)",
                           "publish + barrier", "winner slots read", "cross-warp arg-min", "shfl + list compaction",
                           "per-search overhead (dual update, flip, reset)", "steps", "-"};
  for (int kind = 0; kind < 2; ++kind) {
    uint64_t seed = 12345;
    std::vector<float> h((size_t)n * n);
    std::vector<double> ro(n), co(n);
    for (int i = 0; i < n; ++i) ro[i] = 30 * randn(seed), co[i] = 30 * randn(seed);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) h[(size_t)i * n + j] = (float)(randn(seed) + (kind ? ro[i] + co[j] : 0.0));
    float *d;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const float **dc;
    int32_t *dn, *dl, *ds;
    int64_t *dout, **douts;
    double *dobj;
    cudaMalloc(&dc, 8), cudaMalloc(&dn, 4), cudaMalloc(&dl, 4), cudaMalloc(&ds, 4), cudaMalloc(&dout, 8 * n);
    cudaMalloc(&douts, 8), cudaMalloc(&dobj, 8);
    cudaMemcpy(dc, &d, 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dn, &n, 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dl, &n, 4, cudaMemcpyHostToDevice);
    cudaMemcpy(douts, &dout, 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      int rc = plb_lap_solve_batched(dc, dn, dl, douts, dobj, ds, 1, n, 1, nullptr);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      if (rc) printf("rc %d %s\n", rc, plb_last_error_string());
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    uint32_t tr[2][10];
    cudaMemcpyFromSymbol(tr, plb::plb_lap_trace, sizeof(tr));
    printf("%s n=%d: %.2f ms, %u steps, %.0f ns/step\n", kind ? "offsets" : "randn", n, ms, tr[0][8],
           1e6 * ms / tr[0][8]);
    for (int w = 0; w < 2; ++w) {
      printf("  %s thread:\n", w ? "last" : "first");
      for (int k = 0; k < 8; ++k)
        printf("    %-50s %8.1f clk/step\n", names[k], (double)tr[w][k] / tr[w][8]);
    }
  }
  return 0;
}
