// C-ABI plumbing shared by all entry points: version, per-thread last-error string, plane
// geometry helper.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace plb {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Function attributes and the SM count belong to a DEVICE: cached per (kernel, ordinal) so that a
// process using several GPUs configures each of them (a single `static bool` would only set the
// > 48 KB dynamic shared-memory limit on the first device used).
int ensure_dynamic_smem(const void *func, int bytes, const char *name) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, int> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return (int)e;
  }
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(func, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return PLB_OK;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%s, %d bytes of dynamic shared memory): %s", name, bytes, cudaGetErrorString(e));
    return (int)e;
  }
  done[key] = bytes;
  return PLB_OK;
}

int device_sm_count() {
  static std::mutex mu;
  static std::map<int, int> sms;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lock(mu);
  auto it = sms.find(dev);
  if (it != sms.end()) return it->second;
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  sms[dev] = n;
  return n;
}
}  // namespace plb

extern "C" int plb_version(void) { return 100; }  // 0.1.0

extern "C" const char *plb_last_error_string(void) { return plb::g_err; }

extern "C" int64_t plb_plane_bytes(int64_t rows, int64_t K, int32_t *row_groups, int32_t *k_blocks) {
  if (rows <= 0 || K <= 0) {
    plb::set_error("plb_plane_bytes: rows and K must be positive");
    return PLB_EINVAL;
  }
  int64_t g = plb::ceil_div(rows, 8);
  int64_t kb = plb::ceil_div(K, plb::kPackK);
  if (row_groups) *row_groups = (int32_t)g;
  if (k_blocks) *k_blocks = (int32_t)kb;
  return g * kb * plb::kPanelFloats * 4;
}
