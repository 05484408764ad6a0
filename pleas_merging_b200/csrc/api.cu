// C-ABI plumbing shared by all entry points: version, per-thread last-error string, plane
// geometry helper.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace plb {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
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
