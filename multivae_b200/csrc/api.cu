// Error plumbing + version probe of the C-ABI.
#include <cstring>
#include <cstdlib>

#include "common.cuh"

namespace mv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
}  // namespace mv

extern "C" unsigned long long mv_launch_count(void) { return __atomic_load_n(&mv::g_launches, __ATOMIC_RELAXED); }

extern "C" const char* mv_last_error(void) { return mv::g_err; }

extern "C" int mv_version(int* major, int* minor, int* sm_arch) {
  if (major) *major = 0;
  if (minor) *minor = 1;
  if (sm_arch) *sm_arch = 100;
  return MV_OK;
}
