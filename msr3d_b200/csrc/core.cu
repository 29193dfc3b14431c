// Error reporting, ABI version and launch accounting for libmsr3d_b200.so.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace msr3d {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
int g_pdl = [] {
  const char *e = getenv("MSR3D_PDL");
  return (e != nullptr && e[0] == '0') ? 0 : 1;
}();

int g_tune[TUNE_COUNT] = {0};
static int g_tune_env = [] {   // MSR3D_TUNE="id=value,id=value" (A/B experiments from the shell)
  const char *e = getenv("MSR3D_TUNE");
  while (e != nullptr && *e) {
    int id = 0, v = 0, n = 0;
    if (sscanf(e, "%d=%d%n", &id, &v, &n) == 2 && id >= 0 && id < TUNE_COUNT) g_tune[id] = v;
    else break;
    e += n;
    if (*e == ',') ++e;
  }
  return 0;
}();

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace msr3d

MSR3D_API const char *msr3d_last_error(void) { return msr3d::g_err; }
MSR3D_API int msr3d_abi_version(void) { return 4; }
MSR3D_API unsigned long long msr3d_launch_count(void) {
  return msr3d::g_launches.load(std::memory_order_relaxed);
}
/* Enable (1, default) / disable (0) programmatic dependent launch for subsequent launches. */
MSR3D_API void msr3d_set_pdl(int enabled) { msr3d::g_pdl = enabled ? 1 : 0; }
MSR3D_API int msr3d_get_pdl(void) { return msr3d::g_pdl; }
/* Tuning hook (tools/gemm_sweep.py): id = TUNE_* in csrc/common.cuh, value 0 = library heuristic. */
MSR3D_API int msr3d_set_tunable(int id, int value) {
  if (id < 0 || id >= msr3d::TUNE_COUNT) return msr3d::set_error(MSR3D_ERR_INVALID_ARG, "set_tunable: bad id %d", id);
  msr3d::g_tune[id] = value;
  return MSR3D_OK;
}
