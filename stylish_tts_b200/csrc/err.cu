// Error reporting + library-level queries of libstylish_b200.
#include <stdarg.h>

#include "common.cuh"

namespace sty {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace sty

extern "C" const char* sty_last_error(void) { return sty::g_err; }
extern "C" int sty_version(void) { return 100; }
extern "C" int sty_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}
