// gwd_core.cu -- library-wide state: last-error string, launch counter, device queries.
#include <stdarg.h>
#include <string.h>
#include "gwd_common.cuh"

std::atomic<int64_t> g_gwd_launches{0};
static thread_local char g_err[512] = "";

void gwd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int gwd_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

extern "C" const char* gwd_last_error(void) { return g_err; }
extern "C" int gwd_version(void) { return 100; }
extern "C" int64_t gwd_launch_count(void) { return g_gwd_launches.load(); }
extern "C" void gwd_reset_launch_count(void) { g_gwd_launches.store(0); }
