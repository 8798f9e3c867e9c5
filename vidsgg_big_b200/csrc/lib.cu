// Library-level entry points: error reporting, version, device query.
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace vsg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

long long launches();
}  // namespace vsg

extern "C" long long vsg_launch_count(void) { return vsg::launches(); }
extern "C" const char* vsg_last_error(void) { return vsg::g_err; }
extern "C" int vsg_version(void) { return 1; }
extern "C" int vsg_built_for_sm(void) { return 100; }
extern "C" int vsg_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { vsg::set_error("vsg_device_sm_count: no CUDA device"); return VSG_E_LAUNCH; }
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { vsg::set_error("vsg_device_sm_count: attribute query failed"); return VSG_E_LAUNCH; }
  return n;
}
