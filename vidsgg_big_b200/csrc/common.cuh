// Shared helpers for libvsgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/vsg_b200.h"

namespace vsg {

void set_error(const char* fmt, ...);

void count_launches(int n);

inline int check_launch(const char* what, int n_kernels = 1) {
  count_launches(n_kernels);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VSG_E_LAUNCH;
  }
  return VSG_OK;
}

#define VSG_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::vsg::set_error(__VA_ARGS__);    \
      return VSG_E_INVALID;             \
    }                                   \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();

// cudaFuncSetAttribute is per device (context), so a process-wide `static bool` would leave a second GPU of the same process without
// the raised shared-memory limit: one bit per device ordinal, set after the attribute call succeeded there (thread-safe).
struct PerDeviceFlag {
  std::atomic<unsigned long long> done{0};
  bool is_set(int dev) const { return (done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull; }
  void set(int dev) { done.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};
inline int current_device() { int d = 0; cudaGetDevice(&d); return d; }

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit load that does not pollute L1 (tracks are re-read from L2, not L1)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// upper_bound-1 over a monotone int64 offsets array: largest v with off[v] <= x
__device__ __forceinline__ int find_segment(const int64_t* __restrict__ off, int n_seg, int64_t x) {
  int lo = 0, hi = n_seg;  // invariant off[lo] <= x < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

}  // namespace vsg
