// Shared device helpers and host-side error plumbing for libsigma_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sigma_b200.h"

namespace sigma {

constexpr float kLog2e = 1.4426950408889634f;

// ---- host-side error state (thread-local string, no exceptions across the ABI) ----
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define SIGMA_CHECK_ARG(cond, ...)         \
  do {                                     \
    if (!(cond)) {                         \
      ::sigma::set_error(__VA_ARGS__);     \
      return SIGMA_EINVAL;                 \
    }                                      \
  } while (0)

#define SIGMA_CHECK_CUDA(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::sigma::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
      return SIGMA_ECUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define SIGMA_CHECK_LAUNCH()                      \
  do {                                            \
    ::sigma::count_launch();                      \
    SIGMA_CHECK_CUDA(cudaPeekAtLastError());      \
  } while (0)

// ---- device math ----
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// F.softplus with the default threshold 20 (selective_scan_fwd_kernel.cuh:133,
// selective_scan_interface.py:107).  log1pf keeps small deltas accurate (delta ~ 1e-3 is the
// common case after dt_init); the exp argument is <= 20 so the fast exp is safe.
__device__ __forceinline__ float softplus20(float x) {
  return x <= 20.f ? log1pf(ex2(x * kLog2e)) : x;
}

__device__ __forceinline__ float silu(float x) { return x / (1.f + ex2(-x * kLog2e)); }

// ---- cp.async (LDGSTS) ----
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ float f4_get(const float4 &v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

}  // namespace sigma
