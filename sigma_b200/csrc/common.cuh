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

// ---- shared host-side parameter blocks ----
struct RowNormParams {
  const float *y;          // K slabs
  long long k_stride;      // floats between slabs
  int K;
  const float *gamma, *beta;
  const float *z; long long z_row_stride;       // nullable
  const float *gate;                            // nullable, (rows / rows_per_batch, D)
  float *out;
  long long rows, rows_per_batch;
  long long in_batch_stride, out_batch_stride, out_row_stride;
  int D;
  float eps;
  // row addressing mode (fast kernel only): 0 = plain rows;
  // 1 = PatchMerging2D gather (vmamba.py:619-636): y is (batch, gH, gW, D/4), row (b,i,j) = the four pixels
  //     (2i,2j), (2i+1,2j), (2i,2j+1), (2i+1,2j+1) concatenated, zeros beyond odd gH / gW;
  // 2 = PatchExpand pixel shuffle (MambaDecoder.py:24-28): input rows are (b, h, w, p1, p2) sub-rows of D channels,
  //     row lands at out (b, 2h+p1, 2w+p2)
  int mode = 0, gH = 0, gW = 0;
};

// ---- device math ----
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// F.softplus with the default threshold 20 (selective_scan_fwd_kernel.cuh:133,
// selective_scan_interface.py:107), branch-free and with ONE MUFU op so it schedules inside the scan's inner loop
// (the SFU is the scan's busiest unit):
//   softplus(x) = max(x,0) + log1p(z),  z = exp(-|x|) in (0,1],  log1p(z) = z·q(z)
// q = degree-8 Chebyshev interpolant of log1p(z)/z on [0,1]: relative error of log1p < 2.5e-7 over the whole range
// evaluated in fp32 (the earlier series / MUFU.LG2 split needed a second MUFU op and reached 1e-5 for z > 2^-6).
// For x > 20, z < 2.1e-9 and the sum rounds to x, which is exactly the reference's thresholded branch.
__device__ __forceinline__ float softplus20(float x) {
  const float z = ex2(-fabsf(x) * kLog2e);
  float q = 0.005126102361828089f;
  q = fmaf(q, z, -0.029074065387248993f);
  q = fmaf(q, z, 0.0775160863995552f);
  q = fmaf(q, z, -0.13602247834205627f);
  q = fmaf(q, z, 0.19076880812644958f);
  q = fmaf(q, z, -0.2483539879322052f);
  q = fmaf(q, z, 0.3331812024116516f);
  q = fmaf(q, z, -0.4999944567680359f);
  q = fmaf(q, z, 0.9999999403953552f);
  return fmaf(q, z, fmaxf(x, 0.f));
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2: two fp32 lanes per instruction, same FLOP rate as the
// scalar pipe but half the issue slots — scripts/mufu_bench.cu).  Operands are adjacent register pairs.
struct f2 { float x, y; };
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f2 unpack2(unsigned long long r) {
  f2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
  return d;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)), "l"(pack2(c.x, c.y)));
  return unpack2(rd);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  unsigned long long rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pack2(a.x, a.y)), "l"(pack2(b.x, b.y)));
  return unpack2(rd);
}

// softplus20 of two values at once: the degree-8 polynomial runs as 9 FFMA2 instead of 18 FFMA (the scan kernels are
// issue-limited; the coefficient pairs are loop-invariant register pairs).  Same arithmetic per element as softplus20.
__device__ __forceinline__ f2 softplus20x2(float x0, float x1) {
  const f2 z = f2{ex2(-fabsf(x0) * kLog2e), ex2(-fabsf(x1) * kLog2e)};
  f2 q = fma2(f2{0.005126102361828089f, 0.005126102361828089f}, z, f2{-0.029074065387248993f, -0.029074065387248993f});
  q = fma2(q, z, f2{0.0775160863995552f, 0.0775160863995552f});
  q = fma2(q, z, f2{-0.13602247834205627f, -0.13602247834205627f});
  q = fma2(q, z, f2{0.19076880812644958f, 0.19076880812644958f});
  q = fma2(q, z, f2{-0.2483539879322052f, -0.2483539879322052f});
  q = fma2(q, z, f2{0.3331812024116516f, 0.3331812024116516f});
  q = fma2(q, z, f2{-0.4999944567680359f, -0.4999944567680359f});
  q = fma2(q, z, f2{0.9999999403953552f, 0.9999999403953552f});
  return fma2(q, z, f2{fmaxf(x0, 0.f), fmaxf(x1, 0.f)});
}

__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.f + ex2(-x * kLog2e)); }

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
