// Micro-benchmark: MUFU.EX2 throughput on B200, alone and mixed with the FP32 work of the scan inner loop.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mufu_bench scripts/mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE, int ILP>
__global__ void k(float *out, int iters, float seed) {
  float a[ILP], h[ILP], y = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = seed * (threadIdx.x + i + 1) * 1e-3f; h[i] = 0.f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) { a[i] = ex2(a[i]) - 1.0f; }                       // ex2 + 1 fadd, ILP independent chains
      if (MODE == 1) { float e = ex2(a[i] * -0.37f); h[i] = fmaf(e, h[i], a[i] * 0.5f); y = fmaf(h[i], 0.25f, y); }  // scan element
      if (MODE == 2) { h[i] = fmaf(a[i], h[i], 0.5f); y = fmaf(h[i], 0.25f, y); }   // no MUFU
    }
  }
  float s = y;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
void run(const char *name, int blocks, int threads, int iters) {
  float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, ILP><<<blocks, threads>>>(out, iters, 1.f);
  cudaEventRecord(e0);
  k<MODE, ILP><<<blocks, threads>>>(out, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * ILP;
  printf("%-28s blocks=%4d thr=%4d ILP=%2d: %8.3f ms  %8.1f Gop/s  (%.2f per clk per SM @1.965GHz x148)\n", name, blocks, threads, ILP, ms,
         ops / ms / 1e6, ops / ms / 1e6 / (148 * 1.965));
  cudaFree(out);
}

int main() {
  for (int wps : {1, 2, 4, 8}) {   // warps per SM sub-partition
    int threads = 128 * wps > 1024 ? 1024 : 128 * wps, blocks = 148 * (128 * wps / threads);
    run<0, 16>("ex2 only", blocks, threads, 4096);
    run<1, 16>("scan element (ex2+3fp32)", blocks, threads, 4096);
    run<1, 4>("scan element (ex2+3fp32)", blocks, threads, 4096);
    run<2, 16>("fma only", blocks, threads, 4096);
  }
  return 0;
}
