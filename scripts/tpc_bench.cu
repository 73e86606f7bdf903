// Does a MUFU-heavy CTA run faster when its TPC sibling SM is idle?  (B200: 74 TPCs x 2 SMs.)
// Launches the packed scan element (as scripts/mufu_bench.cu mode 5) with 37/74/111/148/296/444 CTAs of 128 threads and
// prints the time and the number of distinct SMs / TPCs (smid >> 1) used.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/tpc_bench scripts/tpc_bench.cu
#include <cstdio>
#include <set>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up(unsigned long long r) { float2 d; asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r)); return d; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { unsigned long long rd; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)), "l"(pk(c.x, c.y))); return up(rd); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { unsigned long long rd; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y))); return up(rd); }

template <int MODE>
__global__ void k(float *out, int *smid, const float *in, int iters) {
  constexpr int ILP = 16;
  float a2[ILP], h[ILP];
  float y = 0.f, y1 = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a2[i] = -0.01f * (i + 1); h[i] = 0.f; }
  extern __shared__ __align__(16) float sBC[];
  for (int i = threadIdx.x; i < 64 * 40; i += blockDim.x) sBC[i] = 0.5f + 1e-3f * i;
  __syncthreads();
  const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (threadIdx.x == 0) { int s; asm("mov.u32 %0, %%smid;" : "=r"(s)); smid[bid] = s; }
  if (threadIdx.x >= 128) return;   // a 5th warp, as the scan kernel's producer
  float dl = in[threadIdx.x & 31];
  for (int it = 0; it < iters; ++it) {
    dl = dl * 1.0001f + 1e-6f;
    const float dlu = dl;
    const float *row = sBC + (it & 63) * 40;
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) h[i] = ex2(h[i] * 0.5f + dl);
    } else if (MODE == 2) {   // FFMA2 only
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < ILP; i += 2) {
        const float2 hh = ffma2(make_float2(dl, dl), make_float2(h[i], h[i + 1]), make_float2(a2[i], a2[i + 1]));
        h[i] = hh.x; h[i + 1] = hh.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < ILP; i += 4) {
        const float4 bv = *reinterpret_cast<const float4 *>(row + i), cv = *reinterpret_cast<const float4 *>(row + 16 + i);
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const int s = i + 2 * hp;
          const float2 arg = fmul2(make_float2(dl, dl), make_float2(a2[s], a2[s + 1]));
          const float2 a = make_float2(ex2(arg.x), ex2(arg.y));
          const float2 b = fmul2(make_float2(dlu, dlu), hp == 0 ? make_float2(bv.x, bv.y) : make_float2(bv.z, bv.w));
          const float2 hh = ffma2(a, make_float2(h[s], h[s + 1]), b);
          h[s] = hh.x; h[s + 1] = hh.y;
          const float2 yy = ffma2(hh, hp == 0 ? make_float2(cv.x, cv.y) : make_float2(cv.z, cv.w), make_float2(y, y1));
          y = yy.x; y1 = yy.y;
        }
      }
    }
  }
  float s = y + y1;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += h[i];
  out[bid * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int blocks, int threads, int iters, size_t smem = 64 * 40 * 4, dim3 grid3 = dim3(0, 0, 0)) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  dim3 grid = grid3.x ? grid3 : dim3(blocks);
  float *out, *in; int *smid;
  cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&in, 128); cudaMemset(in, 0, 128); cudaMalloc(&smid, 4 * blocks);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, threads, smem>>>(out, smid, in, iters);
  cudaEventRecord(e0);
  k<MODE><<<grid, threads, smem>>>(out, smid, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<int> h(blocks); cudaMemcpy(h.data(), smid, 4 * blocks, cudaMemcpyDeviceToHost);
  std::set<int> sms(h.begin(), h.end()), tpcs; for (int s : h) tpcs.insert(s >> 1);
  int cnt[256] = {0}, mx = 0; for (int s : h) { cnt[s & 255]++; } for (int i = 0; i < 256; ++i) mx = cnt[i] > mx ? cnt[i] : mx;
  printf("%-34s blocks=%4d x %4d thr: %8.3f ms  SMs=%3zu TPCs=%3zu max CTAs/SM=%d  per-warp %.1f cyc/iter\n", name, blocks, threads, ms, sms.size(),
         tpcs.size(), mx, ms * 1e-3 * 1.965e9 / iters);
  cudaFree(out); cudaFree(in); cudaFree(smid);
}

int main() {
  for (int thr : {128, 256, 512})
    for (int b : {37, 74, 111, 148, 296, 444}) {
      run<1>("scan element packed+LDS", b, thr, 8192);
    }
  for (int z : {6, 12, 24, 36, 48}) run<1>("scan-like 160thr 45KB grid(3,4,z)", 12 * z, 160, 8192, 45184, dim3(3, 4, z));
  for (int z : {6, 12, 24, 36}) run<1>("scan-like 160thr 45KB 1-D grid", 12 * z, 160, 8192, 45184);
  for (int b : {37, 74, 148, 296}) run<0>("ex2 only", b, 128, 8192);
  for (int b : {37, 74, 148, 296}) run<2>("FFMA2 only", b, 128, 8192);
  return 0;
}
