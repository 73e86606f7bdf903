// Micro-benchmarks on B200 (sm_100a): MUFU.EX2 throughput, FFMA vs packed FFMA2, and the scan element
// (1 ex2 + 4 fp32 ops) in scalar and packed form.  Elements per clk per SM at 1.965 GHz x 148 SMs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mufu_bench scripts/mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d; asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd)); return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d; asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd)); return d;
}

// MODE 0: ex2 chain; 1: scalar scan element; 2: packed scan element; 3: FFMA only; 4: FFMA2 only;
// 5: packed scan element with B/C fetched by broadcast LDS.128 from shared memory (as the real kernel does);
// 6: as 5 but B/C fetched with one LDS.32 per lane + no broadcast (reference point for LSU cost)
template <int MODE, int ILP>
__global__ void k(float *out, const float *in, int iters) {
  float a2[ILP], h[ILP], Bc[ILP], Cc[ILP];
  float y = 0.f, y1 = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a2[i] = -0.01f * (i + 1); h[i] = 0.f; Bc[i] = 0.5f + 0.01f * i; Cc[i] = 0.25f; }
  __shared__ __align__(16) float sBC[64 * 40];
  for (int i = threadIdx.x; i < 64 * 40; i += blockDim.x) sBC[i] = 0.5f + 1e-3f * i;
  __syncthreads();
  float dl = in[threadIdx.x & 31], u = 1.0f;
  for (int it = 0; it < iters; ++it) {
    dl = dl * 1.0001f + 1e-6f;   // loop-varying so nothing hoists
    const float dlu = dl * u;
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) h[i] = ex2(h[i] * 0.5f + dl);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) { const float a = ex2(dl * a2[i]); h[i] = fmaf(a, h[i], dlu * Bc[i]); y = fmaf(h[i], Cc[i], y); }
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < ILP; i += 2) {
        const float2 arg = fmul2(make_float2(dl, dl), make_float2(a2[i], a2[i + 1]));
        const float2 a = make_float2(ex2(arg.x), ex2(arg.y));
        const float2 b = fmul2(make_float2(dlu, dlu), make_float2(Bc[i], Bc[i + 1]));
        const float2 hh = ffma2(a, make_float2(h[i], h[i + 1]), b);
        h[i] = hh.x; h[i + 1] = hh.y;
        const float2 yy = ffma2(hh, make_float2(Cc[i], Cc[i + 1]), make_float2(y, y1));
        y = yy.x; y1 = yy.y;
      }
    } else if (MODE == 5 || MODE == 6) {
      const float *row = sBC + (it & 63) * 40;
#pragma unroll
      for (int i = 0; i < ILP; i += 4) {
        float4 bv, cv;
        if (MODE == 5) { bv = *reinterpret_cast<const float4 *>(row + i); cv = *reinterpret_cast<const float4 *>(row + 16 + i); }
        else { const float t0 = row[(threadIdx.x + i) & 31]; bv = make_float4(t0, t0, t0, t0); cv = bv; }
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
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) h[i] = fmaf(dl, h[i], Bc[i]);
    } else {
#pragma unroll
      for (int i = 0; i < ILP; i += 2) {
        const float2 hh = ffma2(make_float2(dl, dl), make_float2(h[i], h[i + 1]), make_float2(Bc[i], Bc[i + 1]));
        h[i] = hh.x; h[i + 1] = hh.y;
      }
    }
  }
  float s = y + y1;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int ILP>
void run(const char *name, int wps, int iters) {
  int threads = 128 * wps > 1024 ? 1024 : 128 * wps, blocks = 148 * (128 * wps / threads);
  float *out, *in; cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&in, 128); cudaMemset(in, 0, 128);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, ILP><<<blocks, threads>>>(out, in, iters);
  cudaEventRecord(e0);
  k<MODE, ILP><<<blocks, threads>>>(out, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double el = (double)blocks * threads * iters * ILP;
  printf("%-34s warps/SMSP=%d ILP=%2d: %8.3f ms %9.1f Gelem/s = %6.2f elem/clk/SM\n", name, wps, ILP, ms, el / ms / 1e6, el / ms / 1e6 / (148 * 1.965));
  cudaFree(out); cudaFree(in);
}

int main() {
  for (int wps : {1, 2, 4, 8}) {
    run<0, 16>("ex2 only", wps, 4096);
    run<3, 16>("FFMA only", wps, 4096);
    run<4, 16>("FFMA2 only (2 elem/instr)", wps, 4096);
    run<1, 16>("scan element scalar (ex2+4 fp32)", wps, 4096);
    run<2, 16>("scan element packed (ex2+2 f32x2)", wps, 4096);
    run<5, 16>("packed + B/C by broadcast LDS.128", wps, 4096);
    run<6, 16>("packed + B/C by 1 LDS.32 per 4 st.", wps, 4096);
  }
  return 0;
}
