// Row-wise / stencil kernels of the channels-last pipeline (fp32):
//   row_norm_kernel  : [sum over K direction outputs] -> LayerNorm -> [· SiLU(z)] -> [· gate]   (one warp per row)
//                      = nn.LayerNorm (vmamba.py:1693,2173,...) when K=1, z=gate=NULL, and
//                      = CrossMerge sum + out_norm + y·SiLU(z) (vmamba.py:217-224,1077) otherwise
//   dwconv3x3_silu   : depthwise 3x3 (pad 1) + bias + SiLU on NHWC (vmamba.py:683-692,1072)
// All are HBM-bound; loads/stores are 16-byte, rows are contiguous in the channel dimension.
#include "common.cuh"

namespace sigma {


__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int MAXV>
__global__ void __launch_bounds__(256) row_norm_kernel(const RowNormParams p) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const long long bi = row / p.rows_per_batch, ri = row - bi * p.rows_per_batch;
  const float *in = p.y + bi * p.in_batch_stride + ri * p.D;
  const int nvec = p.D >> 2;
  float4 x[MAXV];
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < MAXV; ++v) {
    const int idx = lane + 32 * v;
    x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < nvec) {
      float4 a = __ldg(reinterpret_cast<const float4 *>(in) + idx);
      for (int k = 1; k < p.K; ++k) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(in + k * p.k_stride) + idx);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      x[v] = a;
      s += (a.x + a.y) + (a.z + a.w);
    }
  }
  const float mean = warp_sum(s) / (float)p.D;
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < MAXV; ++v) {
    if (lane + 32 * v < nvec) {
      const float dx = x[v].x - mean, dy = x[v].y - mean, dz = x[v].z - mean, dw = x[v].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.D + p.eps);
  float *out = p.out + bi * p.out_batch_stride + ri * p.out_row_stride;
  const float *zr = p.z ? p.z + row * p.z_row_stride : nullptr;
  const float *gr = p.gate ? p.gate + bi * p.D : nullptr;
#pragma unroll
  for (int v = 0; v < MAXV; ++v) {
    const int idx = lane + 32 * v;
    if (idx < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4 *>(p.gamma) + idx);
      const float4 b = __ldg(reinterpret_cast<const float4 *>(p.beta) + idx);
      float4 o;
      o.x = fmaf((x[v].x - mean) * rstd, g.x, b.x);
      o.y = fmaf((x[v].y - mean) * rstd, g.y, b.y);
      o.z = fmaf((x[v].z - mean) * rstd, g.z, b.z);
      o.w = fmaf((x[v].w - mean) * rstd, g.w, b.w);
      if (zr) {
        const float4 zz = __ldg(reinterpret_cast<const float4 *>(zr) + idx);
        o.x *= silu(zz.x); o.y *= silu(zz.y); o.z *= silu(zz.z); o.w *= silu(zz.w);
      }
      if (gr) {
        const float4 gg = __ldg(reinterpret_cast<const float4 *>(gr) + idx);
        o.x *= gg.x; o.y *= gg.y; o.z *= gg.z; o.w *= gg.w;
      }
      reinterpret_cast<float4 *>(out)[idx] = o;
    }
  }
}

int row_norm_launch(const RowNormParams &p, cudaStream_t stream) {
  if (p.rows == 0) return SIGMA_OK;
  const int nvec = p.D >> 2;
  const int warps = 8;
  const unsigned grid = (unsigned)((p.rows + warps - 1) / warps);
#define LAUNCH(MV) row_norm_kernel<MV><<<grid, warps * 32, 0, stream>>>(p)
  if (nvec <= 32) LAUNCH(1);
  else if (nvec <= 64) LAUNCH(2);
  else if (nvec <= 128) LAUNCH(4);
  else if (nvec <= 256) LAUNCH(8);
  else if (nvec <= 512) LAUNCH(16);
  else if (nvec <= 1024) LAUNCH(32);
  else { set_error("row_norm: D=%d > 4096 unsupported", p.D); return SIGMA_EUNSUPPORTED; }
#undef LAUNCH
  SIGMA_CHECK_LAUNCH();
  return SIGMA_OK;
}

// ---- depthwise 3x3 + bias + SiLU, NHWC ----
// CTA: 64 channels (16 float4 lanes) x 16 position-threads, each walking a strip of positions; the 9 taps of
// the thread's 4 channels live in registers (loaded once through shared memory from the (D,1,3,3) weight).
constexpr int DW_CH = 64, DW_PT = 16, DW_POS_PER_CTA = 256;

__global__ void __launch_bounds__(256) dwconv3x3_silu_kernel(const float *__restrict__ x, long long x_row_stride,
                                                            long long x_batch_stride, const float *__restrict__ w,
                                                            const float *__restrict__ bias, float *__restrict__ y,
                                                            long long y_batch_stride, int batch, int H, int W, int D) {
  __shared__ __align__(16) float sw[9][DW_CH];
  __shared__ __align__(16) float sb[DW_CH];
  const int c0 = blockIdx.x * DW_CH;
  for (int i = threadIdx.x; i < 9 * DW_CH; i += blockDim.x) {
    const int c = i / 9, tap = i - c * 9;
    sw[tap][c] = (c0 + c < D) ? w[(long long)(c0 + c) * 9 + tap] : 0.f;
  }
  for (int i = threadIdx.x; i < DW_CH; i += blockDim.x) sb[i] = (bias && c0 + i < D) ? bias[c0 + i] : 0.f;
  __syncthreads();
  const int cq = threadIdx.x & 15, pr = threadIdx.x >> 4;
  const int c = c0 + 4 * cq;
  if (c >= D) return;
  float4 wt[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) wt[tap] = *reinterpret_cast<const float4 *>(&sw[tap][4 * cq]);
  const float4 bv = *reinterpret_cast<const float4 *>(&sb[4 * cq]);
  const long long HW = (long long)H * W, total = HW * batch;
  const long long p0 = (long long)blockIdx.y * DW_POS_PER_CTA;
  for (long long p = p0 + pr; p < min(total, p0 + DW_POS_PER_CTA); p += DW_PT) {
    const int b = (int)(p / HW);
    const int hw = (int)(p - (long long)b * HW);
    const int h = hw / W, wq = hw - h * W;
    const float *xb = x + (long long)b * x_batch_stride + c;
    float4 acc = bv;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int hh = h + dy;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ww = wq + dx;
        if (ww < 0 || ww >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(xb + ((long long)hh * W + ww) * x_row_stride));
        const float4 k = wt[(dy + 1) * 3 + (dx + 1)];
        acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y);
        acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
      }
    }
    float4 o;
    o.x = silu(acc.x); o.y = silu(acc.y); o.z = silu(acc.z); o.w = silu(acc.w);
    *reinterpret_cast<float4 *>(y + (long long)b * y_batch_stride + (long long)hw * D + c) = o;
  }
}

int dwconv3x3_silu_launch(const float *x, long long x_row_stride, long long x_batch_stride, const float *w,
                          const float *bias, float *y, long long y_batch_stride, int batch, int H, int W, int D,
                          cudaStream_t stream) {
  const long long total = (long long)batch * H * W;
  if (total == 0) return SIGMA_OK;
  dim3 grid((D + DW_CH - 1) / DW_CH, (unsigned)((total + DW_POS_PER_CTA - 1) / DW_POS_PER_CTA));
  dwconv3x3_silu_kernel<<<grid, 256, 0, stream>>>(x, x_row_stride, x_batch_stride, w, bias, y, y_batch_stride, batch, H, W, D);
  SIGMA_CHECK_LAUNCH();
  return SIGMA_OK;
}

}  // namespace sigma
