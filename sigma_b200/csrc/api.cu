// C-ABI entry points (include/sigma_b200.h): argument checking, dtype staging, error strings.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <atomic>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace sigma {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// implemented in scan_op.cu (generic) / scan_op_tma.cu (TMA-staged) / scan_op_bwd*.cu
size_t scan_op_workspace_bytes(int batch, int dim, int dstate);
size_t scan_op_tma_workspace_bytes(int batch, int dim, int dstate);
size_t scan_op_bwd_workspace_bytes(int batch, int dim, int L, int N, int elem_bytes);
size_t scan_op_bwd_tma_workspace_bytes(int batch, int dim, int L, int N, int elem_bytes);
template <typename T>
bool scan_op_tma_eligible(const void *u, const void *delta, const void *B, const void *C, const void *out, int dim, int L,
                          int N, int G, const sigma_scan_strides &s);
template <typename T>
int scan_op_fwd_tma(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                    const float *bias, void *out, float *x, float *hs, int batch, int dim, int L, int N, int G, int softplus,
                    const sigma_scan_strides &s, void *ws, size_t ws_bytes, int force_split, cudaStream_t stream);
template <typename T>
int scan_op_fwd_generic(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                        const float *bias, void *out, float *x, float *hs, int batch, int dim, int L, int N, int G,
                        int softplus, const sigma_scan_strides &s, void *ws, size_t ws_bytes, int force_split,
                        cudaStream_t stream);
template <typename T>
int scan_op_bwd_tma(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                    const float *bias, const void *dout, void *du, void *ddelta, float *dA, float *dB, float *dC, float *dD,
                    float *dbias, int batch, int dim, int L, int N, int G, int softplus, void *ws, size_t ws_bytes,
                    int force_split, cudaStream_t stream);
template <typename T>
int scan_op_bwd_generic(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                        const float *bias, const void *dout, void *du, void *ddelta, float *dA, float *dB, float *dC,
                        float *dD, float *dbias, int batch, int dim, int L, int N, int G, int softplus, void *ws,
                        size_t ws_bytes, cudaStream_t stream);

// SIGMA_OP_GENERIC=1 forces the generic kernels (A/B timing, tests of the fallback on TMA-eligible shapes)
static bool force_generic() {
  const char *e = getenv("SIGMA_OP_GENERIC");
  return e && atoi(e) > 0;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ---- 16-bit tensors whose rows are not 16-byte aligned (L % 8 != 0, e.g. the 15 x 20 stage) cannot be TMA boxes.  When the
// fp32 image of the call IS eligible (L % 4 == 0), it is cheaper to widen the operands into scratch, run the TMA-staged fp32
// kernels and narrow the results than to take the generic kernels (measured: 0.2-0.6x of the reference kernel there). ----
template <typename T> __device__ __forceinline__ float wide(T v);
template <> __device__ __forceinline__ float wide<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float wide<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T narrow(float v);
template <> __device__ __forceinline__ __half narrow<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 narrow<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// src (n0, n1, n2, L) with element strides (s0, s1, s2, 1) -> dst contiguous fp32
template <typename T>
__global__ void widen_kernel(const T *__restrict__ src, float *__restrict__ dst, int n1, int n2, int L, long long s0, long long s1,
                             long long s2, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i % L);
    long long r = i / L;
    const int i2 = (int)(r % n2); r /= n2;
    const int i1 = (int)(r % n1); r /= n1;
    dst[i] = wide<T>(src[r * s0 + i1 * s1 + i2 * s2 + l]);
  }
}
// src contiguous fp32 (n0, n1, L) -> dst with strides (s0, s1, 1)
template <typename T>
__global__ void narrow_kernel(const float *__restrict__ src, T *__restrict__ dst, int n1, int L, long long s0, long long s1, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i % L);
    long long r = i / L;
    const int i1 = (int)(r % n1); r /= n1;
    dst[r * s0 + i1 * s1 + l] = narrow<T>(src[i]);
  }
}
template <typename T>
static int widen(const void *src, float *dst, int n0, int n1, int n2, int L, long long s0, long long s1, long long s2, cudaStream_t st) {
  const long long total = (long long)n0 * n1 * n2 * L;
  widen_kernel<T><<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>((const T *)src, dst, n1, n2, L, s0, s1, s2, total);
  SIGMA_CHECK_LAUNCH();
  return SIGMA_OK;
}
template <typename T>
static int narrow_to(const float *src, void *dst, int n0, int n1, int L, long long s0, long long s1, cudaStream_t st) {
  const long long total = (long long)n0 * n1 * L;
  narrow_kernel<T><<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(src, (T *)dst, n1, L, s0, s1, total);
  SIGMA_CHECK_LAUNCH();
  return SIGMA_OK;
}

static bool widen_shape_ok(int dim, int L, int N, int G, int elem_bytes) {
  return elem_bytes == 2 && L % 8 != 0 && L % 4 == 0 && (N == 4 || N == 8 || N == 16) && (dim / G) % 32 == 0;
}
static size_t widen_fwd_bytes(int batch, int dim, int L, int N, int G) {
  return 3 * align256((size_t)batch * dim * L * 4) + 2 * align256((size_t)batch * G * N * L * 4);
}
static size_t widen_bwd_bytes(int batch, int dim, int L, int N, int G) {
  return 5 * align256((size_t)batch * dim * L * 4) + 2 * align256((size_t)batch * G * N * L * 4);
}
static sigma_scan_strides contiguous_strides(int dim, int L, int N, int G) {
  sigma_scan_strides st;
  st.u_batch = st.delta_batch = st.out_batch = (int64_t)dim * L;
  st.u_dim = st.delta_dim = st.out_dim = L;
  st.A_dim = N; st.A_dstate = 1;
  st.B_batch = st.C_batch = (int64_t)G * N * L;
  st.B_group = st.C_group = (int64_t)N * L;
  st.B_dstate = st.C_dstate = L;
  return st;
}

template <typename T>
static int scan_fwd_dispatch(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                             const float *bias, void *out, float *x, int batch, int dim, int L, int N, int G, int softplus,
                             const sigma_scan_strides &s, void *ws, size_t ws_bytes, int force_split, cudaStream_t stream) {
  if (!force_generic() && scan_op_tma_eligible<T>(u, delta, B, C, out, dim, L, N, G, s))
    return scan_op_fwd_tma<T>(u, delta, A, B, C, D, bias, out, x, nullptr, batch, dim, L, N, G, softplus, s, ws, ws_bytes,
                              force_split, stream);
  if constexpr (sizeof(T) == 2) {
    const size_t core = align256(scan_op_tma_workspace_bytes(batch, dim, N));
    if (!force_generic() && widen_shape_ok(dim, L, N, G, 2) && ws != nullptr && ws_bytes >= core + widen_fwd_bytes(batch, dim, L, N, G) &&
        s.A_dstate >= 0) {
      char *w = (char *)ws + core;
      const size_t bdl = align256((size_t)batch * dim * L * 4), bgn = align256((size_t)batch * G * N * L * 4);
      float *u32 = (float *)w, *d32 = (float *)(w + bdl), *o32 = (float *)(w + 2 * bdl), *B32 = (float *)(w + 3 * bdl), *C32 = (float *)(w + 3 * bdl + bgn);
      int rc;
      if ((rc = widen<T>(u, u32, batch, dim, 1, L, s.u_batch, s.u_dim, 0, stream))) return rc;
      if ((rc = widen<T>(delta, d32, batch, dim, 1, L, s.delta_batch, s.delta_dim, 0, stream))) return rc;
      if ((rc = widen<T>(B, B32, batch, G, N, L, s.B_batch, s.B_group, s.B_dstate, stream))) return rc;
      if ((rc = widen<T>(C, C32, batch, G, N, L, s.C_batch, s.C_group, s.C_dstate, stream))) return rc;
      sigma_scan_strides cs = contiguous_strides(dim, L, N, G);
      cs.A_dim = s.A_dim; cs.A_dstate = s.A_dstate;
      if ((rc = scan_op_fwd_tma<float>(u32, d32, A, B32, C32, D, bias, o32, x, nullptr, batch, dim, L, N, G, softplus, cs, ws, core, force_split,
                                       stream))) return rc;
      return narrow_to<T>(o32, out, batch, dim, L, s.out_batch, s.out_dim, stream);
    }
  }
  return scan_op_fwd_generic<T>(u, delta, A, B, C, D, bias, out, x, nullptr, batch, dim, L, N, G, softplus, s, ws, ws_bytes,
                                force_split, stream);
}

template <typename T>
static int scan_bwd_dispatch(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                             const float *bias, const void *dout, void *du, void *ddelta, float *dA, float *dB, float *dC,
                             float *dD, float *dbias, int batch, int dim, int L, int N, int G, int softplus, void *ws,
                             size_t ws_bytes, int force_split, cudaStream_t stream) {
  const sigma_scan_strides st = contiguous_strides(dim, L, N, G);
  const bool al = (((uintptr_t)dout | (uintptr_t)du | (uintptr_t)ddelta | (uintptr_t)dB | (uintptr_t)dC) & 15) == 0;
  if (!force_generic() && al && scan_op_tma_eligible<T>(u, delta, B, C, du, dim, L, N, G, st))
    return scan_op_bwd_tma<T>(u, delta, A, B, C, D, bias, dout, du, ddelta, dA, dB, dC, dD, dbias, batch, dim, L, N, G, softplus,
                              ws, ws_bytes, force_split, stream);
  if constexpr (sizeof(T) == 2) {
    const size_t core = align256(scan_op_bwd_tma_workspace_bytes(batch, dim, L, N, 4));
    if (!force_generic() && widen_shape_ok(dim, L, N, G, 2) && ws_bytes >= core + widen_bwd_bytes(batch, dim, L, N, G)) {
      char *w = (char *)ws + core;
      const size_t bdl = align256((size_t)batch * dim * L * 4), bgn = align256((size_t)batch * G * N * L * 4);
      float *u32 = (float *)w, *d32 = (float *)(w + bdl), *g32 = (float *)(w + 2 * bdl), *du32 = (float *)(w + 3 * bdl), *dd32 = (float *)(w + 4 * bdl);
      float *B32 = (float *)(w + 5 * bdl), *C32 = (float *)(w + 5 * bdl + bgn);
      int rc;
      if ((rc = widen<T>(u, u32, batch, dim, 1, L, st.u_batch, st.u_dim, 0, stream))) return rc;
      if ((rc = widen<T>(delta, d32, batch, dim, 1, L, st.u_batch, st.u_dim, 0, stream))) return rc;
      if ((rc = widen<T>(dout, g32, batch, dim, 1, L, st.u_batch, st.u_dim, 0, stream))) return rc;
      if ((rc = widen<T>(B, B32, batch, G, N, L, st.B_batch, st.B_group, st.B_dstate, stream))) return rc;
      if ((rc = widen<T>(C, C32, batch, G, N, L, st.C_batch, st.C_group, st.C_dstate, stream))) return rc;
      if ((rc = scan_op_bwd_tma<float>(u32, d32, A, B32, C32, D, bias, g32, du32, dd32, dA, dB, dC, dD, dbias, batch, dim, L, N, G, softplus, ws,
                                       core, force_split, stream))) return rc;
      if ((rc = narrow_to<T>(du32, du, batch, dim, L, st.u_batch, st.u_dim, stream))) return rc;
      return narrow_to<T>(dd32, ddelta, batch, dim, L, st.u_batch, st.u_dim, stream);
    }
  }
  return scan_op_bwd_generic<T>(u, delta, A, B, C, D, bias, dout, du, ddelta, dA, dB, dC, dD, dbias, batch, dim, L, N, G,
                                softplus, ws, ws_bytes, stream);
}

}  // namespace sigma

using namespace sigma;

extern "C" {
#pragma GCC visibility push(default)

int sigma_abi_version(void) { return 1; }
const char *sigma_last_error(void) { return g_err; }
uint64_t sigma_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t sigma_scan_fwd_workspace_bytes(int batch, int dim, int seqlen, int dstate, int ngroups, int dtype) {
  // every element type is read natively: only the L-segment carries need scratch — plus, for 16-bit calls whose rows are not
  // 16-byte aligned but whose fp32 image is TMA-eligible, the widened operands
  size_t w = align256(std::max(scan_op_workspace_bytes(batch, dim, dstate), scan_op_tma_workspace_bytes(batch, dim, dstate)));
  if (dtype != SIGMA_F32 && widen_shape_ok(dim, seqlen, dstate, ngroups, 2)) w += widen_fwd_bytes(batch, dim, seqlen, dstate, ngroups);
  return w;
}

int sigma_scan_fwd(const void *u, const void *delta, const float *A, const void *B, const void *C,
                   const float *D, const float *delta_bias, void *out, float *x, int batch, int dim,
                   int seqlen, int dstate, int ngroups, int dtype, int delta_softplus,
                   const sigma_scan_strides *st, void *workspace, size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SIGMA_CHECK_ARG(u && delta && A && B && C && out && st, "sigma_scan_fwd: null pointer argument");
  SIGMA_CHECK_ARG(batch > 0 && dim > 0 && seqlen > 0 && dstate > 0 && ngroups > 0,
                  "sigma_scan_fwd: non-positive size (batch=%d dim=%d seqlen=%d dstate=%d ngroups=%d)",
                  batch, dim, seqlen, dstate, ngroups);
  SIGMA_CHECK_ARG(dstate <= 256, "sigma_scan_fwd: dstate=%d > 256 (selective_scan.cpp:198)", dstate);
  SIGMA_CHECK_ARG(dim % ngroups == 0, "sigma_scan_fwd: dim=%d not divisible by ngroups=%d", dim, ngroups);
  SIGMA_CHECK_ARG(dtype == SIGMA_F32 || dtype == SIGMA_F16 || dtype == SIGMA_BF16,
                  "sigma_scan_fwd: unknown dtype %d", dtype);
  if (dtype == SIGMA_F32)
    return scan_fwd_dispatch<float>(u, delta, A, B, C, D, delta_bias, out, x, batch, dim, seqlen, dstate, ngroups, delta_softplus,
                                    *st, workspace, workspace_bytes, 0, stream);
  if (dtype == SIGMA_F16)
    return scan_fwd_dispatch<__half>(u, delta, A, B, C, D, delta_bias, out, x, batch, dim, seqlen, dstate, ngroups, delta_softplus,
                                     *st, workspace, workspace_bytes, 0, stream);
  return scan_fwd_dispatch<__nv_bfloat16>(u, delta, A, B, C, D, delta_bias, out, x, batch, dim, seqlen, dstate, ngroups,
                                          delta_softplus, *st, workspace, workspace_bytes, 0, stream);
}

// test hook (not part of the drop-in surface): force the number of L-segments of the fp32 op kernel
int sigma_scan_fwd_f32_split(const float *u, const float *delta, const float *A, const float *B, const float *C,
                             const float *D, const float *delta_bias, float *out, float *x, int batch, int dim,
                             int seqlen, int dstate, int ngroups, int delta_softplus,
                             const sigma_scan_strides *st, void *workspace, size_t workspace_bytes,
                             int nsplit, void *stream) {
  SIGMA_CHECK_ARG(u && delta && A && B && C && out && st, "sigma_scan_fwd_f32_split: null pointer argument");
  SIGMA_CHECK_ARG(dim % ngroups == 0 && dstate <= 256, "sigma_scan_fwd_f32_split: bad dim/ngroups/dstate");
  return scan_fwd_dispatch<float>(u, delta, A, B, C, D, delta_bias, out, x, batch, dim, seqlen, dstate, ngroups,
                                  delta_softplus, *st, workspace, workspace_bytes, nsplit, (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"

// ---------------------------------------------------------------------------------------------
// fused channels-last pipeline
// ---------------------------------------------------------------------------------------------
namespace sigma {
int row_norm_launch(const RowNormParams &p, cudaStream_t stream);
int layernorm_bwd_launch(const float *x, const float *dy, const float *gamma, float *dx, float *dgamma, float *dbeta, long long rows,
                         int D, float eps, cudaStream_t stream);
int argmax_hist_launch(const float *logits, const void *labels, int label_bytes, unsigned long long *hist,
                       unsigned long long *counts, unsigned char *pred_out, int batch, int ncls, long long HW, cudaStream_t stream);
int dwconv3x3_silu_launch(const float *x, long long x_row_stride, long long x_batch_stride, const float *w,
                          const float *bias, float *y, long long y_batch_stride, int batch, int H, int W, int D,
                          cudaStream_t stream);
size_t ss2d_scan_workspace_bytes(int kind, int batch, int D, int N);
int ss2d_scan_fwd(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A,
                  const float *Ds, float *y, int batch, int H, int W, int D, int N, int R, int Cp, void *ws,
                  size_t ws_bytes, int force_split, cudaStream_t stream, float *dsave = nullptr, float *hsave = nullptr);
size_t ss2d_scan_hs_bytes(int kind, int batch, int H, int W, int D, int N);
int ss2d_pick_segments_hook(long long ctas, int nw, int ntiles, int N);
int gemm_pick_bn_hook(int N, long long m_tiles);
size_t ss2d_scan_bwd_workspace_bytes(int kind, int batch, int H, int W, int D, int N);
int ss2d_scan_bwd(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A, const float *Ds,
                  const float *dy, float *delta, float *dxc, float *ddelta, float *dxdbl, float *dA, float *dDs, float *ddtb, int batch,
                  int H, int W, int D, int N, int R, int Cp, void *ws, size_t ws_bytes, int force_split, cudaStream_t stream,
                  const float *hs_saved = nullptr);
int upsample2x_norm_launch(const float *in, const float *gamma, const float *beta, const float *wcls, int ncls, float *out,
                           int B, int Hin, int Win, int C, float eps, cudaStream_t stream);
int pool_avgmax_partial_launch(const float *x, float *partial, int B, long long L, int C, int nslice, cudaStream_t stream);
int scale_add_launch(const float *a, const float *sa, const float *b, const float *sb, float *out, long long rows,
                     long long rows_per_batch, int C, cudaStream_t stream);
int gemm_tf32_launch(const float *A, long long lda, const float *W, const float *W_lo, const float *bias, const float *residual,
                     long long ldr, const float *rscale, float *C, long long ldc, long long M, int N, int K, cudaStream_t stream);
int split_tf32_launch(const float *x, float *hi, float *lo, long long n, cudaStream_t stream);
int conv3x3_tf32_launch(const float *x, const float *W9, const float *W9_lo, const float *bias, int act, float *y, int B, int H, int W,
                        int Cin, int Cout, cudaStream_t stream);
struct ImagePreParams {
  const unsigned char *src; float *dst; const unsigned char *lsrc; long long *ldst;
  int H0, W0, SH, SW, OH, OW, off_y, off_x, mirror_src, mirror_out, label_pad;
  int clip_y0, clip_x0, clip_y1, clip_x1;
  double scale_y, scale_x, mean[3], stdv[3];
};
int image_pre_launch(const ImagePreParams &p, cudaStream_t stream);
int eval_exp_accumulate_launch(const float *logits, const float *logits_flip, float *acc, int ncls, int TH, int TW, int m_top, int m_left,
                               int vh, int vw, int AH, int AW, int ay, int ax, cudaStream_t stream);
int eval_resize_add_launch(const float *acc, int ncls, int AH, int AW, int m_top, int m_left, int SH, int SW, double *out, int H0, int W0,
                           cudaStream_t stream);
int eval_argmax_hist_launch(const double *score, const unsigned char *labels, unsigned char *pred, unsigned long long *hist,
                            unsigned long long *counts, int ncls, long long HW, cudaStream_t stream);
static bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }
}  // namespace sigma

extern "C" {
#pragma GCC visibility push(default)

int sigma_layernorm_fwd(const float *x, const float *w, const float *b, float *y, int64_t rows, int C, float eps,
                        void *stream) {
  SIGMA_CHECK_ARG(x && w && b && y, "sigma_layernorm_fwd: null pointer");
  SIGMA_CHECK_ARG(C > 0 && C % 4 == 0 && rows >= 0, "sigma_layernorm_fwd: C=%d must be a positive multiple of 4", C);
  SIGMA_CHECK_ARG(al16(x) && al16(w) && al16(b) && al16(y), "sigma_layernorm_fwd: pointers must be 16-byte aligned");
  RowNormParams p{x, 0, 1, w, b, nullptr, 0, nullptr, y, rows, rows > 0 ? rows : 1, 0, 0, C, C, eps};
  return row_norm_launch(p, (cudaStream_t)stream);
}

int sigma_layernorm_bwd(const float *x, const float *dy, const float *w, float *dx, float *dw, float *db, int64_t rows, int C, float eps,
                        void *stream) {
  SIGMA_CHECK_ARG(x && dy && w && dx && dw && db, "sigma_layernorm_bwd: null pointer");
  SIGMA_CHECK_ARG(C > 0 && C % 4 == 0 && rows >= 0, "sigma_layernorm_bwd: C=%d must be a positive multiple of 4", C);
  SIGMA_CHECK_ARG(al16(x) && al16(dy) && al16(w) && al16(dx), "sigma_layernorm_bwd: pointers must be 16-byte aligned");
  return layernorm_bwd_launch(x, dy, w, dx, dw, db, rows, C, eps, (cudaStream_t)stream);
}

int sigma_patch_merge_norm_fwd(const float *x, const float *w, const float *b, float *y, int batch, int H, int W, int C,
                               float eps, void *stream) {
  SIGMA_CHECK_ARG(x && w && b && y, "sigma_patch_merge_norm_fwd: null pointer");
  SIGMA_CHECK_ARG(batch > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "sigma_patch_merge_norm_fwd: bad sizes");
  SIGMA_CHECK_ARG(al16(x) && al16(w) && al16(b) && al16(y), "sigma_patch_merge_norm_fwd: pointers must be 16-byte aligned");
  const int64_t rows = (int64_t)batch * ((H + 1) / 2) * ((W + 1) / 2);
  RowNormParams p{x, 0, 1, w, b, nullptr, 0, nullptr, y, rows, rows, 0, 0, 4 * C, 4 * C, eps};
  p.mode = 1; p.gH = H; p.gW = W;
  return row_norm_launch(p, (cudaStream_t)stream);
}

int sigma_pixel_shuffle_norm_fwd(const float *x, const float *w, const float *b, float *y, int batch, int H, int W, int C,
                                 float eps, void *stream) {
  SIGMA_CHECK_ARG(x && w && b && y, "sigma_pixel_shuffle_norm_fwd: null pointer");
  SIGMA_CHECK_ARG(batch > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "sigma_pixel_shuffle_norm_fwd: bad sizes");
  SIGMA_CHECK_ARG(al16(x) && al16(w) && al16(b) && al16(y), "sigma_pixel_shuffle_norm_fwd: pointers must be 16-byte aligned");
  const int64_t rows = (int64_t)batch * H * W * 4;
  RowNormParams p{x, 0, 1, w, b, nullptr, 0, nullptr, y, rows, rows, 0, 0, C, C, eps};
  p.mode = 2; p.gH = H; p.gW = W;
  return row_norm_launch(p, (cudaStream_t)stream);
}

int sigma_merge_norm_gate_fwd(const float *y, int K, int64_t k_stride, int64_t in_batch_stride, const float *gamma,
                              const float *beta, const float *z, int64_t z_row_stride, const float *gate, float *out,
                              int64_t out_batch_stride, int64_t out_row_stride, int64_t rows, int64_t rows_per_batch,
                              int D, float eps, void *stream) {
  SIGMA_CHECK_ARG(y && gamma && beta && out, "sigma_merge_norm_gate_fwd: null pointer");
  SIGMA_CHECK_ARG(K >= 1 && K <= 8 && D > 0 && D % 4 == 0 && rows >= 0 && rows_per_batch > 0,
                  "sigma_merge_norm_gate_fwd: bad sizes K=%d D=%d rows=%lld rows_per_batch=%lld", K, D, (long long)rows,
                  (long long)rows_per_batch);
  SIGMA_CHECK_ARG(al16(y) && al16(gamma) && al16(beta) && al16(out) && al16(z) && al16(gate) && k_stride % 4 == 0 &&
                      in_batch_stride % 4 == 0 && out_batch_stride % 4 == 0 && out_row_stride % 4 == 0 &&
                      z_row_stride % 4 == 0,
                  "sigma_merge_norm_gate_fwd: pointers / strides must be 16-byte aligned");
  RowNormParams p{y, k_stride, K, gamma, beta, z, z_row_stride, gate, out, rows, rows_per_batch, in_batch_stride,
                  out_batch_stride, out_row_stride, D, eps};
  return row_norm_launch(p, (cudaStream_t)stream);
}

int sigma_dwconv3x3_silu_fwd(const float *x, int64_t x_row_stride, int64_t x_batch_stride, const float *w,
                             const float *bias, float *y, int64_t y_batch_stride, int batch, int H, int W, int D,
                             void *stream) {
  SIGMA_CHECK_ARG(x && w && y, "sigma_dwconv3x3_silu_fwd: null pointer");
  SIGMA_CHECK_ARG(batch > 0 && H > 0 && W > 0 && D > 0 && D % 4 == 0, "sigma_dwconv3x3_silu_fwd: bad sizes");
  SIGMA_CHECK_ARG(al16(x) && al16(y) && x_row_stride % 4 == 0 && x_batch_stride % 4 == 0 && y_batch_stride % 4 == 0,
                  "sigma_dwconv3x3_silu_fwd: pointers / strides must be 16-byte aligned");
  return dwconv3x3_silu_launch(x, x_row_stride, x_batch_stride, w, bias, y, y_batch_stride, batch, H, W, D,
                               (cudaStream_t)stream);
}

int sigma_ss2d_padded_cp(int N, int R) {
  const int opts[] = {4, 8, 12, 16, 24, 32, 48, 64};
  for (int o : opts)
    if (R <= o) return 2 * N + o;
  return -1;
}

size_t sigma_ss2d_scan_workspace_bytes(int kind, int batch, int H, int W, int D, int N) {
  (void)H; (void)W;
  return ss2d_scan_workspace_bytes(kind, batch, D, N);
}

static int ss2d_check(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A,
                      const float *Ds, float *y, int batch, int H, int W, int D, int N, int R, int Cp) {
  SIGMA_CHECK_ARG(xc && xdbl && dtw && dtb && A && Ds && y, "sigma_ss2d_scan_fwd: null pointer");
  SIGMA_CHECK_ARG(kind == SIGMA_DIRS_CROSS4 || kind == SIGMA_DIRS_SEQ2 || kind == SIGMA_DIRS_CROSS,
                  "sigma_ss2d_scan_fwd: unknown kind %d", kind);
  SIGMA_CHECK_ARG(batch > 0 && H > 0 && W > 0 && D > 0 && D % 4 == 0 && R > 0, "sigma_ss2d_scan_fwd: bad sizes");
  SIGMA_CHECK_ARG(kind != SIGMA_DIRS_CROSS || batch % 2 == 0, "sigma_ss2d_scan_fwd: CROSS needs batch = 2·images");
  SIGMA_CHECK_ARG(N == 4 || N == 8 || N == 16, "sigma_ss2d_scan_fwd: d_state=%d unsupported (4, 8, 16)", N);
  SIGMA_CHECK_ARG(Cp == sigma_ss2d_padded_cp(N, R), "sigma_ss2d_scan_fwd: Cp=%d must equal sigma_ss2d_padded_cp(N=%d, R=%d)=%d",
                  Cp, N, R, sigma_ss2d_padded_cp(N, R));
  SIGMA_CHECK_ARG(al16(xc) && al16(xdbl) && al16(y), "sigma_ss2d_scan_fwd: xc / xdbl / y must be 16-byte aligned");
  return SIGMA_OK;
}

int sigma_ss2d_scan_fwd(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb,
                        const float *A, const float *Ds, float *y, int batch, int H, int W, int D, int N, int R, int Cp,
                        void *workspace, size_t workspace_bytes, void *stream) {
  int rc = ss2d_check(kind, xc, xdbl, dtw, dtb, A, Ds, y, batch, H, W, D, N, R, Cp);
  if (rc) return rc;
  return ss2d_scan_fwd(kind, xc, xdbl, dtw, dtb, A, Ds, y, batch, H, W, D, N, R, Cp, workspace, workspace_bytes, 0,
                       (cudaStream_t)stream);
}

// test hook: force the number of L-segments
int sigma_ss2d_scan_fwd_split(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb,
                              const float *A, const float *Ds, float *y, int batch, int H, int W, int D, int N, int R,
                              int Cp, void *workspace, size_t workspace_bytes, int nsplit, void *stream) {
  int rc = ss2d_check(kind, xc, xdbl, dtw, dtb, A, Ds, y, batch, H, W, D, N, R, Cp);
  if (rc) return rc;
  return ss2d_scan_fwd(kind, xc, xdbl, dtw, dtb, A, Ds, y, batch, H, W, D, N, R, Cp, workspace, workspace_bytes, nsplit,
                       (cudaStream_t)stream);
}

// test hooks (host logic only, no CUDA call): the launch heuristics, so that CPU tests can hold them to the recorded sweeps
int sigma_test_pick_segments(long long ctas, int warps_per_cta, int ntiles, int N) { return ss2d_pick_segments_hook(ctas, warps_per_cta, ntiles, N); }
int sigma_test_pick_bn(int N, long long m_tiles) { return gemm_pick_bn_hook(N, m_tiles); }

// training forward: the forward plus what the fused backward needs (delta' slabs, block-start states)
size_t sigma_ss2d_scan_hs_bytes(int kind, int batch, int H, int W, int D, int N) {
  if (kind != SIGMA_DIRS_CROSS4 && kind != SIGMA_DIRS_SEQ2) return 0;
  return ss2d_scan_hs_bytes(kind, batch, H, W, D, N);
}

int sigma_ss2d_scan_fwd_save(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A,
                             const float *Ds, float *y, float *delta, float *hs, int batch, int H, int W, int D, int N, int R, int Cp,
                             void *workspace, size_t workspace_bytes, int nsplit, void *stream) {
  int rc = ss2d_check(kind, xc, xdbl, dtw, dtb, A, Ds, y, batch, H, W, D, N, R, Cp);
  if (rc) return rc;
  SIGMA_CHECK_ARG(delta && hs && al16(delta) && al16(hs), "sigma_ss2d_scan_fwd_save: delta / hs must be non-null and 16-byte aligned");
  SIGMA_CHECK_ARG(kind == SIGMA_DIRS_CROSS4 || kind == SIGMA_DIRS_SEQ2, "sigma_ss2d_scan_fwd_save: kind %d unsupported (CROSS4, SEQ2)", kind);
  SIGMA_CHECK_ARG(N == 4 || N == 16, "sigma_ss2d_scan_fwd_save: d_state=%d unsupported (4, 16)", N);
  return ss2d_scan_fwd(kind, xc, xdbl, dtw, dtb, A, Ds, y, batch, H, W, D, N, R, Cp, workspace, workspace_bytes, nsplit,
                       (cudaStream_t)stream, delta, hs);
}

size_t sigma_ss2d_scan_bwd_workspace_bytes(int kind, int batch, int H, int W, int D, int N) {
  if (kind != SIGMA_DIRS_CROSS4 && kind != SIGMA_DIRS_SEQ2) return 0;
  return ss2d_scan_bwd_workspace_bytes(kind, batch, H, W, D, N);
}

static int ss2d_bwd_entry(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A, const float *Ds,
                          const float *dy, float *delta, float *dxc, float *ddelta, float *dxdbl, float *dA, float *dDs, float *ddtb,
                          int batch, int H, int W, int D, int N, int R, int Cp, void *ws, size_t wsb, int nsplit, void *stream,
                          const float *hs_saved = nullptr) {
  SIGMA_CHECK_ARG(xc && xdbl && dtw && dtb && A && Ds && dy && delta && dxc && ddelta && dxdbl && dA && dDs && ddtb, "sigma_ss2d_scan_bwd: null pointer");
  SIGMA_CHECK_ARG(kind == SIGMA_DIRS_CROSS4 || kind == SIGMA_DIRS_SEQ2, "sigma_ss2d_scan_bwd: kind %d unsupported (CROSS4, SEQ2)", kind);
  SIGMA_CHECK_ARG(batch > 0 && H > 0 && W > 0 && D > 0 && D % 64 == 0 && R > 0, "sigma_ss2d_scan_bwd: bad sizes (D=%d must be a multiple of 64)", D);
  SIGMA_CHECK_ARG(N == 4 || N == 16, "sigma_ss2d_scan_bwd: d_state=%d unsupported (4, 16)", N);
  SIGMA_CHECK_ARG(Cp == sigma_ss2d_padded_cp(N, R), "sigma_ss2d_scan_bwd: Cp=%d must equal sigma_ss2d_padded_cp(N=%d, R=%d)", Cp, N, R);
  SIGMA_CHECK_ARG(al16(xc) && al16(xdbl) && al16(dy) && al16(delta) && al16(dxc) && al16(ddelta) && al16(dxdbl), "sigma_ss2d_scan_bwd: pointers must be 16-byte aligned");
  SIGMA_CHECK_ARG(hs_saved == nullptr || al16(hs_saved), "sigma_ss2d_scan_bwd_saved: hs must be 16-byte aligned");
  return ss2d_scan_bwd(kind, xc, xdbl, dtw, dtb, A, Ds, dy, delta, dxc, ddelta, dxdbl, dA, dDs, ddtb, batch, H, W, D, N, R, Cp, ws, wsb, nsplit,
                       (cudaStream_t)stream, hs_saved);
}

// backward after sigma_ss2d_scan_fwd_save: `delta` and `hs` are INPUTS (what that call wrote); no state sweep runs
int sigma_ss2d_scan_bwd_saved(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A, const float *Ds,
                              const float *dy, const float *delta, const float *hs, float *dxc, float *ddelta, float *dxdbl, float *dA,
                              float *dDs, float *ddtb, int batch, int H, int W, int D, int N, int R, int Cp, void *workspace,
                              size_t workspace_bytes, int nsplit, void *stream) {
  SIGMA_CHECK_ARG(hs != nullptr, "sigma_ss2d_scan_bwd_saved: null hs");
  return ss2d_bwd_entry(kind, xc, xdbl, dtw, dtb, A, Ds, dy, const_cast<float *>(delta), dxc, ddelta, dxdbl, dA, dDs, ddtb, batch, H, W, D, N, R,
                        Cp, workspace, workspace_bytes, nsplit, stream, hs);
}

int sigma_ss2d_scan_bwd(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A, const float *Ds,
                        const float *dy, float *delta, float *dxc, float *ddelta, float *dxdbl, float *dA, float *dDs, float *ddtb, int batch,
                        int H, int W, int D, int N, int R, int Cp, void *workspace, size_t workspace_bytes, void *stream) {
  return ss2d_bwd_entry(kind, xc, xdbl, dtw, dtb, A, Ds, dy, delta, dxc, ddelta, dxdbl, dA, dDs, ddtb, batch, H, W, D, N, R, Cp, workspace,
                        workspace_bytes, 0, stream);
}

// test hook: force the number of L-segments
int sigma_ss2d_scan_bwd_split(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A, const float *Ds,
                              const float *dy, float *delta, float *dxc, float *ddelta, float *dxdbl, float *dA, float *dDs, float *ddtb,
                              int batch, int H, int W, int D, int N, int R, int Cp, void *workspace, size_t workspace_bytes, int nsplit,
                              void *stream) {
  return ss2d_bwd_entry(kind, xc, xdbl, dtw, dtb, A, Ds, dy, delta, dxc, ddelta, dxdbl, dA, dDs, ddtb, batch, H, W, D, N, R, Cp, workspace,
                        workspace_bytes, nsplit, stream);
}

int sigma_upsample2x_norm_fwd(const float *x, const float *w, const float *b, float *y, int batch, int H, int W, int C,
                              float eps, void *stream) {
  SIGMA_CHECK_ARG(x && y && ((w && b) || (!w && !b)), "sigma_upsample2x_norm_fwd: null pointer (w and b may be NULL together)");
  SIGMA_CHECK_ARG(batch > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "sigma_upsample2x_norm_fwd: bad sizes");
  SIGMA_CHECK_ARG(al16(x) && al16(w) && al16(b) && al16(y), "sigma_upsample2x_norm_fwd: pointers must be 16-byte aligned");
  if (!w) return upsample2x_norm_launch(x, nullptr, nullptr, nullptr, 0, y, batch, H, W, C, eps, (cudaStream_t)stream);
  return upsample2x_norm_launch(x, w, b, nullptr, 0, y, batch, H, W, C, eps, (cudaStream_t)stream);
}

int sigma_upsample2x_norm_head_fwd(const float *x, const float *w, const float *b, const float *wcls, int num_classes,
                                   float *logits, int batch, int H, int W, int C, float eps, void *stream) {
  SIGMA_CHECK_ARG(x && w && b && wcls && logits, "sigma_upsample2x_norm_head_fwd: null pointer");
  SIGMA_CHECK_ARG(batch > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && num_classes > 0,
                  "sigma_upsample2x_norm_head_fwd: bad sizes");
  SIGMA_CHECK_ARG(al16(x) && al16(w) && al16(b) && al16(wcls), "sigma_upsample2x_norm_head_fwd: pointers must be 16-byte aligned");
  return upsample2x_norm_launch(x, w, b, wcls, num_classes, logits, batch, H, W, C, eps, (cudaStream_t)stream);
}

int sigma_argmax_hist_fwd(const float *logits, const void *labels, int label_bytes, uint64_t *hist, uint64_t *counts,
                          uint8_t *pred, int batch, int num_classes, int64_t HW, void *stream) {
  SIGMA_CHECK_ARG(logits && labels && hist && counts, "sigma_argmax_hist_fwd: null pointer");
  SIGMA_CHECK_ARG(batch > 0 && HW > 0 && num_classes > 0 && num_classes <= 238, "sigma_argmax_hist_fwd: bad sizes (1 <= classes <= 238: the per-CTA histogram lives in shared memory)");
  return argmax_hist_launch(logits, labels, label_bytes, (unsigned long long *)hist, (unsigned long long *)counts, pred, batch,
                            num_classes, HW, (cudaStream_t)stream);
}

int sigma_pool_avgmax_partial_fwd(const float *x, float *partial, int batch, int64_t L, int C, int nslice, void *stream) {
  SIGMA_CHECK_ARG(x && partial, "sigma_pool_avgmax_partial_fwd: null pointer");
  SIGMA_CHECK_ARG(batch > 0 && L > 0 && C > 0 && C % 4 == 0 && nslice > 0 && nslice <= 65535 && al16(x),
                  "sigma_pool_avgmax_partial_fwd: bad sizes / alignment");
  return pool_avgmax_partial_launch(x, partial, batch, L, C, nslice, (cudaStream_t)stream);
}

int sigma_scale_add_fwd(const float *a, const float *sa, const float *b, const float *sb, float *out, int64_t rows,
                        int64_t rows_per_batch, int C, void *stream) {
  SIGMA_CHECK_ARG(b && sb && out && (a == nullptr || sa != nullptr), "sigma_scale_add_fwd: null pointer");
  SIGMA_CHECK_ARG(rows >= 0 && rows_per_batch > 0 && C > 0 && C % 4 == 0 && al16(a) && al16(sa) && al16(b) && al16(sb) && al16(out),
                  "sigma_scale_add_fwd: bad sizes / alignment");
  return scale_add_launch(a, sa, b, sb, out, rows, rows_per_batch, C, (cudaStream_t)stream);
}

size_t sigma_scan_bwd_workspace_bytes(int batch, int dim, int seqlen, int dstate, int ngroups, int dtype) {
  const int eb = dtype == SIGMA_F32 ? 4 : 2;
  size_t w = align256(std::max(scan_op_bwd_workspace_bytes(batch, dim, seqlen, dstate, eb),
                               scan_op_bwd_tma_workspace_bytes(batch, dim, seqlen, std::min(dstate, 16), eb)));
  if (dtype != SIGMA_F32 && widen_shape_ok(dim, seqlen, dstate, ngroups, 2)) w += widen_bwd_bytes(batch, dim, seqlen, dstate, ngroups);
  return w;
}

static int scan_bwd_entry(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                          const float *delta_bias, const void *dout, void *du, void *ddelta, float *dA, float *dB, float *dC,
                          float *dD, float *ddelta_bias, int batch, int dim, int seqlen, int dstate, int ngroups, int dtype,
                          int delta_softplus, void *workspace, size_t workspace_bytes, int force_split, cudaStream_t stream) {
  SIGMA_CHECK_ARG(u && delta && A && B && C && dout && du && ddelta && dA && dB && dC, "sigma_scan_bwd: null pointer argument");
  SIGMA_CHECK_ARG((D == nullptr || dD != nullptr) && (delta_bias == nullptr || ddelta_bias != nullptr),
                  "sigma_scan_bwd: dD / ddelta_bias required when D / delta_bias are given");
  SIGMA_CHECK_ARG(batch > 0 && dim > 0 && seqlen > 0 && dstate > 0 && ngroups > 0 && dim % ngroups == 0,
                  "sigma_scan_bwd: bad sizes (batch=%d dim=%d seqlen=%d dstate=%d ngroups=%d)", batch, dim, seqlen, dstate, ngroups);
  SIGMA_CHECK_ARG(dtype == SIGMA_F32 || dtype == SIGMA_F16 || dtype == SIGMA_BF16, "sigma_scan_bwd: unknown dtype %d", dtype);
  if (dstate > 16) { set_error("sigma_scan_bwd: d_state=%d > 16 is not supported by the backward kernels", dstate); return SIGMA_EUNSUPPORTED; }
  const size_t need = sigma_scan_bwd_workspace_bytes(batch, dim, seqlen, dstate, ngroups, dtype);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("sigma_scan_bwd: needs %zu workspace bytes, got %zu", need, workspace_bytes);
    return SIGMA_EWORKSPACE;
  }
  if (dtype == SIGMA_F32)
    return scan_bwd_dispatch<float>(u, delta, A, B, C, D, delta_bias, dout, du, ddelta, dA, dB, dC, dD, ddelta_bias, batch, dim,
                                    seqlen, dstate, ngroups, delta_softplus, workspace, workspace_bytes, force_split, stream);
  if (dtype == SIGMA_F16)
    return scan_bwd_dispatch<__half>(u, delta, A, B, C, D, delta_bias, dout, du, ddelta, dA, dB, dC, dD, ddelta_bias, batch, dim,
                                     seqlen, dstate, ngroups, delta_softplus, workspace, workspace_bytes, force_split, stream);
  return scan_bwd_dispatch<__nv_bfloat16>(u, delta, A, B, C, D, delta_bias, dout, du, ddelta, dA, dB, dC, dD, ddelta_bias, batch,
                                          dim, seqlen, dstate, ngroups, delta_softplus, workspace, workspace_bytes, force_split,
                                          stream);
}

int sigma_scan_bwd(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                   const float *delta_bias, const void *dout, void *du, void *ddelta, float *dA, float *dB, float *dC,
                   float *dD, float *ddelta_bias, int batch, int dim, int seqlen, int dstate, int ngroups, int dtype,
                   int delta_softplus, void *workspace, size_t workspace_bytes, void *stream_) {
  return scan_bwd_entry(u, delta, A, B, C, D, delta_bias, dout, du, ddelta, dA, dB, dC, dD, ddelta_bias, batch, dim, seqlen,
                        dstate, ngroups, dtype, delta_softplus, workspace, workspace_bytes, 0, (cudaStream_t)stream_);
}

// test hook: force the number of L-segments of the backward
int sigma_scan_bwd_split(const void *u, const void *delta, const float *A, const void *B, const void *C, const float *D,
                         const float *delta_bias, const void *dout, void *du, void *ddelta, float *dA, float *dB, float *dC,
                         float *dD, float *ddelta_bias, int batch, int dim, int seqlen, int dstate, int ngroups, int dtype,
                         int delta_softplus, void *workspace, size_t workspace_bytes, int nsplit, void *stream_) {
  return scan_bwd_entry(u, delta, A, B, C, D, delta_bias, dout, du, ddelta, dA, dB, dC, dD, ddelta_bias, batch, dim, seqlen,
                        dstate, ngroups, dtype, delta_softplus, workspace, workspace_bytes, nsplit, (cudaStream_t)stream_);
}

int sigma_linear_tf32(const float *A, int64_t lda, const float *W, const float *bias, const float *residual, int64_t ldr,
                      const float *rscale, float *C, int64_t ldc, int64_t M, int N, int K, void *stream) {
  SIGMA_CHECK_ARG(A && W && C, "sigma_linear_tf32: null pointer");
  SIGMA_CHECK_ARG(M >= 0 && M < (1LL << 31) && N > 0 && K > 0, "sigma_linear_tf32: bad sizes M=%lld N=%d K=%d", (long long)M, N, K);
  SIGMA_CHECK_ARG(K % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldc % 4 == 0 && (residual == nullptr || ldr % 4 == 0) && lda >= K && ldc >= N,
                  "sigma_linear_tf32: K, N, lda, ldc, ldr must be multiples of 4 floats (16-byte TMA boxes; the epilogue reads bias / "
                  "rscale / residual as float4)");
  SIGMA_CHECK_ARG(al16(A) && al16(W) && al16(C) && al16(bias) && al16(residual) && al16(rscale),
                  "sigma_linear_tf32: pointers must be 16-byte aligned");
  SIGMA_CHECK_ARG(rscale == nullptr || residual != nullptr, "sigma_linear_tf32: rscale without residual");
  return gemm_tf32_launch(A, lda, W, nullptr, bias, residual, ldr, rscale, C, ldc, M, N, K, (cudaStream_t)stream);
}

int sigma_linear_tf32x3(const float *A, int64_t lda, const float *W_hi, const float *W_lo, const float *bias, const float *residual,
                        int64_t ldr, const float *rscale, float *C, int64_t ldc, int64_t M, int N, int K, void *stream) {
  SIGMA_CHECK_ARG(A && W_hi && W_lo && C, "sigma_linear_tf32x3: null pointer");
  SIGMA_CHECK_ARG(M >= 0 && M < (1LL << 31) && N > 0 && K > 0, "sigma_linear_tf32x3: bad sizes M=%lld N=%d K=%d", (long long)M, N, K);
  SIGMA_CHECK_ARG(K % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldc % 4 == 0 && (residual == nullptr || ldr % 4 == 0) && lda >= K && ldc >= N,
                  "sigma_linear_tf32x3: K, N, lda, ldc, ldr must be multiples of 4 floats");
  SIGMA_CHECK_ARG(al16(A) && al16(W_hi) && al16(W_lo) && al16(C) && al16(bias) && al16(residual) && al16(rscale),
                  "sigma_linear_tf32x3: pointers must be 16-byte aligned");
  SIGMA_CHECK_ARG(rscale == nullptr || residual != nullptr, "sigma_linear_tf32x3: rscale without residual");
  return gemm_tf32_launch(A, lda, W_hi, W_lo, bias, residual, ldr, rscale, C, ldc, M, N, K, (cudaStream_t)stream);
}

int sigma_conv3x3_tf32(const float *x, const float *w9, const float *w9_lo, const float *bias, int act, float *y, int batch, int H, int W,
                       int Cin, int Cout, void *stream) {
  SIGMA_CHECK_ARG(x && w9 && y, "sigma_conv3x3_tf32: null pointer");
  SIGMA_CHECK_ARG(batch >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && Cin % 4 == 0 && Cout % 4 == 0,
                  "sigma_conv3x3_tf32: bad sizes (Cin=%d, Cout=%d must be multiples of 4)", Cin, Cout);
  SIGMA_CHECK_ARG(act == 0 || act == 1, "sigma_conv3x3_tf32: act must be 0 (none) or 1 (GELU)");
  SIGMA_CHECK_ARG(al16(x) && al16(w9) && al16(w9_lo) && al16(bias) && al16(y), "sigma_conv3x3_tf32: pointers must be 16-byte aligned");
  return conv3x3_tf32_launch(x, w9, w9_lo, bias, act, y, batch, H, W, Cin, Cout, (cudaStream_t)stream);
}

int sigma_split_tf32_fwd(const float *x, float *hi, float *lo, int64_t n, void *stream) {
  SIGMA_CHECK_ARG(x && hi && lo && n >= 0, "sigma_split_tf32_fwd: bad arguments");
  return split_tf32_launch(x, hi, lo, n, (cudaStream_t)stream);
}

int sigma_image_pre_fwd(const uint8_t *src, const uint8_t *labels, float *out, int64_t *labels_out, int H0, int W0, int SH, int SW,
                        double scale_y, double scale_x, int OH, int OW, int off_y, int off_x, int mirror_src, int mirror_out,
                        int label_pad, const int *clip4_host, const double *mean3, const double *std3, void *stream) {
  SIGMA_CHECK_ARG(src && out && mean3 && std3, "sigma_image_pre_fwd: null pointer");
  SIGMA_CHECK_ARG((labels == nullptr) == (labels_out == nullptr), "sigma_image_pre_fwd: labels and labels_out go together");
  SIGMA_CHECK_ARG(H0 > 0 && W0 > 0 && SH > 0 && SW > 0 && OH > 0 && OW > 0 && scale_y > 0 && scale_x > 0, "sigma_image_pre_fwd: bad sizes");
  ImagePreParams p;
  p.src = src; p.dst = out; p.lsrc = labels; p.ldst = (long long *)labels_out;
  p.H0 = H0; p.W0 = W0; p.SH = SH; p.SW = SW; p.OH = OH; p.OW = OW; p.off_y = off_y; p.off_x = off_x;
  p.mirror_src = mirror_src; p.mirror_out = mirror_out; p.label_pad = label_pad; p.scale_y = scale_y; p.scale_x = scale_x;
  p.clip_y0 = 0; p.clip_x0 = 0; p.clip_y1 = SH; p.clip_x1 = SW;
  if (clip4_host) {
    p.clip_y0 = std::max(0, clip4_host[0]); p.clip_x0 = std::max(0, clip4_host[1]);
    p.clip_y1 = std::min(SH, clip4_host[0] + clip4_host[2]); p.clip_x1 = std::min(SW, clip4_host[1] + clip4_host[3]);
  }
  for (int c = 0; c < 3; ++c) {
    SIGMA_CHECK_ARG(std3[c] != 0.0, "sigma_image_pre_fwd: std[%d] == 0", c);
    p.mean[c] = mean3[c]; p.stdv[c] = std3[c];
  }
  return image_pre_launch(p, (cudaStream_t)stream);
}

int sigma_eval_exp_accumulate_fwd(const float *logits, const float *logits_flip, float *acc, int ncls, int TH, int TW, int m_top,
                                  int m_left, int vh, int vw, int AH, int AW, int ay, int ax, void *stream) {
  SIGMA_CHECK_ARG(logits && acc, "sigma_eval_exp_accumulate_fwd: null pointer");
  SIGMA_CHECK_ARG(ncls > 0 && m_top >= 0 && m_left >= 0 && vh >= 0 && vw >= 0 && m_top + vh <= TH && m_left + vw <= TW && ay >= 0 &&
                      ax >= 0 && ay + vh <= AH && ax + vw <= AW,
                  "sigma_eval_exp_accumulate_fwd: window (%d+%d, %d+%d) of a %dx%d tile into (%d, %d) of %dx%d", m_top, vh, m_left, vw,
                  TH, TW, ay, ax, AH, AW);
  return eval_exp_accumulate_launch(logits, logits_flip, acc, ncls, TH, TW, m_top, m_left, vh, vw, AH, AW, ay, ax, (cudaStream_t)stream);
}

int sigma_eval_resize_add_fwd(const float *acc, int ncls, int AH, int AW, int m_top, int m_left, int SH, int SW, double *out, int H0,
                              int W0, void *stream) {
  SIGMA_CHECK_ARG(acc && out, "sigma_eval_resize_add_fwd: null pointer");
  SIGMA_CHECK_ARG(ncls > 0 && SH > 0 && SW > 0 && H0 > 0 && W0 > 0 && m_top >= 0 && m_left >= 0 && m_top + SH <= AH && m_left + SW <= AW,
                  "sigma_eval_resize_add_fwd: bad sizes");
  return eval_resize_add_launch(acc, ncls, AH, AW, m_top, m_left, SH, SW, out, H0, W0, (cudaStream_t)stream);
}

int sigma_eval_argmax_hist_fwd(const double *score, const uint8_t *labels, uint8_t *pred, uint64_t *hist, uint64_t *counts,
                               int num_classes, int64_t HW, void *stream) {
  SIGMA_CHECK_ARG(score && pred, "sigma_eval_argmax_hist_fwd: null pointer");
  SIGMA_CHECK_ARG(labels == nullptr || (hist && counts), "sigma_eval_argmax_hist_fwd: labels need hist and counts");
  SIGMA_CHECK_ARG(num_classes >= 1 && num_classes <= 255 && HW >= 0, "sigma_eval_argmax_hist_fwd: bad sizes");
  if (HW == 0) return SIGMA_OK;
  return eval_argmax_hist_launch(score, labels, pred, (unsigned long long *)hist, (unsigned long long *)counts, num_classes, HW,
                                 (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"
