/*
 * sigma_b200 — C-ABI of the B200-native SS2D / selective-scan hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer is a DEVICE pointer unless the
 * name ends in `_host`.  Every function launches on `stream` (a cudaStream_t passed as void*),
 * never synchronises, allocates nothing (scratch comes from the caller through `workspace`),
 * and returns 0 on success or a negative SIGMA_E* code; sigma_last_error() then holds a
 * human-readable message (thread-local).  No exceptions cross this boundary.
 *
 * The reference interface each entry point replaces is cited as file:line relative to the
 * reference repository (zifuwan/Sigma @ 5c619c6).  INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 */
#ifndef SIGMA_B200_H_
#define SIGMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIGMA_OK 0
#define SIGMA_EINVAL (-1)   /* bad shape / stride / dtype / null pointer            */
#define SIGMA_ECUDA (-2)    /* CUDA runtime or driver error (launch, tensor map...) */
#define SIGMA_EWORKSPACE (-3) /* workspace missing or too small                      */
#define SIGMA_EUNSUPPORTED (-4)

/* element types of u / delta / B / C / out (A, D, delta_bias and all states are fp32,
 * selective_scan.cpp:175-180) */
#define SIGMA_F32 0
#define SIGMA_F16 1
#define SIGMA_BF16 2

int sigma_abi_version(void);
const char *sigma_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py: gpu_launches) */
uint64_t sigma_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * a1. selective scan, op level.
 * Replaces `selective_scan_cuda_core.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus,
 * nrows) -> [out, x]`  (csrc/selective_scan/selective_scan.cpp:165-249, kernel
 * selective_scan_fwd_kernel.cuh:64-206).
 *
 *   delta' = softplus?(delta + delta_bias[d]);  h[n,l] = exp(delta'·A[d,n])·h[n,l-1] + delta'·B[g,n,l]·u[l]
 *   out[d,l] = D[d]·u[l] + Σ_n C[g,n,l]·h[n,l],   g = d / (dim / ngroups)
 *
 * Layout: u, delta, out (batch, dim, seqlen) with unit stride along seqlen and the element
 * strides given below; A (dim, dstate) any strides; B, C (batch, ngroups, dstate, seqlen) unit
 * stride along seqlen.  `x` (nullable) receives the chunk-end states
 * (batch, dim, ceil(seqlen/2048), 2·dstate) fp32, interleaved (prod a, h), contiguous
 * (selective_scan.cpp:228, fwd_kernel.cuh:181-184).  D and delta_bias are nullable.
 * ------------------------------------------------------------------------------------------ */
typedef struct sigma_scan_strides {
  int64_t u_batch, u_dim;
  int64_t delta_batch, delta_dim;
  int64_t A_dim, A_dstate;
  int64_t B_batch, B_group, B_dstate;
  int64_t C_batch, C_group, C_dstate;
  int64_t out_batch, out_dim;
} sigma_scan_strides;

size_t sigma_scan_fwd_workspace_bytes(int batch, int dim, int seqlen, int dstate, int ngroups,
                                      int dtype);

int sigma_scan_fwd(const void *u, const void *delta, const float *A, const void *B, const void *C,
                   const float *D, const float *delta_bias, void *out, float *x,
                   int batch, int dim, int seqlen, int dstate, int ngroups, int dtype,
                   int delta_softplus, const sigma_scan_strides *strides,
                   void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * a3. backward of a1.
 * Replaces `selective_scan_cuda_core.bwd(u, delta, A, B, C, D, delta_bias, dout, x,
 * delta_softplus, nrows) -> [du, ddelta, dA, dB, dC, dD, ddelta_bias]`
 * (selective_scan.cpp:251-362, selective_scan_bwd_kernel.cuh:68-274).
 * All tensors contiguous.  du, ddelta: (batch, dim, seqlen) in `dtype`; dA (dim, dstate),
 * dD, ddelta_bias (dim) fp32 — OVERWRITTEN (not accumulated); dB, dC (batch, ngroups, dstate,
 * seqlen) fp32, overwritten.  dD / ddelta_bias may be NULL when D / delta_bias are NULL.
 * ------------------------------------------------------------------------------------------ */
size_t sigma_scan_bwd_workspace_bytes(int batch, int dim, int seqlen, int dstate, int ngroups, int dtype);

int sigma_scan_bwd(const void *u, const void *delta, const float *A, const void *B, const void *C,
                   const float *D, const float *delta_bias, const void *dout,
                   void *du, void *ddelta, float *dA, float *dB, float *dC, float *dD,
                   float *ddelta_bias,
                   int batch, int dim, int seqlen, int dstate, int ngroups, int dtype,
                   int delta_softplus, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * a4+a5 (+a8/a9 cores). Fused multi-direction SS2D scan, channels-last.
 * Replaces, in one launch, CrossScan (vmamba.py:80-98) + the dt_proj einsum (vmamba.py:199) +
 * delta_bias/softplus + SelectiveScan (vmamba.py:213) + the un-flip / un-transpose half of
 * CrossMerge (vmamba.py:100-108) of `cross_selective_scan` (vmamba.py:165-226); with
 * kind=SIGMA_DIRS_SEQ2 the K=2 core of `cross_selective_scan_multimodal_k2` (vmamba.py:369-430)
 * and with kind=SIGMA_DIRS_CROSS the two C-swapped scans of Cross_Mamba_Attention_SSM.forward
 * (vmamba.py:1528-1539).
 *
 *   xc    (batch, Lseq, D)          fp32, channels-last: the dwconv+SiLU output (Lseq = H·W, or
 *                                   2·H·W for SEQ2 = [rgb ‖ x] per image, or H·W per modality m
 *                                   with batch index = 2·b+m for CROSS)
 *   xdbl  (batch, Lseq, K, Cp)      fp32: x_proj output per POSITION and direction, row k =
 *                                   [dt_r (R) | B (N) | C (N) | pad], Cp % 4 == 0
 *   y     (K, batch, Lseq, D)       fp32: direction k's output stored at the POSITION it belongs
 *                                   to (so CrossMerge is a plain sum over k)
 *   dtw (K, D, R), dtb (K, D), A (K·D, N) [= -exp(A_logs)], Ds (K·D)
 * ------------------------------------------------------------------------------------------ */
#define SIGMA_DIRS_CROSS4 0 /* K=4: row-major, column-major, and both reversed (vmamba.py:86-88) */
#define SIGMA_DIRS_SEQ2 1   /* K=2: forward and reversed over a flat sequence (vmamba.py:130-131) */
#define SIGMA_DIRS_CROSS 2  /* K=1 per modality, C taken from the other modality (vmamba.py:1530,1536) */

size_t sigma_ss2d_scan_workspace_bytes(int kind, int batch, int H, int W, int D, int N);

int sigma_ss2d_scan_fwd(int kind, const float *xc, const float *xdbl, const float *dtw,
                        const float *dtb, const float *A, const float *Ds, float *y,
                        int batch, int H, int W, int D, int N, int R, int Cp,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Row-wise / elementwise pieces of a5-a11 (channels-last, fp32).
 * ------------------------------------------------------------------------------------------ */
/* nn.LayerNorm over the last dim (vmamba.py:1693,724,2173; eps=1e-5): y = (x-mean)/sqrt(var+eps)·w+b */
int sigma_layernorm_fwd(const float *x, const float *w, const float *b, float *y, int64_t rows,
                        int C, float eps, void *stream);

/* depthwise 3x3 conv (pad 1) + bias + SiLU, channels-last (vmamba.py:683-692,1072).
 * x rows have stride x_row_stride floats (so the x half of in_proj's output is read in place);
 * w is the nn.Conv2d weight (D,1,3,3) contiguous; y (batch,H,W,D) contiguous.            */
int sigma_dwconv3x3_silu_fwd(const float *x, int64_t x_row_stride, const float *w,
                             const float *bias, float *y, int batch, int H, int W, int D,
                             void *stream);

/* CrossMerge sum + out_norm LayerNorm + gate (vmamba.py:217-224,1077):
 *   yo[r,:] = LN(Σ_k y[k,r,:])·gamma+beta  ·  (z ? SiLU(z[r,:]) : 1)  ·  (gate ? gate[r / rows_per_gate, :] : 1)
 * z rows have stride z_row_stride (the z half of in_proj's output, read in place).           */
int sigma_merge_norm_gate_fwd(const float *y, int K, int64_t k_stride, const float *gamma,
                              const float *beta, const float *z, int64_t z_row_stride,
                              const float *gate, int64_t rows_per_gate, float *yo,
                              int64_t yo_row_stride, int64_t rows, int D, float eps, void *stream);

/* ------------------------------------------------------------------------------------------
 * Dense projections (in_proj / x_proj / out_proj / PatchMerging / PatchExpand linears,
 * vmamba.py:679,725,616,195; MambaDecoder.py:17,39,82-83): tcgen05 (5th-gen tensor core) TF32
 * GEMM, fp32 accumulate in TMEM, TMA-fed:  C[M,N] = A[M,K]·W[N,K]^T (+ bias[N]) (+ residual[M,N]).
 * A rows: lda floats apart (lda % 4 == 0), W (N,K) contiguous, K % 4 == 0; C rows ldc apart.
 * ------------------------------------------------------------------------------------------ */
int sigma_linear_tf32(const float *A, int64_t lda, const float *W, const float *bias,
                      const float *residual, int64_t ldr, float *C, int64_t ldc, int64_t M, int N,
                      int K, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SIGMA_B200_H_ */
