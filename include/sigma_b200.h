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
 * (selective_scan.cpp:228, fwd_kernel.cuh:181-184; the first component is the running product since the START of the
 * sequence, as the reference's prefix callback keeps it).  D and delta_bias are nullable.  fp16 / bf16 are read and
 * written natively.  Calls whose rows are 16-byte aligned, whose channel groups are multiples of 32 and d_state in
 * {4, 8, 16} (every Sigma call) run the TMA-staged kernel (csrc/scan_op_tma.cu); anything else the generic one.
 * `workspace` only holds the L-segment carries (sigma_scan_fwd_workspace_bytes); without it the scan runs unsplit.
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
 * All tensors contiguous (d_state <= 16), fp16 / bf16 natively.  `workspace` holds the forward states of the recompute sweep
 * (one every 16 positions) and the L-segment carries of both directions;
 * the reference's `x` is not needed.  du, ddelta: (batch, dim, seqlen) in `dtype`; dA (dim, dstate),
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
 *   xc    (batch, Lseq, D)          fp32, channels-last: the dwconv+SiLU output.  Lseq = H·W; for
 *                                   SEQ2 Lseq = 2·H·W = [rgb ‖ x] per image; for CROSS batch = 2·images,
 *                                   modality-major: [0, batch/2) rgb, [batch/2, batch) modal-x
 *   xdbl  (batch, Lseq, K, Cp)      fp32: x_proj output per POSITION and direction (K = 4 / 2 / 1), row =
 *                                   [B (N) | C (N) | dt_r (R) | 0-pad], Cp = sigma_ss2d_padded_cp(N, R)
 *   y     (K, batch, Lseq, D)       fp32: direction k's output stored at the POSITION it belongs
 *                                   to (so CrossMerge is a plain sum over k)
 *   dtw (Kw, D, R), dtb (Kw, D), A (Kw·D, N) [= -exp(A_logs)], Ds (Kw·D); Kw = K, or 2 (modalities) for CROSS
 * ------------------------------------------------------------------------------------------ */
#define SIGMA_DIRS_CROSS4 0 /* K=4: row-major, column-major, and both reversed (vmamba.py:86-88) */
#define SIGMA_DIRS_SEQ2 1   /* K=2: forward and reversed over a flat sequence (vmamba.py:130-131) */
#define SIGMA_DIRS_CROSS 2  /* K=1 per modality, C taken from the other modality (vmamba.py:1530,1536) */

int sigma_ss2d_padded_cp(int N, int R); /* row length of xdbl, or -1 if R > 64 */
size_t sigma_ss2d_scan_workspace_bytes(int kind, int batch, int H, int W, int D, int N);

int sigma_ss2d_scan_fwd(int kind, const float *xc, const float *xdbl, const float *dtw,
                        const float *dtb, const float *A, const float *Ds, float *y,
                        int batch, int H, int W, int D, int N, int R, int Cp,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * f1. Backward of the fused scan (training): replaces the autograd of CrossScan (vmamba.py:80-98) + the dt_proj einsum (:199) +
 * SelectiveScan (selective_scan_bwd_kernel.cuh:68-274) + CrossMerge (:100-121) without materialising the (B,4,D,L) copies.
 * kind = SIGMA_DIRS_CROSS4 or SIGMA_DIRS_SEQ2; d_state in {4, 16}; D % 64 == 0.  Inputs as the forward plus
 *   dy      (batch, Lseq, D)      gradient of the MERGED output y = sum_k y_k (CrossMerge is a sum, so every direction sees dy)
 * Outputs (fp32):
 *   dxc     (batch, Lseq, D)      sum over directions of du, accumulated by TMA reduce-add (zeroed inside)
 *   ddelta  (K, batch, Lseq, D)   gradient w.r.t. the PRE-softplus dt_proj output, per direction, at the position it belongs to; the
 *                                 caller finishes d dt_r = ddelta_k · W_dt[k] and dW_dt[k] = ddelta_k^T · dt_r_k with two GEMMs
 *   dxdbl   (batch, Lseq, K, Cp)  dB in columns [0, N), dC in [N, 2N) (zeroed inside; dt_r columns left 0 for the caller)
 *   dA (K·D, N), dDs (K·D), ddtb (K, D)   overwritten
 *   delta   (K, batch, Lseq, D)   scratch: the recomputed softplus(dt_proj) slabs
 * ------------------------------------------------------------------------------------------ */
size_t sigma_ss2d_scan_bwd_workspace_bytes(int kind, int batch, int H, int W, int D, int N);
int sigma_ss2d_scan_bwd(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A, const float *Ds,
                        const float *dy, float *delta, float *dxc, float *ddelta, float *dxdbl, float *dA, float *dDs, float *ddtb, int batch,
                        int H, int W, int D, int N, int R, int Cp, void *workspace, size_t workspace_bytes, void *stream);

/* Training forward + its backward: `sigma_ss2d_scan_fwd_save` is sigma_ss2d_scan_fwd that also writes what the backward would
 * otherwise recompute in a state sweep — delta (K, batch, Lseq, D) = softplus(dt_proj) at the position it belongs to, and
 * hs (sigma_ss2d_scan_hs_bytes) = the scan state at the start of every 16-position block of each direction's walk (the role of
 * the reference's chunk states `x`, selective_scan_fwd_kernel.cuh:176-190).  `sigma_ss2d_scan_bwd_saved` consumes them (delta
 * and hs are inputs) and runs only the reverse sweep.  nsplit = 0 lets the library choose the L-segments.  CROSS4 / SEQ2,
 * d_state 4 / 16. */
size_t sigma_ss2d_scan_hs_bytes(int kind, int batch, int H, int W, int D, int N);
int sigma_ss2d_scan_fwd_save(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A,
                             const float *Ds, float *y, float *delta, float *hs, int batch, int H, int W, int D, int N, int R, int Cp,
                             void *workspace, size_t workspace_bytes, int nsplit, void *stream);
int sigma_ss2d_scan_bwd_saved(int kind, const float *xc, const float *xdbl, const float *dtw, const float *dtb, const float *A, const float *Ds,
                              const float *dy, const float *delta, const float *hs, float *dxc, float *ddelta, float *dxdbl, float *dA,
                              float *dDs, float *ddtb, int batch, int H, int W, int D, int N, int R, int Cp, void *workspace,
                              size_t workspace_bytes, int nsplit, void *stream);

/* ------------------------------------------------------------------------------------------
 * Row-wise / stencil pieces of a5-a11 (channels-last, fp32; D % 4 == 0, 16-byte aligned rows).
 * ------------------------------------------------------------------------------------------ */
/* nn.LayerNorm over the last dim (vmamba.py:1693,724,2173; eps=1e-5): y = (x-mean)/sqrt(var+eps)·w+b */
int sigma_layernorm_fwd(const float *x, const float *w, const float *b, float *y, int64_t rows,
                        int C, float eps, void *stream);

/* Backward of sigma_layernorm_fwd (training path; the reference's autograd of nn.LayerNorm): dx (rows, C); dw (C) = sum over rows
 * of dy·xhat, db (C) = sum over rows of dy — both zeroed inside, then accumulated.  mean / rstd are recomputed from x.
 * C/4 must be one of {8, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384} (every Sigma width up to 1536); else SIGMA_EUNSUPPORTED. */
int sigma_layernorm_bwd(const float *x, const float *dy, const float *w, float *dx, float *dw, float *db, int64_t rows, int C,
                        float eps, void *stream);

/* PatchMerging2D front half (vmamba.py:619-633): y[b,i,j,:] = LayerNorm(cat(x[b,2i,2j], x[b,2i+1,2j], x[b,2i,2j+1],
 * x[b,2i+1,2j+1])) over 4C channels, zero rows beyond an odd H / W (F.pad).  x (batch,H,W,C) -> y (batch,⌈H/2⌉,⌈W/2⌉,4C);
 * the gather is index math inside the LayerNorm kernel (no concatenated tensor).  4C must be 32·k·{2,3,4,6,8,12,16}-shaped
 * (every Sigma width is).                                                                    */
int sigma_patch_merge_norm_fwd(const float *x, const float *w, const float *b, float *y, int batch, int H, int W,
                               int C, float eps, void *stream);

/* PatchExpand back half (MambaDecoder.py:24-30): x is the expand Linear's output (batch,H,W,2,2,C) ("b h w (p1 p2 c)"),
 * y[b,2h+p1,2w+p2,:] = LayerNorm(x[b,h,w,p1,p2,:]); y (batch,2H,2W,C).  The pixel shuffle is the store address.   */
int sigma_pixel_shuffle_norm_fwd(const float *x, const float *w, const float *b, float *y, int batch, int H, int W,
                                 int C, float eps, void *stream);

/* depthwise 3x3 conv (pad 1) + bias + SiLU, channels-last (vmamba.py:683-692,1072).
 * x: position rows x_row_stride floats apart, images x_batch_stride apart (so the x half of
 * in_proj's (.., 2D) output is read in place); w is the nn.Conv2d weight (D,1,3,3) contiguous;
 * y: (H·W, D) rows per image, images y_batch_stride floats apart.                         */
int sigma_dwconv3x3_silu_fwd(const float *x, int64_t x_row_stride, int64_t x_batch_stride,
                             const float *w, const float *bias, float *y, int64_t y_batch_stride,
                             int batch, int H, int W, int D, void *stream);

/* CrossMerge sum + out_norm LayerNorm + gates (vmamba.py:217-224,1077; ConMB: 423-428,1280-1281):
 *   out[r,:] = (LN(Σ_k y[k][r,:])·gamma+beta) · (z ? SiLU(z[r,:]) : 1) · (gate ? gate[r / rows_per_batch, :] : 1)
 * Row r = (b, i) with b = r / rows_per_batch: input row at y + k·k_stride + b·in_batch_stride + i·D,
 * output row at out + b·out_batch_stride + i·out_row_stride, z row at z + r·z_row_stride.     */
int sigma_merge_norm_gate_fwd(const float *y, int K, int64_t k_stride, int64_t in_batch_stride,
                              const float *gamma, const float *beta, const float *z,
                              int64_t z_row_stride, const float *gate, float *out,
                              int64_t out_batch_stride, int64_t out_row_stride, int64_t rows,
                              int64_t rows_per_batch, int D, float eps, void *stream);

/* UpsampleExpand tail (MambaDecoder.py:47-49): y = LayerNorm(bilinear x2 (align_corners=False) of x); x (batch,H,W,C),
 * y (batch,2H,2W,C).  One pass: the upsampled tensor is never materialised un-normalised.
 * w == b == NULL: plain bilinear x2 without the LayerNorm (FinalUpsample_X4's first interpolate, MambaDecoder.py:92). */
int sigma_upsample2x_norm_fwd(const float *x, const float *w, const float *b, float *y, int batch, int H, int W, int C,
                              float eps, void *stream);

/* FinalUpsample_X4 tail + classifier (MambaDecoder.py:95-96, 276-279): logits = Conv1x1(LayerNorm(bilinear x2 (x))).
 * x (batch,H,W,C); wcls (num_classes, C) = the 1x1 conv weight; logits (batch, num_classes, 2H, 2W) NCHW.   */
int sigma_upsample2x_norm_head_fwd(const float *x, const float *w, const float *b, const float *wcls, int num_classes,
                                   float *logits, int batch, int H, int W, int C, float eps, void *stream);

/* ChannelAttention pooling (vmamba.py:1738-1739): per-slice partial sums and maxima over the L positions of each image,
 * x (batch, L, C) channels-last -> partial (batch, nslice, 2, C) [0]=sum [1]=max (the caller finishes the tiny reduction). */
int sigma_pool_avgmax_partial_fwd(const float *x, float *partial, int batch, int64_t L, int C, int nslice, void *stream);

/* out[r,:] = a[r,:]·sa[r / rows_per_batch, :] + b[r,:]·sb[:]   (a, sa nullable: out = b·sb).  CVSSDecoderBlock residuals
 * (vmamba.py:1801,1803) with the channel-attention scaling (vmamba.py:1741) folded in.                      */
int sigma_scale_add_fwd(const float *a, const float *sa, const float *b, const float *sb, float *out, int64_t rows,
                        int64_t rows_per_batch, int C, void *stream);

/* ------------------------------------------------------------------------------------------
 * Dense projections (in_proj / x_proj / out_proj / PatchMerging / PatchExpand / decoder linears,
 * vmamba.py:679,725,616,195; MambaDecoder.py:17,39,82-83): hand-written tcgen05 (5th-gen tensor core) TF32 GEMM,
 * fp32 storage, fp32 accumulate in TMEM, TMA-fed, fused epilogue:
 *     C[M,N] = A[M,K]·W[N,K]^T (+ bias[N]) (+ residual[M,N] (· rscale[N]))
 * A rows lda floats apart, W (N,K) contiguous, C rows ldc apart, residual rows ldr apart; K, lda, ldc, ldr % 4 == 0.
 * rscale is the per-channel residual scale of CVSSDecoderBlock (vmamba.py:1801); bias/residual/rscale nullable.
 * ------------------------------------------------------------------------------------------ */
int sigma_linear_tf32(const float *A, int64_t lda, const float *W, const float *bias, const float *residual, int64_t ldr,
                      const float *rscale, float *C, int64_t ldc, int64_t M, int N, int K, void *stream);

/* The same GEMM with fp32-GRADE products on the TF32 tensor pipe ("tf32x3": A·W = A_hi·W_hi + A_lo·W_hi + A_hi·W_lo, x_hi = x with
 * the low 13 mantissa bits cleared; three tcgen05 MMAs per k-step, activations split in shared memory inside the kernel): what
 * torch's nn.Linear computes with torch.backends.cuda.matmul.allow_tf32 = False, to ~1e-6 relative.  The weights arrive
 * pre-split (W_hi, W_lo each (N, K) contiguous): split them once with sigma_split_tf32_fwd.                                  */
int sigma_linear_tf32x3(const float *A, int64_t lda, const float *W_hi, const float *W_lo, const float *bias, const float *residual,
                        int64_t ldr, const float *rscale, float *C, int64_t ldc, int64_t M, int N, int K, void *stream);
int sigma_split_tf32_fwd(const float *x, float *hi, float *lo, int64_t n, void *stream);

/* Dense 3x3 convolution (pad 1, stride 1) of the ChannelAttentionBlock (vmamba.py:1749-1752), channels-last, as an implicit GEMM on
 * the same tcgen05 kernel: y (batch, H, W, Cout) = conv(x (batch, H, W, Cin), w9) + bias, act = 1 applies the exact (erf) GELU of
 * nn.GELU() in the epilogue.  w9 = the nn.Conv2d weight (Cout, Cin, 3, 3) re-ordered to (3·3, Cout, Cin); every tap's input patch
 * is one shifted 4-D TMA box whose out-of-bounds fill is the zero padding.  w9_lo == NULL: one TF32 MMA per k-step; otherwise
 * tf32x3 with (w9, w9_lo) = sigma_split_tf32_fwd of the re-ordered weight.  Cin, Cout % 4 == 0.                              */
int sigma_conv3x3_tf32(const float *x, const float *w9, const float *w9_lo, const float *bias, int act, float *y, int batch, int H, int W,
                       int Cin, int Cout, void *stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY.md §8(f) rank 2, first piece: the evaluator's per-batch metric on the device (eval.py:22-29,
 * utils/metric.py:8-15).  pred = argmax over classes of logits (batch, classes, H, W) — the index numpy.argmax
 * returns for exp(score) (evaluator.py:520,449); for labels in [0, classes): hist[label·classes + pred] += 1,
 * counts[0] (labeled) += 1, counts[1] (correct) += (pred == label).  hist (classes²) and counts (2) are uint64
 * ACCUMULATORS in device memory (zero them once per evaluation); labels (batch, H, W) are uint8 / int32 / int64
 * (label_bytes = 1 / 4 / 8), anything outside [0, classes) — e.g. 255 — is ignored; pred (batch·H·W uint8) may be NULL.
 * num_classes <= 238 (the per-CTA histogram lives in shared memory).
 * ------------------------------------------------------------------------------------------ */
int sigma_argmax_hist_fwd(const float *logits, const void *labels, int label_bytes, uint64_t *hist,
                          uint64_t *counts, uint8_t *pred, int batch, int num_classes, int64_t HW,
                          void *stream);

/* ------------------------------------------------------------------------------------------
 * the callers either side of the hot path, on the device.
 *
 * sigma_image_pre_fwd — the pre-processing of ONE image into one (3, OH, OW) float32 network input (and optionally its
 * (OH, OW) int64 label map): replaces, in one kernel, TrainPre.__call__ (dataloader/dataloader.py:26-50: random_mirror,
 * random_scale = cv2.resize INTER_LINEAR / INTER_NEAREST, normalize, random_crop_pad_to_shape) and the evaluator's
 * process_image_rgbX (engine/evaluator.py:523-558: normalize + pad_image_to_shape) incl. the multi-scale cv2.resize of
 * sliding_eval_rgbX (:439-444) and the horizontal flip of the padded input (:512-515).  The random choices are made by the
 * caller and passed in.  src (H0, W0, 3) uint8 HWC; labels (H0, W0) uint8 or NULL.
 *   scaled image = cv2.resize(src [flipped horizontally first if mirror_src], (SW, SH), INTER_LINEAR), 8-bit fixed-point
 *                  arithmetic of OpenCV's generic path (scale_y / scale_x = source pixels per scaled pixel: 1/fy, 1/fx, or
 *                  H0/SH, W0/SW); SH == H0 and SW == W0: no resize
 *   out(c, oy, ox) = ((scaled(oy + off_y, ox + off_x, c) / 255) - mean[c]) / std[c]  (double, utils/transforms.py:182-187),
 *                    0 outside the scaled image — or outside the rectangle clip4_host = {y0, x0, h, w} of it when given (a sliding
 *                    window, evaluator.py:478-482) —; label_out = label_pad there (255 in TrainPre); mirror_out flips the OUTPUT.
 * ------------------------------------------------------------------------------------------ */
int sigma_image_pre_fwd(const uint8_t *src, const uint8_t *labels, float *out, int64_t *labels_out, int H0, int W0, int SH, int SW,
                        double scale_y, double scale_x, int OH, int OW, int off_y, int off_x, int mirror_src, int mirror_out,
                        int label_pad, const int *clip4_host, const double *mean3_host, const double *std3_host, void *stream);

/* evaluator.py:505-520 and 481-488: acc[c, ay + y, ax + x] += exp(logits[c, m_top + y, m_left + x] (+ logits_flip[c, m_top + y,
 * TW - 1 - (m_left + x)])) for y < vh, x < vw.  logits (ncls, TH, TW) = one image of the model's NCHW output; logits_flip
 * (nullable) = the output for the horizontally flipped input; acc (ncls, AH, AW) float32 = the scale's score map.         */
int sigma_eval_exp_accumulate_fwd(const float *logits, const float *logits_flip, float *acc, int ncls, int TH, int TW, int m_top,
                                  int m_left, int vh, int vw, int AH, int AW, int ay, int ax, void *stream);

/* evaluator.py:497-499 and 447-448: out (H0, W0, ncls) float64 += cv2.resize(acc[:, m_top : m_top + SH, m_left : m_left + SW]
 * as HWC float32, (W0, H0), INTER_LINEAR) (cv2's float path; identity when SH == H0 and SW == W0).                        */
int sigma_eval_resize_add_fwd(const float *acc, int ncls, int AH, int AW, int m_top, int m_left, int SH, int SW, double *out,
                              int H0, int W0, void *stream);

/* evaluator.py:451 + eval.py:28 + utils/metric.py:8-15: pred = argmax over classes of score (HW, ncls) float64 (first maximum,
 * numpy.argmax); labels (HW) uint8 nullable: hist[label·ncls + pred] += 1, counts[0] (labeled) += 1, counts[1] (correct)
 * += (pred == label) for label < ncls.  hist / counts are uint64 accumulators.                                            */
int sigma_eval_argmax_hist_fwd(const double *score, const uint8_t *labels, uint8_t *pred, uint64_t *hist, uint64_t *counts,
                               int num_classes, int64_t HW, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SIGMA_B200_H_ */
