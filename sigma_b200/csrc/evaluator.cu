// SURVEY.md §8(f) rank 2 — the evaluator's per-image metric on the device (eval.py:22-29, utils/metric.py:8-15):
//   pred = argmax_c logits[b, c, h, w]   (the reference takes argmax of exp(score): same index; first maximum on ties,
//                                          as numpy.argmax)
//   k = label in [0, ncls);  hist[label·ncls + pred] += 1;  labeled += 1;  correct += (pred == label)
// so only ncls² + 2 integers leave the GPU instead of the logits.  Integer work, bit-exact against the reference's numpy.
// One thread per pixel (coalesced over w for every class plane), shared-memory histogram per CTA, 64-bit global atomics.
#include <algorithm>

#include "common.cuh"

namespace sigma {

template <typename LabelT>
__global__ void __launch_bounds__(256) argmax_hist_kernel(const float *__restrict__ logits, const LabelT *__restrict__ labels,
                                                         unsigned long long *__restrict__ hist, unsigned long long *__restrict__ counts,
                                                         unsigned char *__restrict__ pred_out, int batch, int ncls, long long HW) {
  extern __shared__ unsigned int sh[];   // ncls*ncls bins + labeled + correct
  const int nb = ncls * ncls;
  for (int i = threadIdx.x; i < nb + 2; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const long long total = (long long)batch * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const long long b = p / HW, r = p - b * HW;
    const float *lp = logits + b * ncls * HW + r;
    float best = lp[0];
    int arg = 0;
    for (int c = 1; c < ncls; ++c) {
      const float v = __ldg(lp + (long long)c * HW);
      if (v > best) { best = v; arg = c; }          // strict: the first maximum wins (numpy.argmax)
    }
    if (pred_out) pred_out[p] = (unsigned char)arg;
    const long long g = (long long)labels[p];
    if (g >= 0 && g < ncls) {
      atomicAdd(&sh[(int)g * ncls + arg], 1u);
      atomicAdd(&sh[nb], 1u);
      if (arg == (int)g) atomicAdd(&sh[nb + 1], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
  if (threadIdx.x == 0) {
    if (sh[nb]) atomicAdd(&counts[0], (unsigned long long)sh[nb]);
    if (sh[nb + 1]) atomicAdd(&counts[1], (unsigned long long)sh[nb + 1]);
  }
}

int argmax_hist_launch(const float *logits, const void *labels, int label_bytes, unsigned long long *hist,
                       unsigned long long *counts, unsigned char *pred_out, int batch, int ncls, long long HW, cudaStream_t stream) {
  const long long total = (long long)batch * HW;
  if (total == 0) return SIGMA_OK;
  // a CTA's shared counters are 32-bit: bound the pixels per CTA below 2^32 (grid-stride over <= 148*8 CTAs)
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 8);
  const size_t smem = (size_t)(ncls * ncls + 2) * sizeof(unsigned int);
  if (label_bytes == 1)
    argmax_hist_kernel<unsigned char><<<grid, 256, smem, stream>>>(logits, (const unsigned char *)labels, hist, counts, pred_out, batch, ncls, HW);
  else if (label_bytes == 4)
    argmax_hist_kernel<int><<<grid, 256, smem, stream>>>(logits, (const int *)labels, hist, counts, pred_out, batch, ncls, HW);
  else if (label_bytes == 8)
    argmax_hist_kernel<long long><<<grid, 256, smem, stream>>>(logits, (const long long *)labels, hist, counts, pred_out, batch, ncls, HW);
  else { set_error("sigma_argmax_hist_fwd: label_bytes=%d unsupported (1, 4, 8)", label_bytes); return SIGMA_EINVAL; }
  SIGMA_CHECK_LAUNCH();
  return SIGMA_OK;
}

}  // namespace sigma
