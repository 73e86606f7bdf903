"""sigma_b200 — B200-native (sm_100a) SS2D / selective-scan hot path of Sigma (zifuwan/Sigma).

Package layout: `csrc/` hand-written CUDA + the C-ABI (include/sigma_b200.h), `_lib.py` ctypes
binding, `ops.py` the reference's op surface (selective_scan_cuda_core.fwd/bwd, SelectiveScan,
CrossScan...), `modules.py` the reference's nn.Module surface (SS2D, VSSBlock, ConMB, CroMB,
MambaDecoder, EncoderDecoder) on top of the fused kernels, `dropin/` import-path shims so the
reference's train.py / eval.py run unchanged.
"""
__version__ = "0.1.0"
