"""Drop-in for the reference's pybind module `selective_scan_cuda_core`
(models/encoders/selective_scan/csrc/selective_scan/selective_scan.cpp:364-367): same two entry points,
same argument order, same return lists — backed by libsigma_b200 (sigma_scan_fwd / sigma_scan_bwd)."""
from sigma_b200.ops import selective_scan_cuda_core_bwd as bwd  # noqa: F401
from sigma_b200.ops import selective_scan_cuda_core_fwd as _fwd


def fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows):
    return _fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows)
