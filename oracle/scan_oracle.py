"""ctypes binding of the C oracle (oracle/selective_scan_ref.c).  ORACLE — test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsigma_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "selective_scan_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _threads(nthreads):
    """0 = "the caller did not say": honour OMP_NUM_THREADS as it is NOW (the OpenMP runtime read it once when it was
    loaded — possibly before bench.py or torchrun changed it), else leave the runtime's default."""
    if nthreads:
        return int(nthreads)
    try:
        return max(0, int(os.environ.get("OMP_NUM_THREADS", "0")))
    except ValueError:
        return 0


def scan_fwd(u, delta, A, B, C, D=None, bias=None, softplus=False, nthreads=0):
    """numpy in / numpy out.  u, delta (b,d,L); A (d,N); B, C (b,G,N,L) or (b,N,L)."""
    u, delta, A, B, C, D, bias = map(_f, (u, delta, A, B, C, D, bias))
    if B.ndim == 3:
        B, C = B[:, None], C[:, None]
    b, d, L = u.shape
    G, N = B.shape[1], B.shape[2]
    out = np.empty_like(u)
    lib().sigma_oracle_scan_fwd(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(bias), _p(out),
                                b, d, L, N, G, int(bool(softplus)), _threads(nthreads))
    return out


def scan_bwd(u, delta, A, B, C, D, bias, dout, softplus=False, nthreads=0):
    u, delta, A, B, C, D, bias, dout = map(_f, (u, delta, A, B, C, D, bias, dout))
    b, d, L = u.shape
    G, N = B.shape[1], B.shape[2]
    du, ddelta = np.empty_like(u), np.empty_like(u)
    dA = np.empty((d, N), np.float32)
    dB, dC = np.empty_like(B), np.empty_like(C)
    dD, dbias = np.empty(d, np.float32), np.empty(d, np.float32)
    lib().sigma_oracle_scan_bwd(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(bias), _p(dout), _p(du),
                                _p(ddelta), _p(dA), _p(dB), _p(dC), _p(dD), _p(dbias), b, d, L, N, G,
                                int(bool(softplus)), _threads(nthreads))
    return du, ddelta, dA, dB, dC, (dD if D is not None else None), (dbias if bias is not None else None)
