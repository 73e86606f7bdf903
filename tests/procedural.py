"""Deterministic, RNG-library-independent test tensors.

Every tensor is a pure function of (seed, name, shape): numpy PCG64 seeded by
[seed, crc32(name)].  tests/golden/make_golden.py uses these to fill the UNMODIFIED reference
model's state_dict and inputs; the parity tests regenerate the very same values for our modules
(the state_dict keys are identical by construction), so no weights are committed.
"""
import math
import zlib

import numpy as np
import torch


def _rng(seed: int, name: str):
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def randn(seed, name, shape, scale=1.0, shift=0.0):
    a = _rng(seed, name).standard_normal(tuple(shape), dtype=np.float32) * scale + shift
    return torch.from_numpy(a.astype(np.float32))


def rand(seed, name, shape, lo=0.0, hi=1.0):
    a = _rng(seed, name).random(tuple(shape), dtype=np.float32) * (hi - lo) + lo
    return torch.from_numpy(a.astype(np.float32))


def fill_param(seed: int, key: str, shape):
    """Value for one state_dict entry, chosen by the entry's role (inferred from its name)."""
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf in ("A_logs", "A_log_1", "A_log_2"):
        n = shape[-1]  # break the S4D structure A=-(n+1) on purpose
        return torch.log(rand(seed, key, shape, 0.5, n + 0.5))
    if leaf in ("Ds", "D_1", "D_2"):
        return randn(seed, key, shape, 0.2, 1.0)
    if leaf in ("scale1", "scale2"):
        return randn(seed, key, shape, 0.1, 1.0)
    if leaf == "dt_projs_bias" or (leaf == "bias" and ".dt_proj_" in key):
        dt = torch.exp(rand(seed, key, shape) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
        dt = dt.clamp(min=1e-4)
        return dt + torch.log(-torch.expm1(-dt))
    if leaf == "dt_projs_weight" or (leaf == "weight" and ".dt_proj_" in key):
        r = shape[-1]
        return rand(seed, key, shape, -r ** -0.5, r ** -0.5)
    if len(shape) == 1:
        if leaf == "weight":          # LayerNorm gains
            return randn(seed, key, shape, 0.1, 1.0)
        return randn(seed, key, shape, 0.1, 0.0)  # all biases
    if len(shape) == 3:               # x_proj_weight (K, R+2N, D)
        fan_in = shape[-1]
    else:
        fan_in = int(np.prod(shape[1:]))
    return randn(seed, key, shape, fan_in ** -0.5)


@torch.no_grad()
def fill_state_dict(module: torch.nn.Module, seed: int = 7):
    """Overwrite every parameter/buffer of `module` in place, keyed by its state_dict name."""
    sd = module.state_dict()
    for k, v in sd.items():
        if not torch.is_floating_point(v):
            continue
        v.copy_(fill_param(seed, k, v.shape).to(v.dtype))
    return module


def scan_inputs(seed, batch, dim, dstate, seqlen, ngroups=1, dtype=torch.float32,
                has_D=True, has_bias=True):
    """Op-level inputs with the distributions of the reference's own test
    (models/encoders/selective_scan/test_selective_scan.py:153-179)."""
    tag = f"scan/{batch}/{dim}/{dstate}/{seqlen}/{ngroups}"
    A = -0.5 * rand(seed, tag + "/A", (dim, dstate))
    Bm = randn(seed, tag + "/B", (batch, ngroups, dstate, seqlen)).to(dtype)
    Cm = randn(seed, tag + "/C", (batch, ngroups, dstate, seqlen)).to(dtype)
    D = randn(seed, tag + "/D", (dim,)) if has_D else None
    bias = 0.5 * rand(seed, tag + "/bias", (dim,)) if has_bias else None
    u = randn(seed, tag + "/u", (batch, dim, seqlen)).to(dtype)
    delta = (0.5 * rand(seed, tag + "/delta", (batch, dim, seqlen))).to(dtype)
    return u, delta, A, Bm, Cm, D, bias
