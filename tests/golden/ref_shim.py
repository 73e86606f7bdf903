"""Import the UNMODIFIED reference (zifuwan/Sigma @ /root/reference) on a CPU-only box.

Used ONLY by tests/golden/make_golden.py (run in the build container, where /root/reference
exists) to generate the committed fixtures.  Nothing on the GPU box imports this.

The reference needs timm / fvcore / easydict and its own CUDA extension
`selective_scan_cuda_core`; none are installed, so they are replaced by minimal stand-ins and
the extension's `fwd` is routed to the reference's own pure-torch `selective_scan_ref`
(models/encoders/selective_scan/selective_scan/selective_scan_interface.py:86-131).
"""
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = "/root/reference"


def _identity_droppath():
    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, *a, **k):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            return x
    return DropPath


def install():
    """Install stand-in modules and return the imported reference namespace."""
    if "selective_scan_cuda_core" in sys.modules and getattr(
            sys.modules["selective_scan_cuda_core"], "_sigma_ref_shim", False):
        return _namespace()

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.DropPath = _identity_droppath()
    timm_layers.trunc_normal_ = nn.init.trunc_normal_
    timm_layers.to_2tuple = to_2tuple
    timm.models = timm_models
    timm_models.layers = timm_layers
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})

    fv = types.ModuleType("fvcore")
    fvnn = types.ModuleType("fvcore.nn")
    for n in ("FlopCountAnalysis", "flop_count_str", "flop_count", "parameter_count"):
        setattr(fvnn, n, None)
    fv.nn = fvnn
    sys.modules.update({"fvcore": fv, "fvcore.nn": fvnn})

    class EasyDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    ed = types.ModuleType("easydict")
    ed.EasyDict = EasyDict
    sys.modules["easydict"] = ed

    stub = types.ModuleType("selective_scan_cuda_core")
    stub._sigma_ref_shim = True
    sys.modules["selective_scan_cuda_core"] = stub

    sys.path[:0] = [REF_ROOT + "/models/encoders/selective_scan", REF_ROOT]
    from selective_scan import selective_scan_ref  # the reference's own oracle

    calls = []

    def fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows):
        calls.append(dict(u=tuple(u.shape), N=A.shape[1], G=B.shape[1]))
        out = selective_scan_ref(u, delta, A, B, C, D, delta_bias, delta_softplus)
        return out, torch.zeros(1)

    stub.fwd = fwd
    stub.calls = calls
    return _namespace()


def _namespace():
    from selective_scan import selective_scan_ref, selective_scan_fn
    import models.encoders.vmamba as vmamba
    import models.encoders.dual_vmamba as dual_vmamba
    import models.decoders.MambaDecoder as mamba_decoder
    import models.builder as builder
    ns = types.SimpleNamespace(
        selective_scan_ref=selective_scan_ref, selective_scan_fn=selective_scan_fn,
        vmamba=vmamba, dual_vmamba=dual_vmamba, mamba_decoder=mamba_decoder, builder=builder,
        stub=sys.modules["selective_scan_cuda_core"])
    return ns
