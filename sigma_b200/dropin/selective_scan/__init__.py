"""Drop-in for the reference package `selective_scan`
(models/encoders/selective_scan/selective_scan/__init__.py:8): SelectiveScanFn, selective_scan_fn, selective_scan_ref."""
import torch
import torch.nn.functional as F

from sigma_b200.ops import SelectiveScanFn, selective_scan_fn  # noqa: F401


def selective_scan_ref(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False):
    """The reference's pure-torch definition of the op (selective_scan_interface.py:86-131), kept for API parity:
    sequential recurrence in fp32, any device.  Not used by the product path."""
    dtype_in = u.dtype
    u, delta = u.float(), delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:
        delta = F.softplus(delta)
    b, d, L = u.shape
    B, C = B.float(), C.float()
    if B.dim() == 3:
        B = B[:, None]
    if C.dim() == 3:
        C = C[:, None]
    Bx = B.repeat_interleave(d // B.shape[1], dim=1)
    Cx = C.repeat_interleave(d // C.shape[1], dim=1)
    h = u.new_zeros((b, d, A.shape[1]))
    ys = []
    for l in range(L):
        h = torch.exp(delta[:, :, l, None] * A) * h + (delta[:, :, l] * u[:, :, l])[..., None] * Bx[..., l]
        ys.append((h * Cx[..., l]).sum(-1))
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u * D[None, :, None]
    return out.to(dtype_in)
