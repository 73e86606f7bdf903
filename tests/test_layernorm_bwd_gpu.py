"""GPU: sigma_layernorm_bwd (ops.LayerNormFn) against torch's autograd of F.layer_norm — dx, dweight, dbias — over every width with
an instantiation, ragged row counts (partial warp steps), non-contiguous inputs and bf16 autocast.  Bar: 1e-5 of each tensor's scale
(fp32 sums in a different order), output bit-level 1e-6."""
import pytest
import torch
import torch.nn.functional as F

import procedural as P

pytestmark = pytest.mark.gpu
S = 77


def _cmp(name, got, ref, bar):
    sc = float(ref.abs().max()) + 1e-20
    err = float((got.float() - ref.float()).abs().max()) / sc
    assert err <= bar, f"{name}: {err:.2e} of its scale"


@pytest.mark.parametrize("C", [32, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536])
@pytest.mark.parametrize("rows", [1, 7, 300, 4801])
def test_layernorm_bwd_matches_torch(C, rows):
    from sigma_b200 import ops
    ln = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        ln.weight.copy_(P.rand(S, f"ln/w/{C}", (C,), 0.5, 1.5))
        ln.bias.copy_(P.randn(S, f"ln/b/{C}", (C,), 0.1))
    x0 = (P.randn(S, f"ln/x/{C}/{rows}", (rows, C)) * 2 + 0.3).cuda()
    wgt = P.randn(S, f"ln/g/{C}/{rows}", (rows, C)).cuda()
    xr = x0.clone().requires_grad_(True)
    (F.layer_norm(xr, (C,), ln.weight, ln.bias, ln.eps) * wgt).sum().backward()
    ref = [xr.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone()]
    ln.weight.grad = ln.bias.grad = None
    xf = x0.clone().requires_grad_(True)
    y = ops.layer_norm(ln, xf)
    assert y.grad_fn is not None and "LayerNormFn" in type(y.grad_fn).__name__
    (y * wgt).sum().backward()
    _cmp("y", y.detach(), F.layer_norm(x0, (C,), ln.weight, ln.bias, ln.eps), 2e-6)
    for nm, g, r in zip(["dx", "dweight", "dbias"], [xf.grad, ln.weight.grad, ln.bias.grad], ref):
        _cmp(f"C={C} rows={rows} {nm}", g, r, 2e-5)


def test_layernorm_fn_noncontiguous_and_autocast():
    from sigma_b200 import ops
    C = 192
    ln = torch.nn.LayerNorm(C).cuda()
    x0 = P.randn(S, "ln/nc", (2, C, 50)).cuda()
    wgt = P.randn(S, "ln/nc/w", (2, 50, C)).cuda()      # (a plain sum of squares of a LayerNorm output has zero input gradient)
    for amp in (False, True):
        ln.weight.grad = ln.bias.grad = None
        xr = x0.clone().requires_grad_(True)
        xf = x0.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            yr = F.layer_norm(xr.transpose(1, 2), (C,), ln.weight, ln.bias, ln.eps)
            yf = ops.layer_norm(ln, xf.transpose(1, 2))
        assert yf.dtype == yr.dtype
        (yr * wgt).sum().backward()
        gw = ln.weight.grad.clone()
        ln.weight.grad = ln.bias.grad = None
        (yf * wgt).sum().backward()
        _cmp("y", yf.detach(), yr.detach(), 2e-6)
        _cmp("dx", xf.grad, xr.grad, 2e-5)
        _cmp("dw", ln.weight.grad, gw, 2e-5)


def test_unsupported_width_falls_back_to_the_module():
    from sigma_b200 import ops
    ln = torch.nn.LayerNorm(40).cuda()
    x = torch.randn(5, 40, device="cuda", requires_grad=True)
    y = ops.layer_norm(ln, x)
    assert "LayerNormFn" not in type(y.grad_fn).__name__
