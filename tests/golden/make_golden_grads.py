"""Gradient goldens of whole blocks from the UNMODIFIED reference (build container): SS2D inside a VSSBlock, CroMB and
ConMB fusion blocks — torch autograd through the reference's own classes with its CUDA op replaced by its own
`selective_scan_ref` (differentiable), exactly the oracle the reference's op test uses.  Stores the loss, the input
gradients and the gradient of every parameter (small blocks: hidden 32, 6x5 maps).
    python tests/golden/make_golden_grads.py"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE)]
import procedural as P  # noqa: E402
import ref_shim  # noqa: E402

SEED = 7


def main():
    torch.set_num_threads(8)
    ns = ref_shim.install()
    vm = ns.vmamba
    # the shim's fwd is called inside the reference's autograd.Function (no grad): route the Function itself to the
    # differentiable reference implementation for this script
    class _Diff:
        @staticmethod
        def apply(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1):
            return ns.selective_scan_ref(u, delta, A, B, C, D, delta_bias, delta_softplus)
    vm.SelectiveScan = _Diff
    vm.selective_scan_fn_v1 = lambda u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1: \
        ns.selective_scan_ref(u, delta, A, B, C, D, delta_bias, delta_softplus)

    def run(name, mod, *inputs):
        P.fill_state_dict(mod, SEED)
        mod.train()                      # drop_path = 0 in these blocks: training mode == eval mode numerically
        xs = [t.clone().requires_grad_(True) for t in inputs]
        out = mod(*xs)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        loss = sum((o * P.randn(SEED, f"{name}/w{i}", tuple(o.shape))).sum() for i, o in enumerate(outs))
        loss.backward()
        arrays = {"loss": loss.detach().numpy()}
        for i, x in enumerate(xs):
            arrays[f"dx{i}"] = x.grad.numpy()
        for k, p in mod.named_parameters():
            if p.grad is not None:
                arrays["g/" + k] = p.grad.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print("wrote", name, len(arrays), "arrays, loss", float(loss))

    xin = P.randn(SEED, "mod/x", (2, 6, 5, 32))
    xin2 = P.randn(SEED, "mod/x2", (2, 6, 5, 32))
    run("grad_vssblock", vm.VSSBlock(hidden_dim=32, norm_layer=nn.LayerNorm, mlp_ratio=0.0, d_state=16, drop_path=0.0), xin)
    run("grad_cromb", vm.CrossMambaFusionBlock(hidden_dim=32, mlp_ratio=0.0, d_state=4, drop_path=0.0), xin, xin2)
    run("grad_conmb", vm.ConcatMambaFusionBlock(hidden_dim=32, mlp_ratio=0.0, d_state=4, drop_path=0.0), xin, xin2)
    run("grad_cvss_dec", vm.CVSSDecoderBlock(hidden_dim=32, norm_layer=nn.LayerNorm, d_state=4, mlp_ratio=4.0, drop_path=0.0), xin)


if __name__ == "__main__":
    main()
