"""Generate the committed golden fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py        # needs /root/reference; writes tests/golden/*.npz

Every input and every weight is procedural (tests/procedural.py), so only OUTPUTS of the reference
are stored.  The reference op on CPU is its own `selective_scan_ref`
(selective_scan_interface.py:86-131) — the same oracle its test-suite uses
(test_selective_scan.py:181) — and the modules are the reference's own classes, imported through
tests/golden/ref_shim.py.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE)]
import procedural as P  # noqa: E402
import ref_shim  # noqa: E402

SEED = 7


def save(name, **arrays):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


def bwd_goldens(ns):
    """Gradients of the reference's selective_scan_ref by torch autograd — exactly what the
    reference's own test compares its CUDA bwd against (test_selective_scan.py:181-224)."""
    for idx, (b, d, n, L, g, has_D, has_b, sp) in enumerate([
        (2, 24, 8, 200, 2, True, True, True), (1, 16, 16, 333, 1, True, True, True),
        (2, 12, 4, 97, 3, False, False, False), (1, 8, 4, 2500, 1, True, True, True),
    ]):
        u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED + 1, b, d, n, L, g, has_D=has_D, has_bias=has_b)
        dout = P.randn(SEED + 1, f"bwd/dout{idx}", (b, d, L))
        leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (u, dl, A, Bm, Cm, D, bias)]
        with torch.enable_grad():
            out = ns.selective_scan_ref(*leaves, sp)
            out.backward(dout)
        names = ["du", "ddelta", "dA", "dB", "dC", "dD", "dbias"]
        save(f"scan_bwd_case{idx}", cfg=np.array((b, d, n, L, g, int(has_D), int(has_b), int(sp))),
             out=out.detach(), **{nm: t.grad for nm, t in zip(names, leaves) if t is not None})


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    ns = ref_shim.install()
    bwd_goldens(ns)
    if "--only-bwd" in sys.argv:
        return
    vm, dv, md, bd = ns.vmamba, ns.dual_vmamba, ns.mamba_decoder, ns.builder

    # ---------------- op level: selective scan ----------------
    cases = []
    for (b, d, n, L, g, has_D, has_b, sp) in [
        (2, 24, 8, 64, 1, True, True, True), (2, 24, 8, 372, 2, True, True, True),
        (2, 24, 8, 1134, 1, False, True, True), (2, 24, 8, 1134, 2, True, False, False),
        (2, 24, 8, 2048, 1, True, True, False), (1, 12, 4, 4100, 3, True, True, True),
        (3, 8, 16, 37, 1, True, True, True), (2, 32, 1, 300, 4, False, False, False),
    ]:
        u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED, b, d, n, L, g, has_D=has_D, has_bias=has_b)
        out = ns.selective_scan_ref(u, dl, A, Bm, Cm, D, bias, sp)
        cases.append((b, d, n, L, g, int(has_D), int(has_b), int(sp)))
        save(f"scan_case{len(cases) - 1}", out=out, cfg=np.array(cases[-1]))
    # BASELINE config 1: u,delta (2,768,1024), A (768,16), B,C (2,4,16,1024); keep every 16th channel
    u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED, 2, 768, 16, 1024, 4)
    out = ns.selective_scan_ref(u, dl, A, Bm, Cm, D, bias, True)
    save("scan_config1", out_sub=out[:, ::16].contiguous(), cfg=np.array((2, 768, 16, 1024, 4, 1, 1, 1)))

    # ---------------- direction maps ----------------
    x = P.randn(SEED, "cs/x", (2, 5, 6, 7))
    xs = vm.CrossScan.apply(x)
    ys = P.randn(SEED, "cs/ys", (2, 4, 5, 6, 7))
    y = vm.CrossMerge.apply(ys)
    xr, xe = P.randn(SEED, "csm/r", (2, 5, 3, 4)), P.randn(SEED, "csm/e", (2, 5, 3, 4))
    xf = vm.CrossScan_multimodal.apply(xr, xe)
    ysm = P.randn(SEED, "csm/ys", (2, 2, 5, 24))
    y1, y2 = vm.CrossMerge_multimodal.apply(ysm)
    save("cross_scan", xs=xs, y=y, xf=xf, y1=y1.contiguous(), y2=y2.contiguous())

    # ---------------- module level ----------------
    def run(name, mod, *inputs):
        P.fill_state_dict(mod, SEED)
        mod.eval()
        out = mod(*inputs)
        out = out if isinstance(out, (tuple, list)) else (out,)
        save(name, **{f"out{i}": o for i, o in enumerate(out)})

    xin = P.randn(SEED, "mod/x", (2, 6, 5, 32))
    xin2 = P.randn(SEED, "mod/x2", (2, 6, 5, 32))
    run("ss2d_n16", vm.SS2D(d_model=32, d_state=16), xin)
    run("ss2d_n4", vm.SS2D(d_model=32, d_state=4), xin)
    run("vssblock", vm.VSSBlock(hidden_dim=32, norm_layer=nn.LayerNorm, mlp_ratio=0.0, d_state=16), xin)
    run("patchmerge_odd", vm.PatchMerging2D(32, 64), P.randn(SEED, "mod/pm", (2, 5, 7, 32)))
    run("cromb", vm.CrossMambaFusionBlock(hidden_dim=32, mlp_ratio=0.0, d_state=4), xin, xin2)
    run("conmb", vm.ConcatMambaFusionBlock(hidden_dim=32, mlp_ratio=0.0, d_state=4), xin, xin2)
    run("cvss_dec", vm.CVSSDecoderBlock(hidden_dim=32, norm_layer=nn.LayerNorm, d_state=4, mlp_ratio=4.0), xin)

    feats = [P.randn(SEED, f"dec/f{i}", (1, 32 * 2 ** i, 16 // 2 ** i, 24 // 2 ** i)) for i in range(4)]
    run("mamba_decoder", md.MambaDecoder(img_size=[64, 96], in_channels=[32, 64, 128, 256], num_classes=5,
                                         embed_dim=32), feats)
    enc = dv.RGBXTransformer(depths=[1, 1, 2, 1], dims=32, pretrained=None, mlp_ratio=0.0,
                             downsample_version="v1", drop_path_rate=0.2)
    P.fill_state_dict(enc, SEED)
    enc.eval()
    outs = enc(P.randn(SEED, "enc/rgb", (1, 3, 64, 96)), P.randn(SEED, "enc/x", (1, 3, 64, 96)))
    save("rgbx_encoder_small", **{f"out{i}": o for i, o in enumerate(outs)})

    # ---------------- full model: Sigma-tiny ----------------
    sys.path.insert(0, ref_shim.REF_ROOT)
    from utils.metric import hist_info, compute_score
    for tag, (H, W, Bn) in {"sigma_tiny_64x96": (64, 96, 2), "sigma_tiny_72x104_odd": (72, 104, 1)}.items():
        cfg = types.SimpleNamespace(backbone="sigma_tiny", decoder="MambaDecoder", num_classes=9,
                                    image_height=H, image_width=W, pretrained_model=None,
                                    bn_eps=1e-3, bn_momentum=0.1)
        model = bd.EncoderDecoder(cfg, criterion=None, norm_layer=nn.BatchNorm2d)
        P.fill_state_dict(model, SEED)
        model.eval()
        ns.stub.calls.clear()
        rgb = P.randn(SEED, tag + "/rgb", (Bn, 3, H, W))
        mx = P.randn(SEED, tag + "/x", (Bn, 3, H, W))
        logits = model(rgb, mx)
        pred = logits.argmax(1).numpy()
        gt = (P.rand(SEED, tag + "/gt", (Bn, H, W)) * 9).long().clamp(max=8).numpy()
        hist, labeled, correct = hist_info(9, pred, gt)
        iou, mean_iou, _, freq_iou, mean_acc, pix_acc = compute_score(hist, correct, labeled)
        save(tag, logits=logits, miou=np.float64(mean_iou), iou=np.asarray(iou), ncalls=len(ns.stub.calls))
        print(tag, "mIoU", mean_iou, "scan calls", len(ns.stub.calls))


if __name__ == "__main__":
    main()
