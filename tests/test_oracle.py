"""CPU: pin the oracle (oracle/) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Module-level cases take their state_dict KEYS AND SHAPES from
sigma_b200.modules, so they also prove the state_dict contract of SURVEY.md §8b: if a key or shape
differed from the reference's, the procedural fill would differ and the goldens would not match."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import procedural as P
from helpers import GOLDEN, SEED, assert_close, cfg_tiny, golden
from oracle import scan_oracle, sigma_ref

# fp32 torch reference vs double-accumulating oracle: agreement is at fp32 rounding level
RTOL, ATOL = 2e-5, 2e-4


def _np(t):
    return None if t is None else t.numpy()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "scan_case*.npz"))) +
                         [os.path.join(GOLDEN, "scan_config1.npz")], ids=os.path.basename)
def test_scan_fwd_oracle_matches_reference(path):
    g = np.load(path)
    b, d, n, L, G, hD, hb, sp = (int(v) for v in g["cfg"])
    u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED, b, d, n, L, G, has_D=bool(hD), has_bias=bool(hb))
    out = scan_oracle.scan_fwd(_np(u), _np(dl), _np(A), _np(Bm), _np(Cm), _np(D), _np(bias), sp)
    if "out_sub" in g.files:
        assert_close(out[:, ::16], g["out_sub"], RTOL, ATOL, "config1")
    else:
        assert_close(out, g["out"], RTOL, ATOL, os.path.basename(path))


@pytest.mark.parametrize("idx", range(4))
def test_scan_bwd_oracle_matches_reference_autograd(idx):
    g = golden(f"scan_bwd_case{idx}")
    b, d, n, L, G, hD, hb, sp = (int(v) for v in g["cfg"])
    u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED + 1, b, d, n, L, G, has_D=bool(hD), has_bias=bool(hb))
    dout = P.randn(SEED + 1, f"bwd/dout{idx}", (b, d, L))
    res = scan_oracle.scan_bwd(_np(u), _np(dl), _np(A), _np(Bm), _np(Cm), _np(D), _np(bias), _np(dout), sp)
    for name, r in zip(["du", "ddelta", "dA", "dB", "dC", "dD", "dbias"], res):
        if r is None:
            assert name not in g.files
            continue
        ref = g[name]
        assert_close(r, ref, 2e-5, 2e-6 * float(np.abs(ref).max()) + 1e-5, f"case{idx}/{name}")


def test_direction_maps():
    g = golden("cross_scan")
    x = P.randn(SEED, "cs/x", (2, 5, 6, 7))
    assert_close(sigma_ref.cross_scan(x), g["xs"], 0, 0, "CrossScan")
    ys = P.randn(SEED, "cs/ys", (2, 4, 5, 6, 7))
    assert_close(sigma_ref.cross_merge(ys.view(2, 4, 5, 42), 6, 7), g["y"], 1e-6, 1e-6, "CrossMerge")


def _filled(mod):
    P.fill_state_dict(mod, SEED)
    return {k: v.clone() for k, v in mod.state_dict().items()}


def test_module_goldens():
    from sigma_b200 import modules as M
    xin = P.randn(SEED, "mod/x", (2, 6, 5, 32))
    xin2 = P.randn(SEED, "mod/x2", (2, 6, 5, 32))
    with torch.no_grad():
        sd = _filled(M.SS2D(d_model=32, d_state=16))
        assert_close(sigma_ref.ss2d(xin, {"op." + k: v for k, v in sd.items()}, "op"), golden("ss2d_n16")["out0"], RTOL, ATOL, "ss2d_n16")
        sd = _filled(M.SS2D(d_model=32, d_state=4))
        assert_close(sigma_ref.ss2d(xin, {"op." + k: v for k, v in sd.items()}, "op"), golden("ss2d_n4")["out0"], RTOL, ATOL, "ss2d_n4")
        sd = _filled(M.VSSBlock(hidden_dim=32, norm_layer=nn.LayerNorm, mlp_ratio=0.0, d_state=16))
        assert_close(sigma_ref.vss_block(xin, {"b." + k: v for k, v in sd.items()}, "b"), golden("vssblock")["out0"], RTOL, ATOL, "vssblock")
        sd = _filled(M.PatchMerging2D(32, 64))
        assert_close(sigma_ref.patch_merging(P.randn(SEED, "mod/pm", (2, 5, 7, 32)), {"d." + k: v for k, v in sd.items()}, "d"),
                     golden("patchmerge_odd")["out0"], RTOL, ATOL, "patchmerge")
        sd = _filled(M.CrossMambaFusionBlock(hidden_dim=32, mlp_ratio=0.0, d_state=4))
        o = sigma_ref.cromb(xin, xin2, {"c." + k: v for k, v in sd.items()}, "c")
        assert_close(o[0], golden("cromb")["out0"], RTOL, ATOL, "cromb rgb")
        assert_close(o[1], golden("cromb")["out1"], RTOL, ATOL, "cromb x")
        sd = _filled(M.ConcatMambaFusionBlock(hidden_dim=32, mlp_ratio=0.0, d_state=4))
        assert_close(sigma_ref.conmb(xin, xin2, {"c." + k: v for k, v in sd.items()}, "c"), golden("conmb")["out0"], RTOL, ATOL, "conmb")
        sd = _filled(M.CVSSDecoderBlock(hidden_dim=32, norm_layer=nn.LayerNorm, d_state=4, mlp_ratio=4.0))
        assert_close(sigma_ref.cvss_decoder_block(xin, {"d." + k: v for k, v in sd.items()}, "d"), golden("cvss_dec")["out0"], RTOL, ATOL, "cvss_dec")


def test_decoder_and_encoder_goldens():
    from sigma_b200 import modules as M
    with torch.no_grad():
        dec = M.MambaDecoder(img_size=[64, 96], in_channels=[32, 64, 128, 256], num_classes=5, embed_dim=32)
        sd = {"decode_head." + k: v for k, v in _filled(dec).items()}
        feats = [P.randn(SEED, f"dec/f{i}", (1, 32 * 2 ** i, 16 // 2 ** i, 24 // 2 ** i)) for i in range(4)]
        assert_close(sigma_ref.mamba_decoder(feats, sd), golden("mamba_decoder")["out0"], RTOL, ATOL, "mamba_decoder")
        enc = M.RGBXTransformer(depths=[1, 1, 2, 1], dims=32, pretrained=None, mlp_ratio=0.0, downsample_version="v1", drop_path_rate=0.2)
        sd = {"backbone." + k: v for k, v in _filled(enc).items()}
        outs = sigma_ref.rgbx_encoder(P.randn(SEED, "enc/rgb", (1, 3, 64, 96)), P.randn(SEED, "enc/x", (1, 3, 64, 96)), sd)
        g = golden("rgbx_encoder_small")
        for i in range(4):
            assert_close(outs[i], g[f"out{i}"], 5e-5, 5e-4, f"encoder out{i}")


@pytest.mark.parametrize("tag,H,W,Bn", [("sigma_tiny_64x96", 64, 96, 2), ("sigma_tiny_72x104_odd", 72, 104, 1)])
def test_full_model_golden(tag, H, W, Bn):
    """Sigma-tiny logits + mIoU (utils/metric.py) on a fixed synthetic batch."""
    from sigma_b200 import modules as M
    g = golden(tag)
    with torch.no_grad():
        model = M.EncoderDecoder(cfg_tiny(H, W), criterion=None)
        assert len(model.state_dict()) == 668
        sd = _filled(model)
        rgb = P.randn(SEED, tag + "/rgb", (Bn, 3, H, W))
        mx = P.randn(SEED, tag + "/x", (Bn, 3, H, W))
        logits = sigma_ref.encoder_decoder(rgb, mx, sd)
    assert_close(logits, g["logits"], 1e-4, 1e-3, tag)
    gt = (P.rand(SEED, tag + "/gt", (Bn, H, W)) * 9).long().clamp(max=8).numpy()
    iou, miou = sigma_ref.mean_iou(logits.argmax(1).numpy(), gt, 9)
    agree = float((logits.argmax(1).numpy() == g["logits"].argmax(1)).mean())
    assert agree > 0.9995, agree
    assert abs(miou - float(g["miou"])) < 2e-4


def test_oracle_port_matches_reference_at_the_benchmarked_size():
    """Sigma-tiny 480x640 B=1 (BASELINE config 2): the oracle port against the fixture produced by the unmodified
    reference at that size (tests/golden/make_golden_fullsize.py).  ~20-40 s on 8 cores."""
    import contextlib
    import io
    from sigma_b200 import modules as M
    tag, H, W, ncls = "sigma_tiny_480x640", 480, 640, 9
    g = golden(tag)
    with contextlib.redirect_stdout(io.StringIO()):
        model = M.EncoderDecoder(cfg_tiny(H, W, num_classes=ncls), criterion=None)
    P.fill_state_dict(model, SEED)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    rgb = P.randn(SEED, tag + "/rgb", (1, 3, H, W))
    mx = P.randn(SEED, tag + "/x", (1, 3, H, W))
    with torch.no_grad():
        logits = sigma_ref.encoder_decoder(rgb, mx, sd)
    scale = float(g["logits_absmax"])
    err = float(np.abs(logits[:, :, 3::8, 5::8].numpy() - g["logits_sub"]).max())
    assert err <= 2e-4 * scale, f"oracle port vs reference at 480x640: {err:.3e} of {scale:.3e}"
    pred = logits.argmax(1).numpy().astype(np.uint8)
    assert float((pred == g["argmax"]).mean()) >= 0.9995
    gt = (P.rand(SEED, tag + "/gt", (1, H, W)) * ncls).long().clamp(max=ncls - 1).numpy()
    _, miou = sigma_ref.mean_iou(pred, gt, ncls)
    assert abs(miou - float(g["miou"])) < 2e-4
