"""CPU: checkpoint compatibility.  A checkpoint written with the ORIGINAL VMamba key names
(vmamba.py:2111-2147 lists the renames the reference applies when loading) lands in our backbone tensor for tensor; the
reference's `load_model` wrappers (utils/pyt_utils.py:155-192) behave the same; an EncoderDecoder round-trips strictly."""
import contextlib
import io
import os

import torch

import procedural as P
from helpers import SEED, cfg_tiny


def _old_names(sd):
    out = {}
    for k, v in sd.items():
        k2 = k.replace("patch_embed.0.", "patch_embed.proj.").replace("patch_embed.2.", "patch_embed.norm.")
        k2 = k2.replace(".norm.", ".ln_1.") if ".blocks." in k2 and ".op." not in k2 and ".downsample" not in k2 else k2
        k2 = k2.replace(".op.", ".self_attention.")
        out[k2] = v
    return out


def test_vmamba_pretraining_names_load_into_backbone(tmp_path):
    from sigma_b200 import modules as M
    bb = M.Backbone_VSSM(depths=[1, 1, 2, 1], dims=[32, 64, 128, 256], mlp_ratio=0.0, downsample_version="v1", drop_path_rate=0.1)
    P.fill_state_dict(bb, SEED)
    want = {k: v.clone() for k, v in bb.state_dict().items()}
    old = _old_names(want)
    assert any(".ln_1." in k for k in old) and any(".self_attention." in k for k in old) and "patch_embed.proj.weight" in old
    # classification-head entries of a pretraining checkpoint (dropped: Backbone_VSSM deletes the classifier, strict=False)
    old["norm.weight"] = torch.ones(256)
    old["head.weight"] = torch.zeros(1000, 256)
    path = os.path.join(tmp_path, "vssm_ckpt.pth")
    torch.save({"model": old}, path)
    fresh = M.Backbone_VSSM(depths=[1, 1, 2, 1], dims=[32, 64, 128, 256], mlp_ratio=0.0, downsample_version="v1", drop_path_rate=0.1)
    with contextlib.redirect_stdout(io.StringIO()) as log:
        fresh.load_pretrained(path)
    assert "Successfully load ckpt" in log.getvalue()
    got = fresh.state_dict()
    loaded = [k for k in want if not k.startswith("outnorm")]        # outnorm{i} do not exist in a pretraining checkpoint
    assert loaded and all(torch.equal(got[k], want[k]) for k in loaded)


def test_load_model_wrappers_and_strictness(tmp_path):
    from sigma_b200 import checkpoint, modules as M
    with contextlib.redirect_stdout(io.StringIO()):
        src = M.EncoderDecoder(cfg_tiny(64, 96), criterion=None)
        dst = M.EncoderDecoder(cfg_tiny(64, 96), criterion=None)
    P.fill_state_dict(src, SEED + 3)
    sd = src.state_dict()
    for wrap in ("model", "state_dict", "module", None):
        path = os.path.join(tmp_path, f"ck_{wrap}.pth")
        torch.save({wrap: sd} if wrap else sd, path)
        checkpoint.load_model(dst, path)
        assert all(torch.equal(a, b) for a, b in zip(dst.state_dict().values(), sd.values()))
    checkpoint.load_model(dst, dict(sd))                       # a state_dict object instead of a path
    bad = dict(sd)
    bad.pop(next(iter(bad)))
    try:
        checkpoint.load_model(dst, bad)
        raise AssertionError("strict loading must reject a missing key")
    except RuntimeError:
        pass
    wrapped = torch.nn.Module()
    wrapped.module = dst                                        # what DDP's state_dict looks like
    checkpoint.load_model(wrapped, dict(sd), is_restore=True)
