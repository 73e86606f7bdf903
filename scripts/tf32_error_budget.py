"""Where does the TF32 error of the fused path come from?  Sigma-tiny 480x640 against the reference golden with the
projections of one KIND at a time kept in full precision (and the complements): logits error / label agreement.
    python scripts/tf32_error_budget.py"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import procedural as P  # noqa: E402
from helpers import SEED, cfg_tiny, golden  # noqa: E402
from sigma_b200 import fused, modules as M  # noqa: E402

tag, H, W, ncls = "sigma_tiny_480x640", 480, 640, 9
g = golden(tag)
with contextlib.redirect_stdout(io.StringIO()):
    model = M.EncoderDecoder(cfg_tiny(H, W, num_classes=ncls), criterion=None)
P.fill_state_dict(model, SEED)
model = model.cuda().eval()
rgb = P.randn(SEED, tag + "/rgb", (1, 3, H, W)).cuda()
mx = P.randn(SEED, tag + "/x", (1, 3, H, W)).cuda()
scale = float(g["logits_absmax"])
torch.backends.cudnn.allow_tf32 = False
for name, tf32, kinds in [("all fp32", False, set()), ("all tf32", True, set()), ("tf32, x_proj fp32", True, {"x_proj"}),
                          ("tf32, in_proj fp32", True, {"in_proj"}), ("tf32, out_proj fp32", True, {"out_proj"}),
                          ("tf32, x_proj + in_proj fp32", True, {"x_proj", "in_proj"}), ("tf32, only 'dense' (merge / decoder linears) tf32", True, {"x_proj", "in_proj", "out_proj"}),
                          ("tf32, only x_proj tf32", True, {"in_proj", "out_proj", "dense"})]:
    torch.backends.cuda.matmul.allow_tf32 = tf32
    fused._FP32_KINDS = kinds
    with torch.no_grad():
        logits = model(rgb, mx)
    err = float(np.abs(logits[:, :, 3::8, 5::8].float().cpu().numpy() - g["logits_sub"]).max()) / scale
    agree = float((logits.argmax(1).cpu().numpy().astype(np.uint8) == g["argmax"]).mean())
    print(f"{name:55s} logits err {err:.2e} of scale, labels equal {agree:.5f}", flush=True)
fused._FP32_KINDS = set()
