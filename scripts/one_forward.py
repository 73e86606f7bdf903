"""One eager Sigma-tiny forward bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python scripts/one_forward.py
"""
import argparse
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import cfg_tiny  # noqa: E402
from sigma_b200 import modules as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--height", type=int, default=480)
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--model", default="sigma_tiny")
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = M.EncoderDecoder(cfg_tiny(a.height, a.width, backbone=a.model), criterion=None).cuda().eval()
rgb = torch.randn(a.batch, 3, a.height, a.width, device="cuda")
mx = torch.randn(a.batch, 3, a.height, a.width, device="cuda")
with torch.no_grad():
    for _ in range(a.warm):
        model(rgb, mx)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(rgb, mx)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
