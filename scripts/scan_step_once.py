"""One eager forward of the benchmark model between cudaProfilerStart / Stop, for ncu captures of exactly one step:
    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        -k regex:ss2d_scan_kernel --csv --log-file gpurun_out/r02_scan_traffic.csv python scripts/scan_step_once.py --batch 74
(then scripts/ncu_scan_traffic.py turns the CSV into profiles/r02_scan_traffic.json, which bench.py reports as roofline.traffic)"""
import argparse
import contextlib
import io
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigma_b200 import modules as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=74)
ap.add_argument("--model", default="sigma_tiny")
ap.add_argument("--height", type=int, default=480)
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--num-classes", type=int, default=9)
ap.add_argument("--precision", default="tf32x3", choices=["tf32x3", "tf32"], help="dense projections, as bench.py --precision")
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = a.precision == "tf32"
torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(0)
cfg = types.SimpleNamespace(backbone=a.model, decoder="MambaDecoder", num_classes=a.num_classes, image_height=a.height,
                            image_width=a.width, pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1)
with contextlib.redirect_stdout(io.StringIO()):
    model = M.EncoderDecoder(cfg, criterion=None).cuda().eval()
g = torch.Generator(device="cuda").manual_seed(1234)
rgb = torch.randn(a.batch, 3, a.height, a.width, device="cuda", generator=g)
mx = torch.randn(a.batch, 3, a.height, a.width, device="cuda", generator=g)
with torch.no_grad():
    model(rgb, mx)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(rgb, mx)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("one step done")
