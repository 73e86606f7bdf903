"""Which torch (non-sigma) ops still run in the fused forward, and from where: torch.profiler with stacks.
    python scripts/prof_torch_ops.py --batch 32"""
import argparse
import contextlib
import io
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import cfg_tiny  # noqa: E402
from sigma_b200 import modules as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = M.EncoderDecoder(cfg_tiny(480, 640, backbone="sigma_tiny"), criterion=None).cuda().eval()
rgb = torch.randn(a.batch, 3, 480, 640, device="cuda")
mx = torch.randn(a.batch, 3, 480, 640, device="cuda")
with torch.no_grad():
    for _ in range(2):
        model(rgb, mx)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        model(rgb, mx)
        torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if not ev.name.startswith("aten::") or ev.device_time_total <= 0 or ev.cpu_children:
        continue
    frames = [f for f in (ev.stack or []) if "sigma_b200" in f]
    where = frames[0].split("sigma_b200/")[-1] if frames else "?"
    k = (ev.name, where)
    agg[k][0] += 1
    agg[k][1] += ev.device_time_total
for (name, where), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us / 1e3:8.3f} ms {n:4d}x {name:28s} {where}")
