"""torch.profiler table of ONE training step (composed path): which kernels hold the step at 2 images per GPU.
    python scripts/profile_train_step.py [--model sigma_tiny] [--batch 2] > gpurun_out/train_profile.txt"""
import argparse
import contextlib
import io
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigma_b200 import modules as M, train_util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="sigma_tiny")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--num-classes", type=int, default=9)
ap.add_argument("--amp", default="none")
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(0)
cfg = types.SimpleNamespace(backbone=a.model, decoder="MambaDecoder", num_classes=a.num_classes, image_height=480, image_width=640,
                            pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1)
with contextlib.redirect_stdout(io.StringIO()):
    model = M.EncoderDecoder(cfg, criterion=torch.nn.CrossEntropyLoss(ignore_index=255)).cuda().train()
opt = train_util.make_optimizer(model)
step = train_util.TrainStep(model, opt, amp_dtype=torch.bfloat16 if a.amp == "bf16" else None)
rgb = torch.randn(a.batch, 3, 480, 640, device="cuda")
mx = torch.randn(a.batch, 3, 480, 640, device="cuda")
gt = torch.randint(0, a.num_classes, (a.batch, 480, 640), device="cuda")
for _ in range(3):
    step(rgb, mx, gt)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(rgb, mx, gt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
