"""CPU (build container only: needs /root/reference): the drop-in claim of INTEGRATION.md §1 executed.  With
`sigma_b200/dropin` FIRST on the path and the dependency stand-ins of `sigma_b200/dropin/shims` LAST, the UNMODIFIED
reference's train.py / eval.py import graph resolves (train.py:1-30, eval.py:1-17), `models.builder.EncoderDecoder`
IS sigma_b200's, and the objects train.py builds before its loop — segmodel(cfg=config, criterion, norm_layer),
group_weight, AdamW, WarmUpPolyLR, SegEvaluator's class — construct from the reference's own config."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r'''
import importlib, sys
import torch, torch.nn as nn
mods = ["dataloader.dataloader", "models.builder", "dataloader.RGBXDataset", "utils.init_func", "utils.lr_policy", "engine.engine",
        "engine.logger", "utils.pyt_utils", "utils.visualize", "utils.metric", "eval", "tensorboardX"]      # train.py:13-30
loaded = {m: importlib.import_module(m) for m in mods}
import sigma_b200.modules as M
from models.builder import EncoderDecoder as segmodel                      # train.py:15 / eval.py:16
assert segmodel is M.EncoderDecoder, segmodel
assert loaded["models.builder"].__file__.startswith(sys.argv[1]), loaded["models.builder"].__file__
for m in ("dataloader.dataloader", "utils.init_func", "engine.engine", "eval", "utils.metric"):
    assert loaded[m].__file__.startswith("/root/reference/"), (m, loaded[m].__file__)   # the reference's own files, unmodified
import selective_scan_cuda_core, selective_scan                            # vmamba.py:32,26
assert callable(selective_scan_cuda_core.fwd) and callable(selective_scan_cuda_core.bwd) and callable(selective_scan.selective_scan_fn)
from configs.config_MFNet import config                                    # train.py:42
criterion = nn.CrossEntropyLoss(reduction="mean", ignore_index=config.background)          # train.py:76
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    model = segmodel(cfg=config, criterion=criterion, norm_layer=nn.BatchNorm2d)           # train.py:82
from utils.init_func import group_weight
params_list = group_weight([], model, nn.BatchNorm2d, config.lr)                           # train.py:89-90
opt = torch.optim.AdamW(params_list, lr=config.lr, betas=(0.9, 0.999), weight_decay=config.weight_decay)   # :93
from utils.lr_policy import WarmUpPolyLR
pol = WarmUpPolyLR(config.lr, config.lr_power, config.nepochs * config.niters_per_epoch, config.niters_per_epoch * config.warm_up_epoch)
assert pol.get_lr(0) >= 0
n = sum(p.numel() for p in model.parameters())
assert abs(n - 48.29e6) < 0.05e6, n                                        # Sigma-tiny (SURVEY.md §8b)
from eval import SegEvaluator
assert hasattr(SegEvaluator, "func_per_iteration") and hasattr(SegEvaluator, "sliding_eval_rgbX")
from tensorboardX import SummaryWriter
print("DROPIN_OK", type(model).__module__, len(model.state_dict()), n)
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout exists only in the build container")
def test_reference_train_eval_graph_on_dropin(tmp_path):
    dropin = os.path.join(ROOT, "sigma_b200", "dropin")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, dropin, REF, os.path.join(dropin, "shims")]))
    r = subprocess.run([sys.executable, "-c", SCRIPT, dropin], capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "DROPIN_OK sigma_b200.modules 668" in r.stdout, r.stdout[-500:]
