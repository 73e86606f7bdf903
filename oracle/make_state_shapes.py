"""Writes oracle/state_shapes.json: state_dict key -> shape for sigma_tiny / sigma_small / sigma_base (the reference's
checkpoint contract; identical to the keys the goldens were generated with).  Run once in the build
container:  python oracle/make_state_shapes.py   — the CPU arm of bench.py builds its weights from this file so that
it never imports the product (sigma_b200)."""
import contextlib
import io
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigma_b200 import modules as M  # noqa: E402  (generation only)

out = {}
for bb in ("sigma_tiny", "sigma_small", "sigma_base"):
    cfg = types.SimpleNamespace(backbone=bb, decoder="MambaDecoder", num_classes=9, image_height=480, image_width=640,
                                pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1)
    with contextlib.redirect_stdout(io.StringIO()):
        m = M.EncoderDecoder(cfg, criterion=None)
    out[bb] = {k: list(v.shape) for k, v in m.state_dict().items()}
json.dump(out, open(os.path.join(ROOT, "oracle", "state_shapes.json"), "w"))
print({k: len(v) for k, v in out.items()})
