"""Golden outputs of the reference's OWN metric code (utils/metric.py: hist_info, compute_score) on procedural inputs.
Run in the build container (needs /root/reference):  python tests/golden/make_metric_golden.py
Only the reference's outputs are stored; inputs are regenerated from tests/procedural.py by the tests."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
from utils.metric import compute_score, hist_info  # noqa: E402  (the unmodified reference)

from make_metric_golden_inputs import CASES, inputs  # noqa: E402


if __name__ == "__main__":
    out = {}
    for tag, ncls, B, H, W in CASES:
        logits, lab = inputs(tag, ncls, B, H, W)
        hist = np.zeros((ncls, ncls)); labeled = 0; correct = 0
        for b in range(B):                                   # the evaluator's per-image loop (eval.py:22-29, 62-72)
            pred = np.exp(logits[b]).transpose(1, 2, 0).argmax(2)   # evaluator.py:520 (exp) + :449 (argmax)
            h, l, c = hist_info(ncls, pred, lab[b])
            hist += h; labeled += l; correct += c
        iou, miou, _, fiou, macc, pacc = compute_score(hist, correct, labeled)
        out[f"{tag}_hist"] = hist.astype(np.int64)
        out[f"{tag}_labeled"] = np.int64(labeled)
        out[f"{tag}_correct"] = np.int64(correct)
        out[f"{tag}_miou"] = np.float64(miou)
        out[f"{tag}_pixacc"] = np.float64(pacc)
        print(tag, "mIoU", miou, "pixel acc", pacc, "labeled", labeled)
    np.savez_compressed(os.path.join(HERE, "metric.npz"), **out)
