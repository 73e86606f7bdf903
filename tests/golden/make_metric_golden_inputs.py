"""Procedural inputs of the metric goldens (shared by make_metric_golden.py and tests/test_metric.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import procedural as P  # noqa: E402

CASES = [("m9", 9, 2, 48, 64), ("m40", 40, 1, 37, 53), ("m2", 2, 3, 16, 16)]


def inputs(tag, ncls, B, H, W):
    logits = P.randn(91, f"metric/{tag}/logits", (B, ncls, H, W)).numpy()
    lab = (P.rand(91, f"metric/{tag}/labels", (B, H, W)).numpy() * (ncls + 2)).astype(np.int64)
    lab[lab >= ncls] = 255                                   # ignore label, as in the datasets (config.background = 255)
    return logits, lab.astype(np.uint8)
