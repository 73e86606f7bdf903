"""SURVEY.md §8(f) rank 2 (first piece): the evaluator's metric.  CPU: the oracle restatement of utils/metric.py against
goldens produced by the reference's own hist_info / compute_score (tests/golden/make_metric_golden.py).  GPU: the device
kernel (sigma_argmax_hist_fwd) bit-exact against the same goldens and, at the full 480x640 size, against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_metric_golden_inputs import CASES, inputs  # noqa: E402
from oracle import sigma_ref  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "metric.npz"))


@pytest.mark.parametrize("tag,ncls,B,H,W", CASES)
def test_oracle_metric_matches_reference(tag, ncls, B, H, W):
    logits, lab = inputs(tag, ncls, B, H, W)
    hist = np.zeros((ncls, ncls), np.int64); labeled = correct = 0
    for b in range(B):
        h, l, c = sigma_ref.hist_info(ncls, logits[b].argmax(0), lab[b])
        hist += h; labeled += l; correct += c
    assert np.array_equal(hist, G[f"{tag}_hist"]) and labeled == int(G[f"{tag}_labeled"]) and correct == int(G[f"{tag}_correct"])
    iou, miou, _, _, pacc = sigma_ref.compute_score(hist, correct, labeled)
    assert abs(miou - float(G[f"{tag}_miou"])) < 1e-12 and abs(pacc - float(G[f"{tag}_pixacc"])) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("tag,ncls,B,H,W", CASES)
@pytest.mark.parametrize("ldtype", [torch.uint8, torch.int32, torch.int64])
def test_device_metric_bit_exact(tag, ncls, B, H, W, ldtype):
    from sigma_b200.evaluator import DeviceMetric
    logits, lab = inputs(tag, ncls, B, H, W)
    m = DeviceMetric(ncls)
    lt = torch.from_numpy(lab.astype(np.int64)).to(ldtype).cuda()
    pred = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
    for b in range(B):                                   # accumulate per image, as the evaluator does
        m.update(torch.from_numpy(logits[b:b + 1]).cuda(), lt[b:b + 1], pred_out=pred[b:b + 1])
    hist, labeled, correct = m.result()
    assert np.array_equal(hist, G[f"{tag}_hist"]), "confusion matrix differs from the reference's hist_info"
    assert labeled == int(G[f"{tag}_labeled"]) and correct == int(G[f"{tag}_correct"])
    assert np.array_equal(pred.cpu().numpy(), logits.argmax(1).astype(np.uint8))
    _, miou, _, _, pacc = DeviceMetric.compute_score(hist, correct, labeled)
    assert abs(miou - float(G[f"{tag}_miou"])) < 1e-12 and abs(pacc - float(G[f"{tag}_pixacc"])) < 1e-12


@pytest.mark.gpu
def test_device_metric_fullsize_with_ties():
    """480x640, 9 classes, a batch at once; ties (equal logits) must resolve to the first maximum as numpy.argmax does."""
    from sigma_b200.evaluator import DeviceMetric
    rng = np.random.default_rng(5)
    B, ncls, H, W = 3, 9, 480, 640
    logits = rng.integers(-3, 4, size=(B, ncls, H, W)).astype(np.float32)      # many exact ties
    lab = rng.integers(0, ncls + 3, size=(B, H, W)).astype(np.int64)
    lab[lab >= ncls] = 255
    m = DeviceMetric(ncls)
    m.update(torch.from_numpy(logits).cuda(), torch.from_numpy(lab.astype(np.uint8)).cuda())
    hist, labeled, correct = m.result()
    rh, rl, rc = sigma_ref.hist_info(ncls, logits.argmax(1), lab)
    assert np.array_equal(hist, rh) and labeled == rl and correct == rc


def test_device_metric_refuses_cpu_tensors():
    pytest.importorskip("torch")
    from sigma_b200.evaluator import DeviceMetric
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(Exception):
        DeviceMetric(9, device="cpu").update(torch.zeros(1, 9, 4, 4), torch.zeros(1, 4, 4, dtype=torch.uint8))
