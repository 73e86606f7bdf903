"""Device-side segmentation metric: the reference's evaluator moves every image's score map to
the host, takes argmax and builds the confusion matrix with numpy (eval.py:22-29, utils/metric.py:8-33).  Here the
argmax and the confusion matrix are one kernel over the logits that are already in HBM; only classes² + 2 integers
are read back, once per evaluation."""
import numpy as np
import torch

from . import _lib


class DeviceMetric:
    """m = DeviceMetric(num_classes); m.update(logits, labels) per batch; hist, labeled, correct = m.result().

    `logits`: (B, classes, H, W) fp32 CUDA tensor; `labels`: (B, H, W) uint8 / int32 / int64 CUDA tensor, pixels whose
    label is outside [0, classes) (255 in the reference's datasets) are ignored — `hist_info`'s `k` mask."""

    def __init__(self, num_classes, device="cuda"):
        self.n = int(num_classes)
        self.hist = torch.zeros(self.n * self.n, dtype=torch.int64, device=device)
        self.counts = torch.zeros(2, dtype=torch.int64, device=device)

    def update(self, logits, labels, pred_out=None):
        if not (logits.is_cuda and labels.is_cuda):
            raise RuntimeError("sigma_b200.DeviceMetric works on CUDA tensors only (there is no CPU path)")
        if logits.dtype != torch.float32:
            raise TypeError("logits must be float32")
        B, C, H, W = logits.shape
        if C != self.n or tuple(labels.shape) != (B, H, W):
            raise ValueError(f"logits {tuple(logits.shape)} / labels {tuple(labels.shape)} do not match {self.n} classes")
        lb = {torch.uint8: 1, torch.int32: 4, torch.int64: 8}.get(labels.dtype)
        if lb is None:
            raise TypeError("labels must be uint8, int32 or int64")
        logits, labels = logits.contiguous(), labels.contiguous()
        pp = pred_out.data_ptr() if pred_out is not None else None
        rc = _lib.lib().sigma_argmax_hist_fwd(logits.data_ptr(), labels.data_ptr(), lb, self.hist.data_ptr(), self.counts.data_ptr(),
                                              pp, B, C, H * W, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "sigma_argmax_hist_fwd")

    def result(self):
        h = self.hist.cpu().numpy().reshape(self.n, self.n)
        c = self.counts.cpu().numpy()
        return h, int(c[0]), int(c[1])

    @staticmethod
    def compute_score(hist, correct, labeled):
        """utils/metric.py:17-33: (iou per class, mIoU, freq-weighted IoU, mean class accuracy, pixel accuracy)."""
        hist = np.asarray(hist, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
            freq = hist.sum(1) / hist.sum()
            acc = np.diag(hist) / hist.sum(axis=1)
            return iou, float(np.nanmean(iou)), float((iou[freq > 0] * freq[freq > 0]).sum()), float(np.nanmean(acc)), correct / labeled
