"""Per-batch training metrics without per-batch host synchronisation (reference: randlanet/utils/metrics.py:8-59 as
driven by trainer.py:121-131, which does 2 + 3C ``.cpu().item()`` round trips per batch).

Every value ``accuracy`` / ``iou`` return is a function of the batch's C x C confusion matrix
``counts[label][prediction]``.  ``MetricAccumulator.push`` adds one batch's matrix into its own device slot with one
kernel launch (C ABI ``r3d_confusion_counts``) and returns immediately; ``collect`` reads all slots at once — one
device -> host copy per epoch — and evaluates the reference's formulas, including its conventions for absent classes
(per-class accuracy 1.0 when a class has no labels and no correct predictions, IoU 1.0 for an empty union)."""
from typing import List, Tuple

import numpy as np
import torch

from . import _cabi


def confusion_counts(logits: torch.Tensor, labels: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """logits (B,C,N) fp32 CUDA (any strides), labels (B,N) int64 -> int64 (C,C) counts[label][prediction], ADDED into
    ``out`` when given (zero-filled otherwise).  No host synchronisation."""
    _cabi.require_cuda(logits, "logits")
    if logits.dim() == 2:
        logits, labels = logits.unsqueeze(0), labels.unsqueeze(0)
    B, C, N = logits.shape
    if out is None:
        out = torch.zeros((C, C), dtype=torch.int64, device=logits.device)
    labels = labels.contiguous()
    sb, sc, sn = logits.stride()
    with torch.cuda.device(logits.device), _cabi.kernel_timer(f"confusion_counts[M={B * N}]", flops=float(B * N * C),
                                                              bytes=4.0 * B * N * (C + 2)):
        rc = _cabi.lib().r3d_confusion_counts(_cabi.raw(logits.detach()), sb, sc, sn, _cabi.ptr(labels), B, C, N,
                                              _cabi.raw(out), _cabi.stream_ptr(logits.device))
    _cabi.check(rc, "r3d_confusion_counts")
    return out


def accuracy_from_counts(cm: np.ndarray) -> Tuple[float, List[float]]:
    """metrics.py:8-33 from a confusion matrix: overall accuracy and per-class accuracies."""
    cm = np.asarray(cm, dtype=np.float64)
    total = cm.sum()
    overall = float(np.float32(np.trace(cm)) / np.float32(total)) if total > 0 else float("nan")
    per_class = []
    for c in range(cm.shape[0]):
        n_labels, correct = cm[c].sum(), cm[c, c]
        per_class.append(float(correct == 0) if n_labels == 0 else float(np.float32(correct) / np.float32(n_labels)))
    return overall, per_class


def iou_from_counts(cm: np.ndarray) -> Tuple[float, List[float]]:
    """metrics.py:36-59 from a confusion matrix: mean IoU and per-class IoUs."""
    cm = np.asarray(cm, dtype=np.float64)
    per_class = []
    for c in range(cm.shape[0]):
        inter = cm[c, c]
        union = cm[c].sum() + cm[:, c].sum() - inter
        per_class.append(1.0 if union == 0 else float(np.float32(inter) / np.float32(union)))
    return float(np.nanmean(per_class)), per_class


class MetricAccumulator:
    """Confusion matrices of up to ``capacity`` batches, kept on the device until ``collect``."""

    def __init__(self, n_classes: int, capacity: int, device):
        self.slots = torch.zeros((capacity, n_classes, n_classes), dtype=torch.int64, device=device)
        self.n = 0

    def push(self, logits: torch.Tensor, labels: torch.Tensor) -> None:
        if self.n >= self.slots.shape[0]:
            raise RuntimeError("MetricAccumulator is full: collect() first or size it for the epoch")
        confusion_counts(logits, labels, self.slots[self.n])
        self.n += 1

    def collect(self):
        """One device -> host copy.  Returns, per pushed batch, (overall accuracy, per-class accuracies, mIoU, per-class
        IoUs) exactly as the reference computes them batch by batch, and resets the accumulator."""
        cms = self.slots[:self.n].cpu().numpy()
        self.slots.zero_()
        self.n = 0
        return [accuracy_from_counts(cm) + iou_from_counts(cm) for cm in cms]
