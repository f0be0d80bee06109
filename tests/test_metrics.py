"""Training metrics: the confusion-count kernel + host formulas against golden values written by the
REFERENCE's randlanet/utils/metrics.py (accuracy, iou) in oracle/make_golden.py."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

CASES = ["b2c2n300", "b3c3n257", "b1c5n64"]


def _golden(case, tag):
    g = np.load(os.path.join(GOLDEN, "loss_golden.npz"))
    logits, labels = g[f"{case}/logits"], g[f"{case}/labels"]
    C = logits.shape[1]
    if tag:
        labels = np.minimum(labels, max(C - 2, 0))
    return logits, labels, g[f"{case}/{tag}metrics"], C


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag", ["", "absent/"])
def test_metric_formulas_from_counts_match_reference(case, tag):
    """Host side only: counts built with numpy -> the reference's overall / per-class accuracy, mIoU, per-class IoU."""
    metrics = importlib.import_module("3d_recognizer_b200.metrics")
    logits, labels, ref, C = _golden(case, tag)
    pred = logits.argmax(axis=1)
    cm = np.zeros((C, C), dtype=np.int64)
    np.add.at(cm, (labels.reshape(-1), pred.reshape(-1)), 1)
    oa, pca = metrics.accuracy_from_counts(cm)
    miou, pci = metrics.iou_from_counts(cm)
    got = np.array([oa, miou] + pca + pci)
    assert np.allclose(got, ref, rtol=0, atol=1e-7), (got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("transposed", [False, True])
def test_confusion_kernel_and_accumulator_match_reference(case, transposed):
    metrics = importlib.import_module("3d_recognizer_b200.metrics")
    acc = None
    refs = []
    for tag in ("", "absent/"):
        logits, labels, ref, C = _golden(case, tag)
        x = torch.from_numpy(logits).cuda()
        if transposed:                                   # the network hands over a (B,N,C)-major tensor viewed as (B,C,N)
            x = x.transpose(1, 2).contiguous().transpose(1, 2)
        if acc is None:
            acc = metrics.MetricAccumulator(C, 4, x.device)
        acc.push(x, torch.from_numpy(labels).cuda())
        refs.append(ref)
    out = acc.collect()
    assert acc.n == 0 and len(out) == 2
    for (oa, pca, miou, pci), ref in zip(out, refs):
        assert np.allclose(np.array([oa, miou] + pca + pci), ref, rtol=0, atol=1e-7)


@pytest.mark.gpu
def test_confusion_counts_large_batch_sums_to_the_point_count():
    metrics = importlib.import_module("3d_recognizer_b200.metrics")
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn(8, 40960, 2, device="cuda", generator=g).transpose(1, 2)
    labels = (torch.rand(8, 40960, device="cuda", generator=g) < 0.03).long()
    cm = metrics.confusion_counts(logits, labels).cpu().numpy()
    pred = logits.argmax(dim=1)
    ref = np.zeros((2, 2), dtype=np.int64)
    np.add.at(ref, (labels.cpu().numpy().reshape(-1), pred.cpu().numpy().reshape(-1)), 1)
    assert cm.sum() == 8 * 40960 and np.array_equal(cm, ref)
