"""GPU parity of the fused Focal-Tversky / Dice loss kernels (C ABI r3d_tversky_loss_fwd / _bwd) against the tensor-op
formulation of the same file (which follows randlanet/utils/losses.py:66-86) and against the oracle's dice loss."""
import importlib

import pytest
import torch

from oracle import network as onet

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["dice", "tversky", "focal_tversky"])
@pytest.mark.parametrize("B,C,N,transposed", [(8, 2, 2500, True), (2, 3, 1100, True), (1, 2, 16384, False),
                                              (3, 5, 77, True)])
def test_tversky_loss_kernels_vs_tensor_ops(name, B, C, N, transposed):
    losses = importlib.import_module("3d_recognizer_b200.losses")
    g = torch.Generator(device="cuda").manual_seed(B * 100 + C)
    base = torch.randn((B, N, C) if transposed else (B, C, N), device="cuda", generator=g) * 2.0
    labels = torch.randint(0, C, (B, N), device="cuda", generator=g)
    scale = torch.tensor(1.7, device="cuda")
    res = {}
    for kernels in (True, False):
        losses.USE_LOSS_KERNELS = kernels
        try:
            x = base.clone().requires_grad_(True)
            logits = x.transpose(1, 2) if transposed else x          # the network hands over a transposed view
            loss = losses.get_loss(name)(logits, labels)
            (loss * scale).backward()
            res[kernels] = (loss.detach().double().cpu(), x.grad.double().cpu())
        finally:
            losses.USE_LOSS_KERNELS = True
    # fp64 evaluation of the tensor-op formula as the arbiter
    xd = base.double().clone().requires_grad_(True)
    ld = losses.get_loss(name)(xd.transpose(1, 2) if transposed else xd, labels)
    (ld * scale.double()).backward()
    assert abs(float(res[True][0]) - float(ld)) < 2e-6
    denom = float(xd.grad.abs().max())
    e_kernel = float((res[True][1] - xd.grad.cpu()).abs().max()) / denom
    e_ops = float((res[False][1] - xd.grad.cpu()).abs().max()) / denom
    assert e_kernel < max(1e-5, 3 * e_ops), (e_kernel, e_ops)
    if name == "dice":
        logits_cpu = (base.transpose(1, 2) if transposed else base).cpu()
        assert abs(float(onet.dice_loss(logits_cpu, labels.cpu())) - float(res[True][0])) < 2e-6


def test_tversky_loss_degenerate_labels():
    """A class that never occurs (|m_c| = 0) and a batch that is all one class: finite loss and gradients."""
    losses = importlib.import_module("3d_recognizer_b200.losses")
    x = torch.randn(2, 2, 500, device="cuda", requires_grad=True)
    for lab in (torch.zeros(2, 500, dtype=torch.int64, device="cuda"), torch.ones(2, 500, dtype=torch.int64, device="cuda")):
        x.grad = None
        loss = losses.get_loss("dice")(x, lab)
        loss.backward()
        ref = onet.dice_loss(x.detach().cpu(), lab.cpu())
        assert torch.isfinite(loss) and torch.isfinite(x.grad).all()
        assert abs(float(loss) - float(ref)) < 2e-6


LOSS_PARAMS = {"dice": (0.5, 1.0), "tversky": (0.7, 1.0), "focal_tversky": (0.7, 4.0 / 3.0)}   # trainer.py:245-269


@pytest.mark.parametrize("name", list(LOSS_PARAMS))
@pytest.mark.parametrize("case", ["b2c2n300", "b3c3n257", "b1c5n64"])
def test_loss_kernels_vs_reference_golden(name, case):
    """Fused loss kernels against the REFERENCE's FocalTverskyLoss (losses.py:59-87) with the parameter sets of
    Trainer._get_loss: loss value and every d loss / d logits entry (fixtures written by oracle/make_golden.py from
    the imported reference class)."""
    import os

    import numpy as np

    from conftest import GOLDEN
    losses = importlib.import_module("3d_recognizer_b200.losses")
    g = np.load(os.path.join(GOLDEN, "loss_golden.npz"))
    x = torch.from_numpy(g[f"{case}/logits"]).cuda().requires_grad_(True)
    labels = torch.from_numpy(g[f"{case}/labels"]).cuda()
    assert losses.USE_LOSS_KERNELS
    loss = losses.get_loss(name)(x, labels)
    loss.backward()
    ref_loss, ref_grad = float(g[f"{case}/{name}/loss"]), torch.from_numpy(g[f"{case}/{name}/dlogits"])
    assert abs(float(loss) - ref_loss) < 1e-6
    assert float((x.grad.cpu() - ref_grad).abs().max()) < 1e-5 * float(ref_grad.abs().max())


@pytest.mark.parametrize("name", list(LOSS_PARAMS))
def test_loss_kernels_vs_oracle_large(name):
    """The three Tversky variants against the oracle's restatement (pinned to the reference by make_golden.py) at a
    training-step size, logits in the transposed layout the network hands over."""
    losses = importlib.import_module("3d_recognizer_b200.losses")
    alpha, gamma = LOSS_PARAMS[name]
    gen = torch.Generator().manual_seed(5)
    base = torch.randn(4, 40960, 2, generator=gen) * 3.0
    labels = (torch.rand(4, 40960, generator=gen) < 0.03).long()            # fingertip-style class imbalance
    x = base.cuda().requires_grad_(True)
    loss = losses.get_loss(name)(x.transpose(1, 2), labels.cuda())
    loss.backward()
    xo = base.clone().double().requires_grad_(True)
    lo = onet.dice_loss(xo.transpose(1, 2), labels, alpha, gamma)
    lo.backward()
    assert abs(float(loss) - float(lo)) < 2e-6
    assert float((x.grad.cpu().double() - xo.grad).abs().max()) < 2e-5 * float(xo.grad.abs().max())
