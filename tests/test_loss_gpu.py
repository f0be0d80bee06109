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
