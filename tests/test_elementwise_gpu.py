"""GPU parity of the element-wise kernels of the training step (csrc/elementwise.cu): residual sum + LeakyReLU
(modules.py:325) forward / backward against torch, and the Adam update (trainer.py:78-81) against torch.optim.Adam —
eager, with a learning-rate tensor a scheduler writes into, and across a state_dict round trip."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 1000, 32), (1, 37, 8), (3, 4099, 4), (5, 7)])
def test_add_lrelu_vs_torch(shape):
    ops = importlib.import_module("3d_recognizer_b200.ops")
    engine = importlib.import_module("3d_recognizer_b200.engine")
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    a = torch.randn(shape, device="cuda", generator=g).requires_grad_(True)
    b = torch.randn(shape, device="cuda", generator=g).requires_grad_(True)
    a2, b2 = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    dy = torch.randn(shape, device="cuda", generator=g)
    y = engine.residual_lrelu(a, b, 0.01)
    assert "_AddLReluFn" in type(y.grad_fn).__name__
    ref = torch.nn.functional.leaky_relu(a2 + b2, 0.01)
    assert torch.equal(y, ref)
    y.backward(dy)
    ref.backward(dy)
    assert torch.equal(a.grad, a2.grad) and torch.equal(b.grad, b2.grad)
    assert torch.equal(ops.add_lrelu(a.detach(), b.detach(), 0.2), torch.nn.functional.leaky_relu(a2 + b2, 0.2).detach())


@pytest.mark.parametrize("tensor_lr", [False, True])
def test_adam_step_vs_torch(tensor_lr):
    ops = importlib.import_module("3d_recognizer_b200.ops")
    n = 100003
    g = torch.Generator(device="cuda").manual_seed(3)
    p = torch.randn(n, device="cuda", generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-2)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros((), device="cuda")
    lr = torch.tensor(1e-2, device="cuda") if tensor_lr else 1e-2
    for it in range(5):
        grad = torch.randn(n, device="cuda", generator=g) * (0.1 + it)
        ref.grad = grad.clone()
        opt.step()
        ops.adam_step(p, grad, m, v, step, lr, 0.9, 0.999, 1e-8)
        sched.step()
        cur = opt.param_groups[0]["lr"]
        if tensor_lr:
            lr.fill_(cur)
        else:
            lr = cur
    assert float(step) == 5.0
    assert torch.allclose(p, ref.detach(), rtol=2e-6, atol=2e-7)
    st = opt.state[ref]
    # the moments differ by the rounding of one fused multiply-add per step (absolute size ~ ulp of the gradient)
    assert torch.allclose(m, st["exp_avg"], rtol=1e-5, atol=1e-6) and torch.allclose(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-7)
