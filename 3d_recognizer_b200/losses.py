"""Loss functions that drive the training step (reference: randlanet/utils/losses.py; factory
randlanet/utils/trainer.py:245-269).  (B,C,N) logits + (B,N) int64 labels -> scalar.

Written against class-probability sums rather than materialised one-hot tensors: with
p = softmax(logits) over the class axis and m_c = [label == c],
    TP_c = sum m_c p_c,   FN_c = sum m_c (1 - p_c) = |m_c| - TP_c,   FP_c = sum (1 - m_c) p_c = sum p_c - TP_c
which is algebraically the reference's expression (losses.py:66-86)."""
import torch
import torch.nn.functional as F

EPS = 1e-7


USE_LOSS_KERNELS = True       # CUDA logits: three launches instead of ~40 tensor-op launches (csrc/loss.cu)


class _TverskyFn(torch.autograd.Function):
    """Focal-Tversky loss on the sm_100a kernels (C ABI ``r3d_tversky_loss_fwd`` / ``_bwd``)."""

    @staticmethod
    def forward(ctx, logits, labels, alpha, gamma, first):
        from . import _cabi, ops
        B, C, N = logits.shape
        dev = logits.device
        labels = labels.contiguous()
        acc = ops.zeros(3 * C, torch.float64, dev)
        out = torch.empty(1 + 2 * C, dtype=torch.float32, device=dev)          # loss | coef (2,C)
        sb, sc, sn = logits.stride()
        with torch.cuda.device(dev), _cabi.kernel_timer(f"tversky_loss_fwd[M={B * N}]", flops=8.0 * B * N * C,
                                                        bytes=4.0 * B * N * (C + 2)):
            rc = _cabi.lib().r3d_tversky_loss_fwd(_cabi.raw(logits), sb, sc, sn, _cabi.ptr(labels), B, C, N, first,
                                                  float(alpha), float(gamma), EPS, _cabi.ptr(acc), _cabi.raw(out[:1]),
                                                  _cabi.raw(out[1:]), _cabi.stream_ptr(dev))
        _cabi.check(rc, "r3d_tversky_loss_fwd")
        ctx.save_for_backward(logits, labels, out)
        return out[0]

    @staticmethod
    def backward(ctx, gout):
        from . import _cabi
        logits, labels, out = ctx.saved_tensors
        B, C, N = logits.shape
        dev = logits.device
        dlogits = torch.empty_strided(logits.shape, logits.stride(), dtype=torch.float32, device=dev)
        sb, sc, sn = logits.stride()
        gout = gout.contiguous().float()
        with torch.cuda.device(dev), _cabi.kernel_timer(f"tversky_loss_bwd[M={B * N}]", flops=10.0 * B * N * C,
                                                        bytes=4.0 * B * N * (2 * C + 2)):
            rc = _cabi.lib().r3d_tversky_loss_bwd(_cabi.raw(logits), sb, sc, sn, _cabi.ptr(labels), B, C, N,
                                                  _cabi.raw(out[1:]), _cabi.ptr(gout), _cabi.raw(dlogits),
                                                  _cabi.stream_ptr(dev))
        _cabi.check(rc, "r3d_tversky_loss_bwd")
        return dlogits, None, None, None, None


def focal_tversky(logits: torch.Tensor, labels: torch.Tensor, alpha: float, gamma: float,
                  neglect_background: bool = True) -> torch.Tensor:
    C = logits.size(-2)
    if (USE_LOSS_KERNELS and logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 3 and C <= 16
            and labels.dtype == torch.int64 and (C > 1 or not neglect_background)):
        return _TverskyFn.apply(logits, labels, alpha, gamma, 1 if neglect_background else 0)
    p = F.softmax(logits, dim=-2)                               # (B,C,N)
    first = 1 if neglect_background else 0
    terms = []
    for c in range(first, C):
        m = (labels == c).to(p.dtype)                           # (B,N)
        pc = p[:, c, :]
        tp = (m * pc).sum()
        fn = (m * (1 - pc)).sum()
        fp = ((1 - m) * pc).sum()
        ti = (tp + EPS) / (tp + alpha * fn + (1 - alpha) * fp + EPS)
        terms.append((1 - ti) ** gamma)
    return torch.stack(terms).mean()


def focal(logits: torch.Tensor, labels: torch.Tensor, gamma: float = 2.0) -> torch.Tensor:
    B, C, N = logits.size()
    y_true = F.one_hot(labels, C).transpose(-1, -2).to(logits.dtype).clamp(EPS, 1.0 - EPS)
    y_pred = F.softmax(logits, dim=-2).clamp(EPS, 1.0 - EPS)
    return (-y_true * torch.log(y_pred) * (1 - y_pred) ** gamma).sum() / (B * N)


def get_loss(name: str):
    """Same names and default parameters as Trainer._get_loss (trainer.py:245-269)."""
    if name == "cross_entropy":
        return lambda lg, lb: F.cross_entropy(lg, lb)
    if name == "focal":
        return lambda lg, lb: focal(lg, lb, 2.0)
    if name == "dice":
        return lambda lg, lb: focal_tversky(lg, lb, 0.5, 1.0)
    if name == "tversky":
        return lambda lg, lb: focal_tversky(lg, lb, 0.7, 1.0)
    if name == "focal_tversky":
        return lambda lg, lb: focal_tversky(lg, lb, 0.7, 4.0 / 3.0)
    raise ValueError(f"Loss function {name} not known!")
