"""Loss functions that drive the training step (reference: randlanet/utils/losses.py; factory
randlanet/utils/trainer.py:245-269).  (B,C,N) logits + (B,N) int64 labels -> scalar.

Written against class-probability sums rather than materialised one-hot tensors: with
p = softmax(logits) over the class axis and m_c = [label == c],
    TP_c = sum m_c p_c,   FN_c = sum m_c (1 - p_c) = |m_c| - TP_c,   FP_c = sum (1 - m_c) p_c = sum p_c - TP_c
which is algebraically the reference's expression (losses.py:66-86)."""
import torch
import torch.nn.functional as F

EPS = 1e-7


def focal_tversky(logits: torch.Tensor, labels: torch.Tensor, alpha: float, gamma: float,
                  neglect_background: bool = True) -> torch.Tensor:
    C = logits.size(-2)
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
