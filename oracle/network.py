"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional CPU restatement (plain torch fp32 ops, autograd-capable) of the reference network
``randlanet/utils/modules.py``.  It is driven by a *state_dict* with the reference's
parameter names, so it also pins the checkpoint schema.

Every function cites the reference lines it follows.  KNN is the canonical exact search of
``oracle.knn.knn_exact`` for BOTH the encoder and the decoder's 1-NN (the reference's intended
``kdtree`` wiring, modules.py:135-138; the shipped FAISS / matmul back-ends are approximate /
inexact — SURVEY.md F6-F8).

Pinned against the reference's own modules imported from /root/reference by
``oracle/make_golden.py`` (max |Δlogit| printed there; fixtures in tests/golden/).
"""
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .knn import knn_exact

BN_EPS = 1e-6        # modules.py:87, :497
BN_MOMENTUM = 0.99   # modules.py:87, :497  (running = 0.01*old + 0.99*new)

DEFAULT_SETTINGS = dict(n_classes=2, n_points=10000, n_features=0, n_neighbors=32, decimation=4,
                        layer_sizes=[16, 64, 128, 256], knn="approximate", upsampling="nni")


def state_dict_schema(settings: dict) -> "Dict[str, Tuple[Tuple[int, ...], torch.dtype]]":
    """Parameter/buffer names, shapes and dtypes of ``RandLANet.state_dict()``
    (module tree of modules.py:495-530, SharedMLP :83-90, AttentivePooling :234-237)."""
    s = dict(DEFAULT_SETTINGS, **settings)
    out: Dict[str, Tuple[Tuple[int, ...], torch.dtype]] = {}

    def bn(prefix, c):
        out[prefix + ".weight"] = ((c,), torch.float32)
        out[prefix + ".bias"] = ((c,), torch.float32)
        out[prefix + ".running_mean"] = ((c,), torch.float32)
        out[prefix + ".running_var"] = ((c,), torch.float32)
        out[prefix + ".num_batches_tracked"] = ((), torch.int64)

    def smlp(prefix, cin, cout, transpose=False, with_bn=True):
        out[prefix + ".conv.weight"] = ((cin, cout, 1, 1) if transpose else (cout, cin, 1, 1), torch.float32)
        out[prefix + ".conv.bias"] = ((cout,), torch.float32)
        if with_bn:
            bn(prefix + ".batch_norm", cout)

    n_in = 3 + s["n_features"]
    out["fc_start.weight"] = ((8, n_in), torch.float32)
    out["fc_start.bias"] = ((8,), torch.float32)
    bn("bn_start.0", 8)
    c = 8
    for l, d in enumerate(s["layer_sizes"]):
        p = f"encoder.{l}"
        smlp(p + ".mlp1", c, d // 2)
        smlp(p + ".mlp2", d, 2 * d)
        smlp(p + ".shortcut", c, 2 * d)
        smlp(p + ".mlp_rpe1", 10, d // 2)
        smlp(p + ".mlp_rpe2", d // 2, d // 2)
        out[p + ".pool1.score_fn.0.weight"] = ((d, d), torch.float32)
        smlp(p + ".pool1.mlp", d, d // 2)
        out[p + ".pool2.score_fn.0.weight"] = ((d, d), torch.float32)
        smlp(p + ".pool2.mlp", d, d)
        c = 2 * d
    smlp("mlp", c, c)
    c *= 2
    for j, d in enumerate(s["layer_sizes"][::-1][1:]):
        smlp(f"decoder.{j}", c, 2 * d, transpose=True)
        c = 4 * d
    smlp(f"decoder.{len(s['layer_sizes']) - 1}", c, 8, transpose=True)
    smlp("fc_end.0", 8, 64)
    smlp("fc_end.1", 64, 32)
    smlp("fc_end.3", 32, s["n_classes"], with_bn=False)
    return out


def synth_state_dict(settings: dict, seed: int = 0) -> "Dict[str, torch.Tensor]":
    """Deterministic non-trivial weights (numpy MT19937, key order of the schema) shared by the
    golden generator and the tests; BN running stats are non-default so eval-mode folding is
    exercised.  Not an init scheme of the reference — a test vector."""
    rng = np.random.RandomState(seed)
    sd = {}
    for name, (shape, dtype) in state_dict_schema(settings).items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(7, dtype=torch.int64)
        elif name.endswith("running_var"):
            sd[name] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif name.endswith("running_mean"):
            sd[name] = torch.from_numpy(rng.normal(0, 0.2, shape).astype(np.float32))
        elif name.endswith("batch_norm.weight") or name == "bn_start.0.weight":
            sd[name] = torch.from_numpy(rng.uniform(0.7, 1.3, shape).astype(np.float32))
        elif name.endswith(".bias"):
            sd[name] = torch.from_numpy(rng.normal(0, 0.1, shape).astype(np.float32))
        else:
            fan_in = shape[0] if (name.startswith("decoder") and name.endswith("conv.weight")) else shape[1]
            sd[name] = torch.from_numpy((rng.normal(0, 1, shape) * np.sqrt(2.0 / fan_in)).astype(np.float32))
    return sd


# --------------------------------------------------------------------------- KNN (modules.py:107-150)
def knn(xyz: torch.Tensor, xyz_query: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """KNN.forward: indices int64 (B,N,K) and NON-squared distances (sqrt at modules.py:149)."""
    idx, d2 = knn_exact(xyz.detach().cpu().numpy(), xyz_query.detach().cpu().numpy(), k)
    return torch.from_numpy(idx), torch.sqrt(torch.from_numpy(d2)).to(xyz.dtype)


# --------------------------------------------------------------------------- activation branches (test instrument)
class Kinks:
    """Instrument for the piecewise-linear activations of the network (ReLU / LeakyReLU, modules.py:96-103, :325,
    :565-566).  The reference's arithmetic is unchanged; this only (a) RECORDS, per activation site in call order, the
    elements whose pre-activation lies within ``threshold`` x rms(site) of the kink, and (b) lets a test PIN branches:
    ``pins[site]`` = (flat indices, branch values): those elements take the given branch (True = the positive one)
    whatever sign the run itself computes.  Two correct fp32 evaluations — this port included: its threaded CPU
    reductions are not bit-reproducible from run to run — may put such an element on different sides of the kink (see
    grad_parity); with the ambiguous elements enumerated by an fp64 run and ALL of them pinned, every oracle run is
    deterministic in its branches and every gradient entry can be held to the strict tolerance with no outliers
    (tests/test_kink_pinned_gpu.py)."""

    def __init__(self, threshold: float = 0.0, pins: "Optional[Dict[int, Tuple[torch.Tensor, torch.Tensor]]]" = None):
        self.threshold, self.pins = threshold, pins or {}
        self.site = 0
        self.found: "List[Tuple[int, int, float]]" = []     # (site, flat index, pre-activation / rms)

    def activate(self, u: torch.Tensor, slope: float) -> torch.Tensor:
        site = self.site
        self.site += 1
        if self.threshold > 0:
            rms = float(u.detach().double().pow(2).mean().sqrt())
            near = (u.detach().abs() <= self.threshold * rms).reshape(-1).nonzero().reshape(-1)
            flat = u.detach().reshape(-1)
            self.found += [(site, int(i), float(flat[i]) / max(rms, 1e-300)) for i in near]
        mask = u > 0
        if site in self.pins:
            mask = mask.contiguous().clone()          # flat indices are in logical (row-major) order, as recorded
            idx, val = self.pins[site]
            mask.view(-1)[idx] = val
        return torch.where(mask, u, u * slope)


_KINKS: "Optional[Kinks]" = None


class kinks:
    """``with kinks(Kinks(...)):`` routes every activation of ``forward`` through the instrument."""

    def __init__(self, k: Kinks):
        self.k = k

    def __enter__(self):
        global _KINKS
        self.prev, _KINKS = _KINKS, self.k
        return self.k

    def __exit__(self, *exc):
        global _KINKS
        _KINKS = self.prev
        return False


def _activation(y: torch.Tensor, slope: float) -> torch.Tensor:
    """relu (slope 0) / leaky_relu; same values as F.relu / F.leaky_relu."""
    if _KINKS is not None:
        return _KINKS.activate(y, slope)
    return F.relu(y) if slope == 0.0 else F.leaky_relu(y, slope)


# --------------------------------------------------------------------------- SharedMLP (modules.py:60-104)
def _bn(sd, prefix, x, training):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)


def shared_mlp(sd, prefix, x, training, act=None, transpose=False, with_bn=True):
    w, b = sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"]
    y = F.conv_transpose2d(x, w, b) if transpose else F.conv2d(x, w, b)
    if with_bn:
        y = _bn(sd, prefix + ".batch_norm", y, training)
        if training:
            sd[prefix + ".batch_norm.num_batches_tracked"] += 1
    if act == "relu":
        y = _activation(y, 0.0)
    elif isinstance(act, float):
        y = _activation(y, act)
    return y


# --------------------------------------------------------------------------- LocSE (modules.py:153-221)
def relative_position_encoding(xyz, idx, dist):
    """cat[p_i, p_j, p_i - p_j, |p_i - p_j|] -> (B,10,N,K)  (modules.py:170-186)."""
    B, N, K = idx.shape
    pj = xyz[torch.arange(B)[:, None, None], idx]          # neighbour coordinates (B,N,K,3)
    pi = xyz.unsqueeze(2).expand(B, N, K, 3)
    enc = torch.cat((pi, pj, pi - pj, dist.unsqueeze(-1)), dim=-1)  # (B,N,K,10)
    return enc.permute(0, 3, 1, 2)


def gather_neighbors(feat, idx):
    """feat (B,C,N,1), idx (B,N,K) -> (B,C,N,K)  (modules.py:209-217)."""
    B, C = feat.shape[:2]
    _, N, K = idx.shape
    flat = idx.reshape(B, 1, N * K).expand(B, C, N * K)
    return torch.gather(feat.squeeze(-1), 2, flat).reshape(B, C, N, K)


# --------------------------------------------------------------------------- AttentivePooling (modules.py:224-253)
def attentive_pooling(sd, prefix, x, training):
    w = sd[prefix + ".score_fn.0.weight"]                       # Linear(d,d,bias=False)
    scores = F.softmax(F.linear(x.permute(0, 2, 3, 1), w), dim=-2)  # softmax over K
    pooled = torch.sum(scores.permute(0, 3, 1, 2) * x, dim=-1, keepdim=True)
    return shared_mlp(sd, prefix + ".mlp", pooled, training, act="relu")


# --------------------------------------------------------------------------- LFA (modules.py:298-325)
def local_feature_aggregation(sd, prefix, xyz, feat, k, training):
    idx, dist = knn(xyz, xyz, k)
    f = shared_mlp(sd, prefix + ".mlp1", feat, training, act=0.2)
    r1 = shared_mlp(sd, prefix + ".mlp_rpe1", relative_position_encoding(xyz, idx, dist), training, act="relu")
    p1 = attentive_pooling(sd, prefix + ".pool1", torch.cat((r1, gather_neighbors(f, idx)), dim=1), training)
    r2 = shared_mlp(sd, prefix + ".mlp_rpe2", r1, training, act="relu")     # input is r1 (modules.py:321)
    p2 = attentive_pooling(sd, prefix + ".pool2", torch.cat((r2, gather_neighbors(p1, idx)), dim=1), training)
    return _activation(shared_mlp(sd, prefix + ".mlp2", p2, training)
                       + shared_mlp(sd, prefix + ".shortcut", feat, training), 0.01)


# --------------------------------------------------------------------------- UpSampler (modules.py:328-456)
def upsample(approach: str, feat, xyz, xyz_up):
    """feat (B,F,N1,1), xyz (B,N1,3), xyz_up (B,N2,3) -> (B,F,N2,1)."""
    if approach == "none":
        return feat
    if approach == "nni":                                            # modules.py:343-364
        idx, _ = knn(xyz, xyz_up, 1)
        return gather_neighbors(feat, idx)
    if approach in ("nna", "idw", "isdw"):                           # modules.py:366-414, :434-449
        # note: "nna" also reaches inverse_distance_weighting=True (default arg, modules.py:372,435)
        power = 2.0 if approach == "isdw" else 1.0
        idx, dist = knn(xyz, xyz_up, 8)
        nf = gather_neighbors(feat, idx)
        eps = 1e-7
        w = (1.0 + eps) / (dist ** power + eps)
        w = w / torch.sum(w, dim=-1, keepdim=True)
        return torch.sum(w.unsqueeze(1) * nf, dim=-1, keepdim=True)
    raise ValueError(f"Upsampling approach {approach} not understood!")


# --------------------------------------------------------------------------- RandLANet.forward (modules.py:542-611)
def min_n_points(settings: dict) -> int:
    s = dict(DEFAULT_SETTINGS, **settings)
    L = len(s["layer_sizes"])
    return max(s["n_neighbors"] * s["decimation"] ** (L - 1), 2 * s["decimation"] ** L)


def forward(sd: "Dict[str, torch.Tensor]", settings: dict, inp: torch.Tensor,
            permutation: Optional[np.ndarray] = None, training: bool = False,
            dropout_p: float = 0.5) -> torch.Tensor:
    """inp (B,N,3+F) -> logits (B,C,N).  ``permutation`` None draws np.random.permutation(N) from the
    global numpy RNG at the same point as modules.py:571.  In training mode the BN running
    statistics inside ``sd`` are updated in place, like the reference's buffers."""
    s = dict(DEFAULT_SETTINGS, **settings)
    B, N, dim = inp.shape
    assert dim == 3 + s["n_features"], "Input should have shape (B, N, 3 + F)!"
    assert N >= min_n_points(s), f"Input point cloud should have at least {min_n_points(s)} points!"
    dec, k, L = s["decimation"], s["n_neighbors"], len(s["layer_sizes"])

    xyz = inp[..., :3] if inp.dtype == torch.float64 else inp[..., :3].float()   # fp64 runs are the arbiter in tests
    feat = F.linear(inp, sd["fc_start.weight"], sd["fc_start.bias"]).transpose(-2, -1).unsqueeze(-1)
    feat = _activation(_bn(sd, "bn_start.0", feat, training), 0.2)
    if training:
        sd["bn_start.0.num_batches_tracked"] += 1
    if permutation is None:
        permutation = np.random.permutation(N)
    perm = torch.from_numpy(np.asarray(permutation))
    xyz, feat = xyz[:, perm], feat[:, :, perm]

    stack: List[torch.Tensor] = []
    ratio = 1
    xyz_l, feat_l = xyz, feat
    for l in range(L):
        feat = local_feature_aggregation(sd, f"encoder.{l}", xyz_l, feat_l, k, training)
        stack.append(feat)
        ratio *= dec
        xyz_l, feat_l = xyz[:, : N // ratio], feat[:, :, : N // ratio]
    feat = shared_mlp(sd, "mlp", feat_l, training, act="relu")
    for j in range(L):
        up = upsample("nni", feat, xyz[:, : N // ratio], xyz[:, : dec * N // ratio])
        feat = shared_mlp(sd, f"decoder.{j}", torch.cat((up, stack.pop()), dim=1), training, act="relu",
                          transpose=True)
        ratio //= dec
    feat = feat[:, :, torch.argsort(perm)]
    feat = shared_mlp(sd, "fc_end.0", feat, training, act="relu")
    feat = shared_mlp(sd, "fc_end.1", feat, training, act="relu")
    feat = F.dropout(feat, dropout_p, training)
    return shared_mlp(sd, "fc_end.3", feat, training, with_bn=False).squeeze(-1)


# --------------------------------------------------------------------------- Model.predict pieces (model.py:123-235)
def sample_points(n_points: int, n_sample_points: int, consistent: bool = False) -> np.ndarray:
    """preprocessing.py:35-62 + :6-32: np.random.choice without replacement, then (when up-sampling)
    with replacement; ``consistent`` reseeds the GLOBAL numpy RNG to 0 for each draw and restores it."""
    def choice(a, size, replace):
        if consistent:
            st = np.random.get_state()
            np.random.seed(0)
        v = np.random.choice(a, size, replace, None)
        if consistent:
            np.random.set_state(st)
        return v
    ids = choice(n_points, min(n_sample_points, n_points), False)
    if n_sample_points > n_points:
        ids = np.r_[ids, choice(n_points, n_sample_points - n_points, True)]
    return ids


def predict(sd, settings: dict, xyz: np.ndarray, prepostprocess: bool = True) -> np.ndarray:
    """Model.predict for n_features=0 (model.py:146-235): consistent pre-sampling, eval forward, class
    softmax, up-sampling back to the full cloud with settings['upsampling'].  Returns confidences."""
    s = dict(DEFAULT_SETTINGS, **settings)
    batched = xyz.ndim == 3
    if not batched:
        xyz = xyz[None]
    if s["upsampling"] == "none":
        prepostprocess = False
    with torch.no_grad():
        inp = torch.from_numpy(xyz.astype(np.float32))
        if prepostprocess:
            ids = sample_points(xyz.shape[1], s["n_points"], consistent=True)
            sub = inp[:, ids, :]
            logits = forward(sd, s, sub)
            conf = torch.softmax(logits, dim=-2).unsqueeze(3)
            out = upsample(s["upsampling"], conf, sub[:, :, :3], inp[:, :, :3]).squeeze(-1).numpy()
        else:
            out = torch.softmax(forward(sd, s, inp), dim=-2).numpy()
    return out if batched else out[0]


def dice_loss(logits: torch.Tensor, labels: torch.Tensor, alpha=0.5, gamma=1.0, eps=1e-7) -> torch.Tensor:
    """FocalTverskyLoss(alpha=.5, gamma=1, neglect_background) = the trainer's default "dice"
    (losses.py:59-87, trainer.py:245-269) — used only to drive backward in gradient parity tests."""
    C = logits.size(-2)
    y_true = torch.eye(C, device=labels.device)[labels].transpose(-1, -2).permute(1, 0, 2).flatten(1)[1:]
    y_pred = F.softmax(logits, dim=-2).permute(1, 0, 2).flatten(1)[1:]
    tp = torch.sum(y_true * y_pred, dim=1)
    fn = torch.sum(y_true * (1 - y_pred), dim=1)
    fp = torch.sum((1 - y_true) * y_pred, dim=1)
    ti = (tp + eps) / (tp + alpha * fn + (1 - alpha) * fp + eps)
    return ((1 - ti) ** gamma).mean()


def grad_parity(got: "Dict[str, torch.Tensor]", ref: "Dict[str, torch.Tensor]", outliers: float = 0.01):
    """Worst relative gradient error over parameters: per tensor max|got-ref| / max|ref|, after setting aside
    the ``outliers`` fraction (at least one element) of largest deviations, which must still stay below 10 %.

    Why outliers are set aside.  The network is piecewise linear (ReLU / LeakyReLU).  Every forward holds a few
    pre-activations within fp32 round-off of zero (measured: min|u|/max|u| ~ 1e-8 per layer), and any two correct
    fp32 evaluations — different summation order, atomics — may put such an element on different sides of the
    kink.  That changes ONE (row, channel) term of the per-channel reductions by O(1) and shows up as an isolated
    deviation of 1e-3..1e-2 in one element of a bias / BatchNorm gradient; everything else agrees to ~1e-6.
    A genuine defect moves whole tensors, not one element in a hundred.

    A conv bias that feeds a train-mode BatchNorm (every ``*.conv.bias`` with a ``batch_norm`` sibling) has a
    mathematically ZERO gradient — the batch mean cancels it — so both sides hold only summation round-off
    there; those tensors are compared against the largest gradient in the network instead of against
    themselves.  Returns (worst_rel, name)."""
    scale = max(float(g.abs().max()) for g in ref.values())
    worst, worst_name = 0.0, ""
    for name, g in ref.items():
        mine = got[name]
        if mine is None:              # "no gradient" stands for an exactly zero gradient
            mine = torch.zeros_like(g)
        diff = (mine.detach().cpu().float().reshape(-1) - g.reshape(-1)).abs()
        cancelled = name == "fc_start.bias" or (          # fc_start feeds bn_start (modules.py:565-566)
            name.endswith(".conv.bias") and (name[:-len("conv.bias")] + "batch_norm.weight") in ref)
        denom = scale if cancelled else max(float(g.abs().max()), 1e-30)
        rel = torch.sort(diff / denom, descending=True).values
        n_out = min(max(1, int(np.ceil(outliers * rel.numel()))), rel.numel() - 1) if outliers > 0 else 0
        if n_out and float(rel[0]) >= 0.1:
            return float(rel[0]), name + " (outlier above 10 %)"
        r = float(rel[n_out]) if rel.numel() > n_out else 0.0
        if r > worst:
            worst, worst_name = r, name
    return worst, worst_name


def grad_parity_l2(got: "Dict[str, torch.Tensor]", ref: "Dict[str, torch.Tensor]"):
    """Worst per-tensor relative L2 error ||got - ref|| / ||ref|| (dead conv biases excluded).  Used for clouds
    beyond the golden sizes, where one flipped ReLU branch (see grad_parity) moves a whole channel row of the small
    encoding-MLP gradients — more than 1 % of their entries — while the tensor as a whole stays within ~1e-4..1e-3."""
    worst, worst_name = 0.0, ""
    for name, g in ref.items():
        dead = name == "fc_start.bias" or (
            name.endswith(".conv.bias") and (name[:-len("conv.bias")] + "batch_norm.weight") in ref)
        if dead or got[name] is None:
            continue
        r = float((got[name].detach().cpu().double() - g.double()).norm() / max(float(g.double().norm()), 1e-30))
        if r > worst:
            worst, worst_name = r, name
    return worst, worst_name


def grad_parity_fraction(got: "Dict[str, torch.Tensor]", ref: "Dict[str, torch.Tensor]", tol: float = 1e-4) -> float:
    """Fraction of all gradient entries (dead conv biases excluded) within ``tol`` * max|ref tensor| of the reference."""
    ok = total = 0
    for name, g in ref.items():
        dead = name == "fc_start.bias" or (
            name.endswith(".conv.bias") and (name[:-len("conv.bias")] + "batch_norm.weight") in ref)
        if dead or got[name] is None:
            continue
        d = (got[name].detach().cpu().double() - g.double()).abs()
        ok += int((d <= tol * max(float(g.abs().max()), 1e-30)).sum())
        total += d.numel()
    return ok / max(total, 1)


def grad_fixture_view(g: torch.Tensor, limit: int = 1024) -> torch.Tensor:
    """Golden fixtures keep gradients whole up to ``limit`` elements and a strided sample of the
    flattened tensor beyond that (keeps tests/golden small); tests view both sides through this."""
    if g is None:
        return None
    flat = g.detach().reshape(-1)
    if flat.numel() <= limit:
        return flat
    stride = -(-flat.numel() // limit)
    return flat[::stride]
