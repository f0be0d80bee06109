"""Execution engine behind ``modules.RandLANet.forward`` (reference: randlanet/utils/modules.py:542-611).

All tensors are POINT-MAJOR: features (B, N, C), neighbourhoods (B, N, K, C).  A point's channels are
contiguous, so a neighbour gather is one contiguous C-float read and "random down-sampling"
(modules.py:586-589: the first N/4^l points of the permuted cloud) is a prefix view per cloud.

Two execution paths share this file:

* ``forward_kernels``  — inference (no autograd): hand-written sm_100a kernels through the C ABI for
  every stage (KNN, fused LocSE + attentive pooling per LFA block, pointwise MLPs, 1-NN up-sampling).
* ``forward_autograd`` — training / whenever gradients are required: the exact CUDA KNN plus
  differentiable tensor ops for the rest (being replaced stage by stage by autograd.Functions over
  backward kernels).

Neither path runs on the CPU; ``ops`` raises if a tensor is not on a CUDA device.
"""
import os
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

from . import ops


# ----------------------------------------------------------------------------------------- helpers
def conv_weight_2d(smlp) -> torch.Tensor:
    """(C_out, C_in) view of a SharedMLP's 1x1 conv weight.  Conv2d stores (C_out,C_in,1,1);
    ConvTranspose2d stores (C_in,C_out,1,1) (checkpoint schema)."""
    w = smlp.conv.weight
    if isinstance(smlp.conv, torch.nn.ConvTranspose2d):
        return w.view(w.shape[0], w.shape[1]).t()
    return w.view(w.shape[0], w.shape[1])


def _activation(y: torch.Tensor, act) -> torch.Tensor:
    if act is None:
        return y
    if isinstance(act, torch.nn.ReLU):
        return F.relu(y)
    if isinstance(act, torch.nn.LeakyReLU):
        return F.leaky_relu(y, act.negative_slope)
    return act(y)


def batch_norm_lastdim(bn: torch.nn.BatchNorm2d, y: torch.Tensor) -> torch.Tensor:
    """BatchNorm2d semantics (modules.py:86-90) for a channel-last tensor: statistics over every
    leading position (B*N or B*N*K), biased variance for normalisation, running stats updated with
    momentum 0.99 and the unbiased variance in training mode."""
    c = y.shape[-1]
    out = F.batch_norm(y.reshape(-1, c), bn.running_mean, bn.running_var, bn.weight, bn.bias,
                       bn.training, bn.momentum, bn.eps)
    if bn.training and bn.track_running_stats:
        bn.num_batches_tracked.add_(1)
    return out.view_as(y)


_SIDE_STREAMS = {}
OVERLAP_WEIGHT_GRADS = os.environ.get("R3D_NO_OVERLAP") is None


SERIALIZE_FORKS = False      # bench.py's instrumented pass: every fork runs inline, so per-kernel events time one kernel


class _fork:
    """Runs the enclosed launches on a side stream of the current device, ordered after everything queued on the
    current stream so far; ``join()`` makes the current stream wait for them.  The per-point layers of a small cloud
    are latency-bound kernels on a handful of SMs each: the weight-gradient row reduction and the input-gradient
    GEMM of a layer are independent and overlap almost perfectly (also inside a CUDA-graph capture, where the fork
    and join become graph edges)."""

    def __init__(self, device, lane: int = 0):
        self.cur = torch.cuda.current_stream(device)
        # lane 0: short forks joined within the same function; lane 1: forks that stay open across several launches
        # of the main stream (a shared stream would serialise the short forks behind the long one)
        key = (device.index if device.index is not None else torch.cuda.current_device(), lane)
        if key not in _SIDE_STREAMS:
            _SIDE_STREAMS[key] = torch.cuda.Stream(device)
            ops.SIDE_STREAMS.add(_SIDE_STREAMS[key].cuda_stream)
        self.side = _SIDE_STREAMS[key]
        self.ctx = None

    def __enter__(self):
        self.inline = SERIALIZE_FORKS
        if self.inline:
            return self
        self.side.wait_stream(self.cur)
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if not self.inline:
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if not self.inline:
            self.cur.wait_stream(self.side)


def _once(ctx, what: str) -> None:
    """The backward kernels accumulate into scratch that the forward zero-filled once: a second backward over the
    same graph (retain_graph=True, gradient penalties) would add onto stale sums.  Refuse it loudly."""
    if getattr(ctx, "_r3d_consumed", False):
        raise RuntimeError(f"{what}: backward was already run on this graph; the fused kernels accumulate into "
                           "forward-allocated scratch and cannot be differentiated twice (rerun the forward)")
    ctx._r3d_consumed = True


class _SharedMLPTrainFn(torch.autograd.Function):
    """SharedMLP with train-mode BatchNorm on dense rows x (M,Cin): the sm_100a per-point kernels forward
    (GEMM + batch statistics, normalise + activation) and backward (BatchNorm backward, dx GEMM, dW row-reduction)."""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, bn, act, slope):
        x = x.contiguous()
        cout, cin = w.shape
        M = x.shape[0]
        scratch = ops.zeros(4 * cout, torch.float64, x.device)   # forward stats | backward stats
        stats = scratch[:2 * cout]
        # operand scales of the tensor-core weight-gradient kernel: [0] max |x| (written by the forward GEMM when it
        # runs on the tensor cores), [1] max |dz| (written by the BatchNorm backward)
        ctx.absmax = ops.zeros((2,), torch.float32, x.device) if ops.pc_wgrad_ok(M, cout, cin) else None
        if ops.pc_gemm_ok(M, cin, cout):
            ctx.absmax_x_known = ctx.absmax is not None
            z = ops.pc_gemm(x, w.contiguous(), stats=stats, absmax_out=None if ctx.absmax is None else ctx.absmax[0:1])
            y, save = ops.bn_apply(z, stats, bn, bias, act, slope)
        else:
            ctx.absmax_x_known = False
            z, y, save = ops.pointwise_bn(x, w.contiguous(), stats, bn, bias, act, slope)
        ctx.act, ctx.slope = act, slope
        ctx.save_for_backward(x, w, z, save, beta, scratch)
        return y

    @staticmethod
    def backward(ctx, dy):
        _once(ctx, "SharedMLP")
        x, w, z, save, beta, scratch = ctx.saved_tensors
        cout, cin = w.shape
        M = x.shape[0]
        am = ctx.absmax
        dz, dgamma, dbeta = ops.bn_backward(dy, z, save, beta, ctx.act, ctx.slope, stats2=scratch[2 * cout:],
                                            absmax_out=None if am is None else am[1:2])

        def weight_grad():
            if am is None:
                return ops.rowreduce_gemm(dz, x)
            return ops.pc_wgrad(dz, x, am[1:2], am[0:1] if ctx.absmax_x_known else None)

        def input_grad():
            if ops.pc_gemm_ok(M, cout, cin):
                return ops.pc_gemm(dz, w.contiguous(), transposed=True)
            return ops.pointwise(dz.unsqueeze(0), w.contiguous()).squeeze(0)

        if OVERLAP_WEIGHT_GRADS and ctx.needs_input_grad[0]:
            with _fork(dz.device) as f:
                dw = weight_grad()
            dx = input_grad()
            f.join()
        else:
            dx = input_grad() if ctx.needs_input_grad[0] else None
            dw = weight_grad()
        # the conv bias cancels against the batch mean: its gradient is exactly zero and is reported as None (the
        # optimiser then leaves the parameter alone, which is what a zero gradient does)
        return dx, dw, None, dgamma, dbeta, None, None, None


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on dense rows (SharedMLP without BatchNorm: the class logits, modules.py:525-527)."""

    @staticmethod
    def forward(ctx, x, w, bias):
        x = x.contiguous()
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None          # the attentive-pooling score Linear has none (modules.py:234-237)
        return ops.pointwise(x.unsqueeze(0), w.contiguous(), None, bias.contiguous() if ctx.has_bias else None,
                             w_out_in=True).squeeze(0)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        if OVERLAP_WEIGHT_GRADS and ctx.needs_input_grad[0]:
            with _fork(dy.device) as f:               # parameter gradients next to the input-gradient GEMM
                dw, db = ops.rowreduce_gemm(dy, x), (dy.sum(dim=0) if ctx.has_bias else None)
            dx = ops.pointwise(dy.unsqueeze(0), w.contiguous()).squeeze(0)
            f.join()
            return dx, dw, db
        dx = ops.pointwise(dy.unsqueeze(0), w.contiguous()).squeeze(0) if ctx.needs_input_grad[0] else None
        return dx, ops.rowreduce_gemm(dy, x), (dy.sum(dim=0) if ctx.has_bias else None)


class _SharedMLPEvalFn(torch.autograd.Function):
    """SharedMLP with EVAL-mode BatchNorm (running statistics) under autograd, on the same per-point kernels as the
    training layer: GEMM, normalise + activation from statistics sums synthesised out of the running mean / variance,
    and for the backward the BatchNorm reduce pass (dgamma, dbeta) plus the dz pass on zero batch sums (dz = a du)."""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, bn, act, slope):
        x = x.contiguous()
        M = x.shape[0]
        z = ops.pointwise(x.unsqueeze(0), w.contiguous(), w_out_in=True).squeeze(0)
        mean = (bn.running_mean - bias).double()              # of z, which carries no conv bias
        stats = torch.cat((mean * M, (bn.running_var.double() + mean * mean) * M))
        y, save = ops.bn_apply(z, stats, bn, None, act, slope, track=False)
        ctx.act, ctx.slope = act, slope
        ctx.save_for_backward(x, w, z, save, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, z, save, beta = ctx.saved_tensors
        dz, dgamma, dbeta = ops.bn_backward_fixed(dy, z, save, beta, ctx.act, ctx.slope)
        dx = ops.pointwise(dz.unsqueeze(0), w.contiguous()).squeeze(0) if ctx.needs_input_grad[0] else None
        return dx, ops.rowreduce_gemm(dz, x), save[0] * dbeta, dgamma, dbeta, None, None, None


def _kernel_layer_ok(x: torch.Tensor, bn) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and (bn is None or bn.weight.shape[0] % 4 == 0)


def shared_mlp(smlp, x: torch.Tensor) -> torch.Tensor:
    """SharedMLP on a channel-last tensor (..., C_in) -> (..., C_out).  fp32 tensors on a CUDA device run the per-point
    kernels, forward and backward, with batch statistics (train mode) or the running ones (eval mode under autograd);
    the CPU host-logic tests and the fp64 arbiters of the GPU tests take the same math as differentiable tensor ops."""
    bn = smlp.batch_norm
    if USE_POINTWISE_KERNELS and _kernel_layer_ok(x, bn):
        w = conv_weight_2d(smlp)
        x2 = x.reshape(-1, x.shape[-1])
        if bn is None:
            y = _LinearFn.apply(x2, w, smlp.conv.bias)
            return _activation(y, smlp.activation).view(*x.shape[:-1], w.shape[0])
        act, slope = _act_of(smlp)
        fn = _SharedMLPTrainFn if bn.training else _SharedMLPEvalFn
        y = fn.apply(x2, w, smlp.conv.bias, bn.weight, bn.bias, bn, act, slope)
        return y.view(*x.shape[:-1], w.shape[0])
    y = F.linear(x, conv_weight_2d(smlp), smlp.conv.bias)
    if bn is not None:
        y = batch_norm_lastdim(bn, y)
    return _activation(y, smlp.activation)


# train-mode per-point layers on the sm_100a kernels (False: differentiable tensor ops, for A/B tests)
USE_POINTWISE_KERNELS = True


def gather_points(feat: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """feat (B,Ns,C), idx (B,Nq,K) -> (B,Nq,K,C): rows of ``feat`` at the neighbour indices."""
    B, Ns, C = feat.shape
    _, Nq, K = idx.shape
    flat = (idx.long() + (torch.arange(B, device=idx.device) * Ns).view(B, 1, 1)).reshape(-1)
    return feat.reshape(B * Ns, C).index_select(0, flat).view(B, Nq, K, C)


def relative_position_encoding(xyz: torch.Tensor, idx: torch.Tensor, dist: torch.Tensor) -> torch.Tensor:
    """(B,N,K,10) = [p_i, p_j, p_i - p_j, |p_i - p_j|] (modules.py:170-186)."""
    pj = gather_points(xyz, idx)
    pi = xyz.unsqueeze(2).expand_as(pj)
    return torch.cat((pi, pj, pi - pj, dist.unsqueeze(-1)), dim=-1)


def attentive_pooling(pool, x: torch.Tensor) -> torch.Tensor:
    """x (B,N,K,d) -> (B,N,n_out): softmax over K of Linear(x), weighted sum, SharedMLP
    (modules.py:246-253)."""
    scores = F.softmax(F.linear(x, pool.score_fn[0].weight), dim=2)
    return shared_mlp(pool.mlp, (scores * x).sum(dim=2))


def lfa_block(lfa, xyz: torch.Tensor, feat: torch.Tensor) -> torch.Tensor:
    """LocalFeatureAggregation (modules.py:298-325): xyz (B,N,3), feat (B,N,n_in) -> (B,N,2d)."""
    nn_ = ops.knn(xyz, xyz, lfa._n_neighbors, idx64=True, dist=True)
    idx, dist = nn_["idx64"], nn_["dist"]
    f = shared_mlp(lfa.mlp1, feat)
    r1 = shared_mlp(lfa.mlp_rpe1, relative_position_encoding(xyz, idx, dist))
    p1 = attentive_pooling(lfa.pool1, torch.cat((r1, gather_points(f, idx)), dim=-1))
    r2 = shared_mlp(lfa.mlp_rpe2, r1)                     # fed by r1, not by the raw encoding (modules.py:321)
    p2 = attentive_pooling(lfa.pool2, torch.cat((r2, gather_points(p1, idx)), dim=-1))
    return F.leaky_relu(shared_mlp(lfa.mlp2, p2) + shared_mlp(lfa.shortcut, feat), 0.01)


class _UpsampleFn(torch.autograd.Function):
    """ops.upsample (gather at the K nearest coarse points, weighted sum, optional concat with the decoder skip) with
    its scatter-add backward (ops.upsample_bwd): one launch each way."""

    @staticmethod
    def forward(ctx, approach, feat, idx, dist, skip):
        ctx.approach, ctx.n_coarse, ctx.n_feat = approach, feat.shape[1], feat.shape[2]
        ctx.save_for_backward(idx, dist)
        return ops.upsample(approach, feat, idx, dist, skip)

    @staticmethod
    def backward(ctx, dout):
        idx, dist = ctx.saved_tensors
        need_skip = ctx.needs_input_grad[4]
        dfeat, dskip = ops.upsample_bwd(ctx.approach, dout, idx, dist, ctx.n_coarse, ctx.n_feat, need_skip)
        return None, dfeat if ctx.needs_input_grad[1] else None, None, None, dskip


def upsample(approach: str, feat: torch.Tensor, xyz: torch.Tensor, xyz_up: torch.Tensor,
             idx: torch.Tensor = None, skip: torch.Tensor = None, channel_major: bool = False) -> torch.Tensor:
    """UpSampler (modules.py:343-456) on point-major features: feat (B,N1,F) -> (B,N2,F) [(B,N2,F+Fs) with the decoder
    skip (B,N2,Fs) concatenated, modules.py:600-602; (B,F,N2) with ``channel_major``].  ``idx``: the 1-NN indices
    (B,N2,1) when the caller has searched already.  One neighbour search and one gather launch; differentiable in
    ``feat`` and ``skip``."""
    if approach not in ops.UP_WEIGHTING:
        raise ValueError(f"Upsampling approach {approach} not understood!")
    dist = None
    if not (feat.is_cuda and feat.dtype == torch.float32):
        # fp64 arbiters and the CPU host-logic tests (forward() itself refuses non-CUDA inputs): the same math as
        # differentiable tensor ops, like shared_mlp
        if approach == "nni":
            if idx is None:
                idx = ops.knn(xyz, xyz_up, 1, idx64=True, dist=False)["idx64"]
            up = gather_points(feat, idx[:, :, :1]).squeeze(2)
        else:
            nn_ = ops.knn(xyz, xyz_up, 8, idx64=True, dist=True)
            w = (1.0 + 1e-7) / (nn_["dist"] ** ops.UP_WEIGHTING[approach][1] + 1e-7)
            w = w / w.sum(dim=-1, keepdim=True)
            up = (w.unsqueeze(-1) * gather_points(feat, nn_["idx64"])).sum(dim=2)
        up = up if skip is None else torch.cat((up, skip), dim=-1)
        return up.transpose(1, 2) if channel_major else up
    if approach == "nni":
        if idx is None:
            idx = ops.knn(xyz, xyz_up, 1, idx64=False, idx32=True, dist=False)["idx32"]
    else:
        # "nna" reaches the inverse-distance branch too (default argument, modules.py:372 / :435)
        nn_ = ops.knn(xyz, xyz_up, 8, idx64=False, idx32=True, dist=True)
        idx, dist = nn_["idx32"], nn_["dist"]
    if torch.is_grad_enabled() and (feat.requires_grad or (skip is not None and skip.requires_grad)):
        out = _UpsampleFn.apply(approach, feat, idx, dist, skip)
        return out.transpose(1, 2) if channel_major else out
    return ops.upsample(approach, feat, idx, dist, skip, channel_major)


# ------------------------------------------------------------- training path: fused LFA with autograd
class _LfaPoolFn(torch.autograd.Function):
    """One fused LocSE + attentive-pooling launch (ops.lfa_pool) with its mirrored backward kernel
    (ops.lfa_pool_bwd).  Differentiable inputs: feat, ws (d,d), and — as fp64 tensors — w1 (h,10), a1, c1 (h)
    [, w2 (h,h), a2, c2]; a*/c* are the per-channel affine maps that BatchNorm (+ conv bias) reduces to.

    Why fp64 edges.  The gradient of mlp_rpe1/2 is the sum of a direct term (a (.) sum du x^T, from here) and the
    BatchNorm-statistics term (from _BnFromMomentsFn); with un-centred inputs (absolute coordinates) the two are
    ~1e3 times larger than their sum.  Accumulating the sums in fp64 and letting autograd add the fp64 pieces keeps
    the result at fp32 accuracy (measured at N=16384: 8e-4 relative in fp32, 3e-6 in fp64); the kernels themselves
    run in fp32 on the float copies w1f, a1f, c1f [, w2f, a2f, c2f]."""

    @staticmethod
    def forward(ctx, stage, xyz, idx32, feat, ws, w1, a1, c1, w2, a2, c2, w1f, a1f, c1f, w2f, a2f, c2f):
        w2T = w2f.t().contiguous() if stage == 2 else None
        wsT = ws.t().contiguous()
        if ops.lfa_pool_tc_supported(ws.shape[0], idx32.shape[2], idx32.shape[0] * idx32.shape[1]):
            pooled = ops.lfa_pool_tc(stage, xyz, idx32, feat, w1f, a1f, c1f, w2f if stage == 2 else None,
                                     a2f if stage == 2 else None, c2f if stage == 2 else None, ws.contiguous())
        else:
            pooled = ops.lfa_pool(stage, xyz, idx32, feat, w1f, a1f, c1f, w2T, a2f if stage == 2 else None,
                                  c2f if stage == 2 else None, wsT)
        ctx.stage = stage
        ctx.save_for_backward(xyz, idx32, feat, ws, w1, a1, w2 if stage == 2 else None, a2 if stage == 2 else None,
                              w1f, a1f, c1f, w2f if stage == 2 else None, a2f if stage == 2 else None,
                              c2f if stage == 2 else None, w2T, wsT)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        xyz, idx32, feat, ws, w1, a1, w2, a2, w1f, a1f, c1f, w2f, a2f, c2f, w2T, wsT = ctx.saved_tensors
        stage = ctx.stage
        w2s = (w2f * a2f.unsqueeze(1)).contiguous() if stage == 2 else None
        dfeat, dws, g1, g2m, g2c = ops.lfa_pool_bwd(stage, xyz, idx32, feat, w1f, a1f, c1f, w2T, a2f, c2f, w2s, wsT,
                                                    ws.contiguous(), dpooled)
        gm = g1[:, :10]
        dw1, da1, dc1 = a1.unsqueeze(1) * gm, (w1 * gm).sum(dim=1), g1[:, 10]
        dw2 = da2 = dc2 = None
        if stage == 2:
            dw2, da2, dc2 = a2.unsqueeze(1) * g2m, (w2 * g2m).sum(dim=1), g2c[:, 10]
        return (None, None, None, dfeat, dws, dw1, da1, dc1, dw2, da2, dc2) + (None,) * 6


class _LfaPool1TrainFn(torch.autograd.Function):
    """Stage 1 of an LFA block in TRAINING mode: rpe -> r1 = relu(BN_batch(W1 rpe)) -> PFA gather of f -> attentive
    pooling, with mlp_rpe1's BatchNorm expressed through the moments of the encoding (ops.bn_from_moments).
    Differentiable inputs: feat, ws, and mlp_rpe1's parameters w1 (conv weight), gamma1, beta1.  The gradient of
    mlp_rpe1 collects contributions from BOTH halves of the block (r1 also feeds mlp_rpe2): the kernels of both
    accumulate du1 (x) [rpe, 1] into the shared fp64 buffer ``g1`` (stage 2's backward always runs first: its
    input p1 is computed from this function's output), and this backward turns the total into (dW1, dgamma1,
    dbeta1) with one kernel (ops.lfa_rpe1_grads).  The fp64 accumulation matters: the direct term and the
    BatchNorm-statistics term are ~1e3 times larger than their sum with un-centred inputs (measured at N=16384:
    8e-4 relative in fp32, 3e-6 in fp64)."""

    @staticmethod
    def forward(ctx, xyz, idx32, feat, ws, w1, gamma1, beta1, w2, gamma2, beta2, w1f, a1f, c1f, m, save1, g1, count,
                shared):
        wsT = shared["wT"][0] if "wT" in shared else ws.t().contiguous()
        if ops.lfa_pool_tc_supported(ws.shape[0], idx32.shape[2], idx32.shape[0] * idx32.shape[1]):
            pooled = ops.lfa_pool_tc(1, xyz, idx32, feat, w1f, a1f, c1f, None, None, None, ws.contiguous())
        else:
            pooled = ops.lfa_pool(1, xyz, idx32, feat, w1f, a1f, c1f, None, None, None, wsT)
        ctx.count = count
        ctx.w1_shape = w1.shape
        ctx.shared = shared
        ctx.save_for_backward(xyz, idx32, feat, ws, wsT, gamma1, w1f, a1f, c1f, m, save1, g1)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        _once(ctx, "LocalFeatureAggregation stage 1")
        xyz, idx32, feat, ws, wsT, gamma1, w1f, a1f, c1f, m, save1, g1 = ctx.saved_tensors
        dfeat, dws, _, _, _ = ops.lfa_pool_bwd(1, xyz, idx32, feat, w1f, a1f, c1f, None, None, None, None, wsT,
                                               ws.contiguous(), dpooled, g1_acc=g1)
        # stage 2's second BatchNorm pass (side stream, see _LfaPool2TrainFn.backward) has been adding to g1 too
        dw2, dgamma2, dbeta2, fork, keep_alive = ctx.shared.pop("stage2", (None, None, None, None, None))
        if fork is not None:
            fork.join()
        del keep_alive
        dw1, dgamma1, dbeta1 = ops.lfa_rpe1_grads(w1f, m[10], m, ctx.count, gamma1.detach().contiguous(), save1, g1)
        return (None, None, dfeat, dws, dw1.view(ctx.w1_shape), dgamma1, dbeta1, dw2, dgamma2, dbeta2) + (None,) * 8


class _LfaPool2TrainFn(torch.autograd.Function):
    """Stage 2 of an LFA block in TRAINING mode: r2 = relu(BN_batch(W2 r1)), PFA gather of p1, attentive pooling.
    Forward = the same fused kernel as _LfaPoolFn; backward = the standard two BatchNorm passes
    (ops.lfa_pool2_bwd_train: everything down to du2 and the batch sums; ops.lfa_bn2_coeffs: the per-channel
    coefficients; ops.lfa_bn2_bwd: dz2, dW2, dr1, du1, and du1 (x) [rpe, 1] added to the block's shared ``g1``
    accumulator, see _LfaPool1TrainFn).  Differentiable inputs: feat (= p1), ws, w2 (conv weight), gamma2, beta2.
    a2f/c2f/save2 are this step's batch statistics (constants here: their dependence on w2 and r1 IS the BatchNorm
    backward that pass 2 carries out)."""

    @staticmethod
    def forward(ctx, xyz, idx32, feat, ws, w2, w1f, a1f, c1f, a2f, c2f, save2, g1, shared):
        h = w1f.shape[0]
        w2f = w2.detach().view(h, h)
        if "wT" in shared:
            _, wsT, w2T = shared["wT"]
        else:
            w2T = w2f.t().contiguous()
            wsT = ws.t().contiguous()
        if ops.lfa_pool_tc_supported(ws.shape[0], idx32.shape[2], idx32.shape[0] * idx32.shape[1]):
            ctx.tc_cache = {}           # d = 256: the r2 rows the forward materialised, reused by pass 1 of the backward
            pooled = ops.lfa_pool_tc(2, xyz, idx32, feat, w1f, a1f, c1f, w2f.contiguous(), a2f, c2f, ws.contiguous(),
                                     cache=ctx.tc_cache)
        else:
            ctx.tc_cache = None
            pooled = ops.lfa_pool(2, xyz, idx32, feat, w1f, a1f, c1f, w2T, a2f, c2f, wsT)
        ctx.w2_shape = w2.shape
        ctx.shared = shared
        ctx.save_for_backward(xyz, idx32, feat, ws, w1f, a1f, c1f, w2f, w2T, wsT, a2f, c2f, save2, g1)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        _once(ctx, "LocalFeatureAggregation stage 2")
        xyz, idx32, feat, ws, w1f, a1f, c1f, w2f, w2T, wsT, a2f, c2f, save2, g1 = ctx.saved_tensors
        h = w1f.shape[0]
        dfeat, dws, du2, sums = ops.lfa_pool2_bwd_train(xyz, idx32, feat, w1f, a1f, c1f, w2T, a2f, c2f, wsT,
                                                        ws.contiguous(), dpooled, w_rpe2=w2f.contiguous(),
                                                        cache=ctx.tc_cache)
        # The second pass only produces parameter gradients (dW2, and mlp_rpe1's through g1): it runs on the side
        # stream, overlapping pool1.mlp's and stage 1's backward, and is joined in _LfaPool1TrainFn.backward, which
        # also hands (dW2, dgamma2, dbeta2) to autograd (w2 is a pass-through input there).
        with _fork(dfeat.device, lane=1) as fork:
            bn2, dgamma2, dbeta2 = ops.lfa_bn2_coeffs(sums, a2f, c2f, save2, float(idx32.numel()))
            _, dw2 = ops.lfa_bn2_bwd(xyz, idx32, w1f, a1f, c1f, du2, w2T, w2f, bn2, h, g1_acc=g1)
            dw2 = dw2.float().view(ctx.w2_shape)
        # du2 and w2T were allocated on the main stream and are read by the side stream: everything the forked launches
        # read stays referenced until the join (a freed block returns to the main stream's pool -- autograd drops the
        # saved tensors when this function returns -- and could be handed out again while pass 2 still reads it)
        ctx.shared["stage2"] = (dw2, dgamma2, dbeta2, fork, (du2, sums, bn2, w2T, a2f, c2f, save2))
        return (None, None, dfeat, dws) + (None,) * 9


class _R1MomentsFn(torch.autograd.Function):
    """(sum r1 (h), sum r1 r1^T (h,h)) in fp64 over all (point, neighbour) rows, r1 = relu(a1 (W1 rpe) + c1):
    the statistics mlp_rpe2's train-mode BatchNorm needs.  w1, a1, c1: fp64 autograd edges, w1f.. their fp32 copies."""

    @staticmethod
    def forward(ctx, xyz, idx32, w1, a1, c1, w1f, a1f, c1f):
        m_r1, s_r1 = ops.lfa_moments(1, xyz, idx32, 2 * w1f.shape[0], w1f, a1f, c1f)
        ctx.save_for_backward(xyz, idx32, w1, a1, w1f, a1f, c1f)
        return s_r1[:, 10].clone(), m_r1

    @staticmethod
    def backward(ctx, g_sum, g_m):
        xyz, idx32, w1, a1, w1f, a1f, c1f = ctx.saved_tensors
        gsym = (g_m + g_m.t()).float().contiguous()
        g1 = ops.lfa_moments(2, xyz, idx32, 2 * w1f.shape[0], w1f, a1f, c1f, gsym=gsym,
                             gsum=g_sum.float().contiguous())
        gm = g1[:, :10]
        return None, None, a1.unsqueeze(1) * gm, (w1 * gm).sum(dim=1), g1[:, 10], None, None, None


class _BnFromMomentsFn(torch.autograd.Function):
    """Train-mode BatchNorm of y = W x + b expressed through the input moments (sums s, second-moment sums m over
    ``count`` rows, fp64): mean_y = W mu + b, var_y = diag(W Cov W^T) (biased, as BatchNorm normalises;
    modules.py:86-90).  Returns the per-channel affine (a, c) with bn(y) = a (W x) + c as fp64 tensors (see
    _LfaPoolFn) and updates the running statistics like BatchNorm2d (momentum 0.99, unbiased running variance).
    ``w`` is the fp64 autograd edge of the conv weight, ``wf`` its fp32 values.  One small kernel forward, one or two
    backward (ops.bn_from_moments*)."""

    @staticmethod
    def forward(ctx, w, wf, gamma, beta, s, m, bn, bias, count):
        a, c, save = ops.bn_from_moments(wf, s, m, count, bn, bias)
        ctx.count = count
        ctx.save_for_backward(wf, gamma, s, m, save)
        return a.double(), c.double()

    @staticmethod
    def backward(ctx, ga, gc):
        wf, gamma, s, m, save = ctx.saved_tensors
        need = ctx.needs_input_grad[4] or ctx.needs_input_grad[5]
        dw, dgamma, dbeta, dm, ds = ops.bn_from_moments_bwd(wf, s, m, ctx.count, gamma.contiguous(), save, ga, gc, need)
        return dw, None, dgamma, dbeta, ds, dm, None, None, None


def _eval_affine(smlp):
    bn = smlp.batch_norm
    a = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    return a, bn.bias + (smlp.conv.bias - bn.running_mean) * a


def lfa_transposed_weights(lfa):
    """(pool1 score weight^T, pool2 score weight^T, mlp_rpe2 weight^T): the [in][out] layouts the fused kernels
    stream.  Plain values (no autograd): the functions differentiate with respect to the parameters themselves."""
    h = lfa.mlp_rpe2.conv.weight.shape[0]
    with torch.no_grad():
        return (lfa.pool1.score_fn[0].weight.t().contiguous(), lfa.pool2.score_fn[0].weight.t().contiguous(),
                lfa.mlp_rpe2.conv.weight.view(h, h).t().contiguous())


def lfa_block_fused(lfa, xyz: torch.Tensor, feat: torch.Tensor, idx: torch.Tensor = None, wT=None) -> torch.Tensor:
    """LocalFeatureAggregation (modules.py:298-325) with gradients: KNN and the two fused LocSE + pooling
    halves run on the sm_100a kernels (forward AND backward); the per-point layers around them are
    differentiable tensor ops.  Works in train mode (batch statistics) and eval mode (running statistics)."""
    K = lfa._n_neighbors
    # the residual branch depends on the block input only: it runs on a side stream next to the whole main chain
    # (autograd replays it on that stream too, so its backward overlaps mlp2's)
    sc_fork = None
    if OVERLAP_WEIGHT_GRADS and feat.is_cuda:
        with _fork(feat.device, lane=2) as sc_fork:
            sc = shared_mlp(lfa.shortcut, feat)
    if idx is None:
        idx = ops.knn(xyz, xyz, K, idx64=False, idx32=True, dist=False)["idx32"]
    f = shared_mlp(lfa.mlp1, feat)
    r1m, r2m = lfa.mlp_rpe1, lfa.mlp_rpe2
    w1f = r1m.conv.weight.detach().view(-1, 10)
    h = w1f.shape[0]
    w2f = r2m.conv.weight.detach().view(h, h)
    d = 2 * h
    ws1, ws2 = lfa.pool1.score_fn[0].weight, lfa.pool2.score_fn[0].weight
    if r1m.batch_norm.training:
        # batch statistics of mlp_rpe1 / mlp_rpe2 from the moments of their inputs (no (B,N,K,h) tensor is ever
        # materialised); everything between the parameters and the pooled features is inside the two functions
        bn1, bn2 = r1m.batch_norm, r2m.batch_norm
        count = float(xyz.shape[0] * xyz.shape[1] * K)
        g1 = ops.zeros((h, 16), torch.float64, xyz.device)       # mlp_rpe1's gradient accumulator, both halves
        shared = {} if wT is None else {"wT": wT}
        # the statistics kernels of stage 2 depend on the encoding only, not on pooled1: they run on the side stream
        # next to stage 1's pooling kernel and pool1.mlp (every kernel here fills a fraction of the SMs at this size)
        with torch.no_grad():
            m = ops.lfa_moments(0, xyz, idx, d)                   # (16,16) fp64, no parameters involved
            a1f, c1f, save1 = ops.bn_from_moments(w1f, m[10], m, count, bn1, r1m.conv.bias)
            with _fork(xyz.device, lane=1) as fork:
                m_r1, s_r1 = ops.lfa_moments(1, xyz, idx, d, w1f, a1f, c1f)
                a2f, c2f, save2 = ops.bn_from_moments(w2f, s_r1[:, 10], m_r1, count, bn2, r2m.conv.bias)
        pooled1 = _LfaPool1TrainFn.apply(xyz, idx, f, ws1, r1m.conv.weight, bn1.weight, bn1.bias, r2m.conv.weight,
                                         bn2.weight, bn2.bias, w1f, a1f, c1f, m, save1, g1, count, shared)
        p1 = shared_mlp(lfa.pool1.mlp, pooled1)
        fork.join()
        pooled2 = _LfaPool2TrainFn.apply(xyz, idx, p1, ws2, r2m.conv.weight, w1f, a1f, c1f, a2f, c2f, save2, g1,
                                         shared)
    else:
        w1 = r1m.conv.weight.view(h, 10).double()        # fp64 autograd edges (see _LfaPoolFn)
        w2 = r2m.conv.weight.view(h, h).double()
        a1, c1 = (t.double() for t in _eval_affine(r1m))
        a1f, c1f = a1.detach().float(), c1.detach().float()
        pooled1 = _LfaPoolFn.apply(1, xyz, idx, f, ws1, w1, a1, c1, None, None, None, w1f, a1f, c1f, None, None, None)
        p1 = shared_mlp(lfa.pool1.mlp, pooled1)
        a2, c2 = (t.double() for t in _eval_affine(r2m))
        a2f, c2f = a2.detach().float(), c2.detach().float()
        pooled2 = _LfaPoolFn.apply(2, xyz, idx, p1, ws2, w1, a1, c1, w2, a2, c2, w1f, a1f, c1f, w2f, a2f, c2f)
    p2 = shared_mlp(lfa.pool2.mlp, pooled2)
    if sc_fork is not None:
        sc_fork.join()
    else:
        sc = shared_mlp(lfa.shortcut, feat)
    return residual_lrelu(shared_mlp(lfa.mlp2, p2), sc, 0.01)


class _SkipAndPrefixFn(torch.autograd.Function):
    """An encoder level's output is used twice: whole, as the decoder's skip connection, and its first n points per cloud
    (random down-sampling = a prefix of the permuted cloud, modules.py:583) as the next level's input.  Plain autograd
    turns that into a zero-filled full-size tensor, a copy of the prefix gradient into it and a full-size sum; here the
    prefix gradient is added onto the skip gradient's first n rows (a quarter of the tensor, one launch)."""

    @staticmethod
    def forward(ctx, out, n):
        ctx.n, ctx.shape = n, out.shape
        return out.view_as(out), out[:, :n].contiguous()

    @staticmethod
    def backward(ctx, dskip, dprefix):
        if dprefix is None:
            return dskip, None
        if dskip is None:
            total = torch.zeros(ctx.shape, dtype=dprefix.dtype, device=dprefix.device)
        else:
            # In this network dskip is the fresh tensor the decoder's up-sampling backward allocated (ops.upsample_bwd) and
            # this node is its only consumer, so the prefix gradient is accumulated into it in place; anything that is a
            # view of other storage, or not dense, is copied first.
            total = dskip if (dskip.is_contiguous() and dskip._base is None) else dskip.clone(memory_format=torch.contiguous_format)
        total[:, :ctx.n].add_(dprefix)
        return total, None


class _AddLReluFn(torch.autograd.Function):
    """LeakyReLU(a + b) — the residual sum that closes an LFA block (modules.py:325) — as one launch each way
    (ops.add_lrelu / add_lrelu_bwd): the backward's single output is the gradient of both summands."""

    @staticmethod
    def forward(ctx, a, b, slope):
        y = ops.add_lrelu(a, b, slope)
        ctx.slope = slope
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        d = ops.add_lrelu_bwd(dy, y, ctx.slope)
        return d, d, None


def residual_lrelu(a: torch.Tensor, b: torch.Tensor, slope: float = 0.01) -> torch.Tensor:
    if a.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape:
        return _AddLReluFn.apply(a, b, slope)
    return F.leaky_relu(a + b, slope)


# ------------------------------------------------- row-form LFA block: any n_neighbors, any layer size
class _GatherConcatFn(torch.autograd.Function):
    """X = [r ; feat at the neighbours] over the (B*N*K) rows (ops.lfa_gather_concat) and its scatter-add backward."""

    @staticmethod
    def forward(ctx, r, feat, idx32):
        ctx.save_for_backward(idx32)
        return ops.lfa_gather_concat(r, feat, idx32)

    @staticmethod
    def backward(ctx, dout):
        (idx32,) = ctx.saved_tensors
        dr, dfeat = ops.lfa_gather_concat_bwd(dout, idx32, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dr, dfeat, None


class _AttnPoolFn(torch.autograd.Function):
    """pooled = sum_k softmax_k(S) X over the K rows of every point (ops.lfa_attn_pool / lfa_attn_pool_bwd)."""

    @staticmethod
    def forward(ctx, S, X, K):
        ctx.K = K
        ctx.save_for_backward(S, X)
        return ops.lfa_attn_pool(S, X, K)

    @staticmethod
    def backward(ctx, dpooled):
        S, X = ctx.saved_tensors
        dS, dX = ops.lfa_attn_pool_bwd(S, X, dpooled, ctx.K)
        return dS, dX, None


def lfa_block_rows(lfa, xyz: torch.Tensor, feat: torch.Tensor, idx: torch.Tensor = None, wT=None) -> torch.Tensor:
    """LocalFeatureAggregation (modules.py:298-325) for settings outside the fused kernels' template lists (any
    n_neighbors >= 1, any layer size): the neighbourhood rows (B*N*K, C) are materialised, mlp_rpe1/2 and the score
    Linear run as per-point layers over them (train-mode BatchNorm statistics over all B*N*K rows, modules.py:86-90),
    gather + concat and softmax-pooling are the row kernels of csrc/lfa_rows.cu.  Differentiable; same result as
    `lfa_block_fused` to fp32 round-off where both apply (tests/test_lfa_rows_gpu.py)."""
    K = lfa._n_neighbors
    B, N, _ = xyz.shape
    if idx is None:
        idx = ops.knn(xyz, xyz, K, idx64=False, idx32=True, dist=False)["idx32"]
    sc = shared_mlp(lfa.shortcut, feat)
    f = shared_mlp(lfa.mlp1, feat)
    r1 = shared_mlp(lfa.mlp_rpe1, ops.lfa_rpe_rows(xyz, idx))
    X1 = _GatherConcatFn.apply(r1, f, idx)
    S1 = _LinearFn.apply(X1, lfa.pool1.score_fn[0].weight, None)
    p1 = shared_mlp(lfa.pool1.mlp, _AttnPoolFn.apply(S1, X1, K).view(B, N, -1))
    r2 = shared_mlp(lfa.mlp_rpe2, r1)                      # fed by r1, not by the raw encoding (modules.py:321)
    X2 = _GatherConcatFn.apply(r2, p1, idx)
    S2 = _LinearFn.apply(X2, lfa.pool2.score_fn[0].weight, None)
    p2 = shared_mlp(lfa.pool2.mlp, _AttnPoolFn.apply(S2, X2, K).view(B, N, -1))
    return residual_lrelu(shared_mlp(lfa.mlp2, p2), sc, 0.01)


def lfa_block_auto(lfa, xyz: torch.Tensor, feat: torch.Tensor, idx: torch.Tensor = None, wT=None) -> torch.Tensor:
    """The fused kernels where they are instantiated (d in ops.LFA_FUSED_WIDTHS, K in ops.LFA_FUSED_NEIGHBORS), the
    row form elsewhere."""
    d = 2 * lfa.mlp_rpe1.conv.weight.shape[0]
    if ops.lfa_fused_supported(d, lfa._n_neighbors):
        return lfa_block_fused(lfa, xyz, feat, idx, wT)
    return lfa_block_rows(lfa, xyz, feat, idx, wT)


# LFA implementation used by forward_autograd: the fused kernels (row form for shapes they are not built for), or
# (tests / debugging) the plain tensor-op composition `lfa_block`
LFA_IMPL = lfa_block_auto


# ------------------------------------------------------------------------------------ full forward
def _device_permutation(permutation, device) -> torch.Tensor:
    """The host-drawn permutation (modules.py:571) as an int64 device tensor; a device tensor passes through
    (CUDA-graph replays update a static permutation buffer instead of copying inside the graph)."""
    if isinstance(permutation, torch.Tensor):
        return permutation.to(device=device, dtype=torch.int64)
    return torch.from_numpy(np.ascontiguousarray(permutation)).to(device, non_blocking=True)


def forward_autograd(net, inp: torch.Tensor, permutation) -> torch.Tensor:
    s = net.settings
    dec, L = s.decimation, len(s.layer_sizes)
    B, N, _ = inp.shape
    perm = _device_permutation(permutation, inp.device)
    if inp.is_cuda:
        ops.ZEROS.begin()          # one zero-filled scratch block for this forward/backward (ops._ZeroPool)

    # The reference permutes coordinates and features after fc_start (modules.py:565-573).  fc_start + BatchNorm act
    # per point (the batch statistics are sums over points), so permuting the INPUT once gives the same tensors with
    # one gather instead of two and no scatter at the very end of the backward.
    inp = inp.float().index_select(1, perm)
    bn0 = net.bn_start[0]
    if USE_POINTWISE_KERNELS and _kernel_layer_ok(inp, bn0):
        fn0 = _SharedMLPTrainFn if bn0.training else _SharedMLPEvalFn
        feat = fn0.apply(inp.reshape(B * N, -1), net.fc_start.weight, net.fc_start.bias, bn0.weight,
                         bn0.bias, bn0, "lrelu", float(net.bn_start[1].negative_slope)).view(B, N, -1)
    else:
        feat = F.linear(inp, net.fc_start.weight, net.fc_start.bias)
        feat = F.leaky_relu(batch_norm_lastdim(bn0, feat), net.bn_start[1].negative_slope)
    xyz = inp[..., :3].contiguous()

    # Every neighbour search depends on the coordinates only.  With the fused kernels the searches of the
    # down-sampled levels and of the decoder's 1-NN up-sampling run on a side stream while level 0 is processed
    # (each is a few-microsecond kernel on a fraction of the SMs; ~75 us of a 3 ms step on the main chain otherwise).
    pre_fork, enc_idx, dec_idx, enc_wT = None, {}, {}, {}
    if LFA_IMPL in (lfa_block_fused, lfa_block_auto) and OVERLAP_WEIGHT_GRADS and xyz.is_cuda and L > 1:
        with torch.no_grad(), _fork(xyz.device, lane=3) as pre_fork:
            if net.training:
                for lvl in range(1, L):              # weight transposes of the later levels: off the main chain too
                    enc_wT[lvl] = lfa_transposed_weights(net.encoder[lvl])
            sizes = [N]
            for lvl in range(L):
                sizes.append(sizes[-1] // dec)       # points kept after level lvl (floor, modules.py:583)
            for lvl in range(1, L):
                n_k = sizes[lvl]
                K = net.encoder[lvl]._n_neighbors
                enc_idx[lvl] = ops.knn(xyz[:, :n_k], xyz[:, :n_k], K, idx64=False, idx32=True, dist=False)["idx32"]
            for lvl in range(L):                     # decoder stage lvl: sizes[L - lvl] -> sizes[L - lvl - 1] points
                dec_idx[lvl] = ops.knn(xyz[:, :sizes[L - lvl]], xyz[:, :sizes[L - lvl - 1]], 1, idx64=False, idx32=True,
                                       dist=False)["idx32"]

    skips: List[torch.Tensor] = []
    n_l = N
    cur = feat
    for lvl, lfa in enumerate(net.encoder):
        if lvl == 1 and pre_fork is not None:
            pre_fork.join()
        out = (LFA_IMPL(lfa, xyz[:, :n_l], cur, enc_idx[lvl], enc_wT.get(lvl)) if lvl in enc_idx
               else LFA_IMPL(lfa, xyz[:, :n_l], cur))
        n_l //= dec
        # random down-sampling = a prefix of the permuted cloud (modules.py:583); one dense copy here instead of one in
        # every consumer of the strided view (mlp1 and shortcut of the next block, forward and backward)
        if out.is_cuda and out.requires_grad:
            out, cur = _SkipAndPrefixFn.apply(out, n_l)
        else:
            cur = out[:, :n_l].contiguous() if out.is_cuda else out[:, :n_l]
        skips.append(out)
    cur = shared_mlp(net.mlp, cur)
    for lvl, stage in enumerate(net.decoder):
        n_up = skips[-1].shape[1]          # N // dec^(l-1): the encoder level this stage returns to
        # 1-NN up-sampling + skip concat (modules.py:596-602): one gather launch, one scatter launch in the backward
        cur = shared_mlp(stage, upsample("nni", cur, xyz[:, :n_l], xyz[:, :n_up], dec_idx.get(lvl), skip=skips.pop()))
        n_l = n_up
    # The reference undoes the permutation before fc_end (modules.py:608).  fc_end acts per point (its BatchNorm batch
    # statistics are sums over all points), so undoing it AFTER fc_end gives the same logits while the gather and its
    # scatter-add backward move n_classes channels per point instead of 32.
    cur = shared_mlp(net.fc_end[0], cur)
    cur = shared_mlp(net.fc_end[1], cur)
    cur = F.dropout(cur, net.fc_end[2].p, net.fc_end[2].training)
    cur = shared_mlp(net.fc_end[3], cur)
    inv = torch.empty_like(perm)                  # inverse permutation by scatter (argsort is a 30 us radix sort)
    inv.scatter_(0, perm, torch.arange(perm.numel(), device=perm.device))
    return cur.index_select(1, inv).transpose(1, 2)


def _require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("3d_recognizer_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")


# ------------------------------------------------------------------- inference path: C-ABI kernels only
def _fold(smlp):
    """Eval-mode SharedMLP -> (wT (Cin,Cout), scale (Cout), shift (Cout)): y = scale*(W x) + shift with the
    conv bias and the BatchNorm running statistics (eps 1e-6, modules.py:87) folded in."""
    w = conv_weight_2d(smlp).detach()
    bias = smlp.conv.bias.detach()
    bn = smlp.batch_norm
    if bn is None:
        return w.t().contiguous(), torch.ones_like(bias), bias.clone()
    scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
    shift = bn.bias.detach() + (bias - bn.running_mean) * scale
    return w.t().contiguous(), scale.contiguous(), shift.contiguous()


def _act_of(smlp):
    a = smlp.activation
    if a is None:
        return None, 0.0
    if isinstance(a, torch.nn.ReLU):
        return "relu", 0.0
    if isinstance(a, torch.nn.LeakyReLU):
        return "lrelu", float(a.negative_slope)
    raise ValueError(f"unsupported activation {a}")


def _state_version(net) -> tuple:
    return tuple(t._version for t in net.state_dict(keep_vars=True).values())


def folded_parameters(net) -> dict:
    """Kernel-ready eval-mode parameters, cached on the module and rebuilt when any parameter or buffer
    changed (tensor version counters)."""
    ver = _state_version(net)
    cache = getattr(net, "_r3d_folded", None)
    if cache is not None and cache["version"] == ver:
        return cache
    with torch.no_grad():
        c = {"version": ver}
        bn0 = net.bn_start[0]
        s0 = bn0.weight * torch.rsqrt(bn0.running_var + bn0.eps)
        c["start"] = (net.fc_start.weight.t().contiguous(), s0.contiguous(),
                      (bn0.bias + (net.fc_start.bias - bn0.running_mean) * s0).contiguous())
        enc = []
        for lfa in net.encoder:
            e = {}
            e["mlp1"] = _fold(lfa.mlp1)
            w1T, a1, b1 = _fold(lfa.mlp_rpe1)
            e["rpe1"] = (w1T.t().contiguous(), a1, b1)                      # (h,10) [out][in]
            e["rpe2"] = _fold(lfa.mlp_rpe2)                                 # wT (h,h) [in][out]
            e["score1"] = lfa.pool1.score_fn[0].weight.detach().t().contiguous()
            e["score2"] = lfa.pool2.score_fn[0].weight.detach().t().contiguous()
            # stored [out][in] layouts for the tensor-core kernel
            e["score1_oi"] = lfa.pool1.score_fn[0].weight.detach().contiguous()
            e["score2_oi"] = lfa.pool2.score_fn[0].weight.detach().contiguous()
            e["rpe2_oi"] = conv_weight_2d(lfa.mlp_rpe2).detach().contiguous()
            e["pool1"] = _fold(lfa.pool1.mlp)
            e["pool2"] = _fold(lfa.pool2.mlp)
            # residual sum mlp2(p2) + shortcut(x) as ONE layer over [p2 ; x]: BN scales folded into W
            w2T, s2, t2 = _fold(lfa.mlp2)
            wsT, ss, ts = _fold(lfa.shortcut)
            e["res"] = (torch.cat((w2T * s2, wsT * ss), dim=0).contiguous(), None, (t2 + ts).contiguous())
            enc.append(e)
        c["encoder"] = enc
        c["mlp"] = _fold(net.mlp)
        c["decoder"] = [_fold(m) for m in net.decoder]
        c["end0"] = _fold(net.fc_end[0])
        c["end1"] = _fold(net.fc_end[1])
        c["end3"] = _fold(net.fc_end[3])
    net._r3d_folded = c
    return c


def forward_kernels(net, inp: torch.Tensor, permutation: np.ndarray) -> torch.Tensor:
    """Eval-mode forward on the sm_100a kernels only: per LFA block one KNN, two fused LocSE+pooling
    launches and four per-point layers; per decoder stage one 1-NN and one gather+concat+MLP launch."""
    s = net.settings
    dec, k = s.decimation, s.n_neighbors
    B, N, _ = inp.shape
    P = folded_parameters(net)
    dev = inp.device
    perm64 = _device_permutation(permutation, dev)
    perm = perm64.to(torch.int32)
    inp = inp.float().contiguous()

    # fc_start + bn_start + LeakyReLU(0.2) fused with the point permutation (modules.py:565-573)
    w, sc, sh = P["start"]
    cur = ops.pointwise(inp, w, sc, sh, "lrelu", 0.2, gidx=perm)
    xyz = inp[..., :3].index_select(1, perm64).contiguous()

    # Every neighbour search depends on the coordinates only: the searches of the down-sampled levels and of the decoder's
    # 1-NN up-sampling run on a side stream while level 0 is processed (as in forward_autograd; 1.8 of 18 ms at 32 x 65 536
    # points were searches queued behind kernels they do not depend on).
    L = len(net.encoder)
    sizes = [N]
    for _ in range(L):
        sizes.append(sizes[-1] // dec)           # points kept after each level (floor, modules.py:583)
    pre_fork, dec_fork, enc_idx, dec_idx = None, None, {}, {}
    if OVERLAP_WEIGHT_GRADS and L > 1:
        with _fork(dev, lane=3) as pre_fork:     # joined before level 1
            for lvl in range(1, L):
                enc_idx[lvl] = ops.knn(xyz[:, :sizes[lvl]], xyz[:, :sizes[lvl]], k, idx64=False, idx32=True,
                                       dist=False)["idx32"]
        with _fork(dev, lane=4) as dec_fork:     # joined before the decoder
            for lvl in range(L):                 # decoder stage lvl: sizes[L - lvl] -> sizes[L - lvl - 1] points
                dec_idx[lvl] = ops.knn(xyz[:, :sizes[L - lvl]], xyz[:, :sizes[L - lvl - 1]], 1, idx64=False, idx32=True,
                                       dist=False)["idx32"]

    skips: List[torch.Tensor] = []
    n_l = N
    for lvl, (lfa, e) in enumerate(zip(net.encoder, P["encoder"])):
        x = cur[:, :n_l]
        xyz_l = xyz[:, :n_l]
        if lvl == 1 and pre_fork is not None:
            pre_fork.join()
        idx = enc_idx[lvl] if lvl in enc_idx else ops.knn(xyz_l, xyz_l, k, idx64=False, idx32=True, dist=False)["idx32"]
        w, sc, sh = e["mlp1"]
        f = ops.pointwise(x, w, sc, sh, "lrelu", 0.2)
        w1, a1, b1 = e["rpe1"]
        w2T, a2, b2 = e["rpe2"]
        tc = ops.lfa_pool_tc_supported(2 * w1.shape[0], k, idx.shape[0] * idx.shape[1])
        pooled = (ops.lfa_pool_tc(1, xyz_l, idx, f, w1, a1, b1, None, None, None, e["score1_oi"]) if tc else
                  ops.lfa_pool(1, xyz_l, idx, f, w1, a1, b1, None, None, None, e["score1"]))
        w, sc, sh = e["pool1"]
        p1 = ops.pointwise(pooled, w, sc, sh, "relu")
        pooled = (ops.lfa_pool_tc(2, xyz_l, idx, p1, w1, a1, b1, e["rpe2_oi"], a2, b2, e["score2_oi"]) if tc else
                  ops.lfa_pool(2, xyz_l, idx, p1, w1, a1, b1, w2T, a2, b2, e["score2"]))
        w, sc, sh = e["pool2"]
        p2 = ops.pointwise(pooled, w, sc, sh, "relu")
        w, sc, sh = e["res"]
        cur = ops.pointwise(p2, w, sc, sh, "lrelu", 0.01, xb=x)
        skips.append(cur)
        n_l //= dec
    w, sc, sh = P["mlp"]
    cur = ops.pointwise(cur[:, :n_l], w, sc, sh, "relu")
    if dec_fork is not None:
        dec_fork.join()
    for lvl, (w, sc, sh) in enumerate(P["decoder"]):
        skip = skips.pop()
        n_up = skip.shape[1]
        nn1 = (dec_idx[lvl] if lvl in dec_idx else
               ops.knn(xyz[:, :n_l], xyz[:, :n_up], 1, idx64=False, idx32=True, dist=False)["idx32"])
        cur = ops.pointwise(cur, w, sc, sh, "relu", gidx=nn1.view(B, n_up), xb=skip)
        n_l = n_up
    inv = torch.empty_like(perm)
    inv[perm64] = torch.arange(N, dtype=torch.int32, device=dev)
    w, sc, sh = P["end0"]
    cur = ops.pointwise(cur, w, sc, sh, "relu", gidx=inv)              # inverse permutation (modules.py:608)
    w, sc, sh = P["end1"]
    cur = ops.pointwise(cur, w, sc, sh, "relu")
    w, sc, sh = P["end3"]                                              # Dropout is the identity in eval mode
    return ops.pointwise(cur, w, sc, sh, None, transpose_out=True)


def forward(net, inp: torch.Tensor, permutation: np.ndarray) -> torch.Tensor:
    if inp.device != net.device:
        inp = inp.to(net.device)
    _require_cuda(inp)
    if not net.training and not torch.is_grad_enabled():
        return forward_kernels(net, inp, permutation)
    return forward_autograd(net, inp, permutation)
