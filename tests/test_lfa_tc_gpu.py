"""GPU parity of the tensor-core ("channel-lane") training kernels (r3d_lfa_tc_bwd modes 1-4, csrc/lfa_cl_bwd.cu)
against the FP32 CUDA-core kernels they replace (r3d_lfa_pool_bwd, r3d_lfa_pool2_bwd_train, r3d_lfa_bn2_bwd,
r3d_lfa_moments), which in turn are checked against autograd of the tensor-op composition and the reference's golden
vectors (tests/test_train_gpu.py).  Every width the kernels are built for, ragged sizes, K = 16 and 32."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


@pytest.fixture(scope="module")
def ops():
    return importlib.import_module("3d_recognizer_b200.ops")


class _force:
    """Run the block with the tensor-core kernels forced on for every built width, or forced off."""

    def __init__(self, ops, on):
        self.ops, self.on = ops, on

    def __enter__(self):
        o = self.ops
        self.saved = (o.USE_TENSOR_CORES, o.TC_BWD_WIDTHS, o.TC_WIDTHS, o.TC_MOM_WIDTHS, o.TC_FORCE)
        o.USE_TENSOR_CORES = self.on
        o.TC_FORCE = True                                  # also below the size thresholds of the dispatcher
        o.TC_BWD_WIDTHS = o.TC_WIDTHS = o.TC_MOM_WIDTHS = o.TC_ALL_WIDTHS

    def __exit__(self, *exc):
        o = self.ops
        o.USE_TENSOR_CORES, o.TC_BWD_WIDTHS, o.TC_WIDTHS, o.TC_MOM_WIDTHS, o.TC_FORCE = self.saved
        return False


def _case(d, K, B, N, seed):
    h = d // 2
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = dict(
        xyz=torch.rand(B, N, 3, device="cuda", generator=g),
        feat=torch.randn(B, N, h, device="cuda", generator=g),
        w1=torch.randn(h, 10, device="cuda", generator=g),
        a1=torch.rand(h, device="cuda", generator=g) + 0.5,
        b1=torch.randn(h, device="cuda", generator=g) * 0.3,
        w2=(torch.randn(h, h, device="cuda", generator=g) / h ** 0.5).contiguous(),
        a2=torch.rand(h, device="cuda", generator=g) + 0.5,
        b2=torch.randn(h, device="cuda", generator=g) * 0.3,
        ws=(torch.randn(d, d, device="cuda", generator=g) / d ** 0.5).contiguous(),
        # heavy-tailed upstream gradient of realistic magnitude (a mean loss over B*N points)
        dp=torch.randn(B, N, d, device="cuda", generator=g) * torch.exp(torch.randn(B, N, 1, device="cuda", generator=g))
        * (1e-3 / (B * N)),
    )
    return t


SHAPES = [(16, 16), (32, 16), (64, 16), (128, 16), (256, 16), (16, 32), (64, 32), (128, 32), (256, 32)]
SIZES = [(2, 1000), (3, 37), (1, 9001)]


@pytest.mark.parametrize("d,K", SHAPES)
@pytest.mark.parametrize("B,N", SIZES)
def test_stage1_backward_tc_vs_fp32(ops, d, K, B, N):
    t = _case(d, K, B, N, d + K + N)
    idx = ops.knn(t["xyz"], t["xyz"], K, idx64=False, idx32=True, dist=False)["idx32"]
    out = {}
    for on in (False, True):
        with _force(ops, on):
            out[on] = ops.lfa_pool_bwd(1, t["xyz"], idx, t["feat"], t["w1"], t["a1"], t["b1"], None, None, None, None,
                                       t["ws"].t().contiguous(), t["ws"], t["dp"])
    ops.check_tc_status(t["xyz"].device)
    for name, a, b in zip(("dfeat", "dw_score", "g1"), out[True][:3], out[False][:3]):
        assert rel(a, b) < 2e-5, (name, rel(a, b))


def _du2_canonical_tc(du2, B, N, K, d):
    """tensor-core tile layout [tile][lane = sub*h + c][row = p*K + k] -> (B*N, K, h)"""
    h, sub, pts = d // 2, 128 // d, 64 // K
    nt = du2.numel() // 4096
    return du2.view(nt, sub, h, pts, K).permute(0, 1, 3, 4, 2).reshape(nt * sub * pts, K, h)[:B * N]


def _du2_to_tc(can, B, N, K, d):
    h, sub, pts = d // 2, 128 // d, 64 // K
    nt = -(-(B * N) // (sub * pts))
    full = torch.zeros(nt * sub * pts, K, h, device=can.device)
    full[:B * N] = can
    return full.view(nt, sub, pts, K, h).permute(0, 1, 4, 2, 3).contiguous().view(-1)


def _du2_canonical_fp32(du2, B, N, K, d, P):
    """CUDA-core tile layout [b][tile][h][P*K] -> (B*N, K, h)"""
    h = d // 2
    T = -(-N // P)
    return du2.view(B, T, h, P, K).permute(0, 1, 3, 4, 2).reshape(B, T * P, K, h)[:, :N].reshape(B * N, K, h)


@pytest.mark.parametrize("d,K", SHAPES)
@pytest.mark.parametrize("B,N", SIZES)
def test_stage2_train_backward_tc_vs_fp32(ops, d, K, B, N):
    """Both BatchNorm passes.  Pass 1: dfeat and dw_score directly; du2 element by element (the two kernel families round
    mlp_rpe2's pre-activation differently, so a handful of elements within round-off of the ReLU kink may take the other
    branch: <= 1e-5 of the elements, and only as zero <-> non-zero); the batch sums against sums of the kernel's own
    du2.  Pass 2: both families on the SAME du2 and coefficients."""
    import importlib
    cabi = importlib.import_module("3d_recognizer_b200._cabi")
    engine = importlib.import_module("3d_recognizer_b200.engine")
    t = _case(d, K, B, N, 7 * d + K + N)
    h = d // 2
    nn_ = ops.knn(t["xyz"], t["xyz"], K, idx64=True, idx32=True, dist=True)
    idx = nn_["idx32"]
    rows = float(B * N * K)
    with _force(ops, False):
        m_r1, s_r1 = ops.lfa_moments(1, t["xyz"], idx, d, t["w1"], t["a1"], t["b1"])
    bn = torch.nn.BatchNorm2d(h, eps=1e-6, momentum=0.99).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.7, 1.3)
        bn.bias.uniform_(-0.2, 0.2)
    a2, c2, save2 = ops.bn_from_moments(t["w2"], s_r1[:, 10], m_r1, rows, bn, None)
    rpe = engine.relative_position_encoding(t["xyz"].double(), nn_["idx64"], nn_["dist"].double()).reshape(-1, 10)
    r1 = torch.relu(rpe @ t["w1"].double().t() * t["a1"].double() + t["b1"].double())
    r2 = torch.relu(r1 @ t["w2"].double().t() * a2.double() + c2.double()).view(B * N, K, h)
    out = {}
    for on in (False, True):
        with _force(ops, on):
            dfeat, dws, du2, sums = ops.lfa_pool2_bwd_train(t["xyz"], idx, t["feat"], t["w1"], t["a1"], t["b1"],
                                                            t["w2"].t().contiguous(), a2, c2, t["ws"].t().contiguous(),
                                                            t["ws"], t["dp"], w_rpe2=t["w2"])
        if on and isinstance(du2, tuple):
            can = _du2_canonical_tc(du2[0], B, N, K, d)
            assert abs(float(du2[1][1]) - float(can.abs().max())) <= 1e-6 * float(can.abs().max())    # absmax for pass 2
        else:                                          # d = 256 hands pass 2 the CUDA-core layout from either family
            can = _du2_canonical_fp32(du2, B, N, K, d, cabi.lib().r3d_lfa_tile_points_for(K, d, B, N))
        # the kernel's batch sums are the sums of its own du2
        assert rel(sums[0], can.double().sum(dim=(0, 1))) < 1e-5
        assert rel(sums[1], (can.double() * r2).sum(dim=(0, 1))) < 1e-5
        out[on] = dict(dfeat=dfeat, dw_score=dws, du2=can, sums=sums)
    ops.check_tc_status(t["xyz"].device)
    assert rel(out[True]["dfeat"], out[False]["dfeat"]) < 2e-5
    assert rel(out[True]["dw_score"], out[False]["dw_score"]) < 2e-5
    a, b = out[True]["du2"], out[False]["du2"]
    bad = (a - b).abs() > 2e-5 * b.abs().max()
    assert int(bad.sum()) <= max(2, int(1e-5 * a.numel())), int(bad.sum())
    assert bool((((a == 0) | (b == 0)) | ~bad).all()), "du2 differs by more than a ReLU branch flip"
    if d == 256:
        return                                         # pass 2 of the widest level runs on the CUDA-core kernel either way
    # pass 2 on identical inputs
    bn2, _, _ = ops.lfa_bn2_coeffs(out[False]["sums"], a2, c2, save2, rows)
    scal = torch.tensor([0.0, float(b.abs().max())], device="cuda")
    res = {}
    for on, du2 in ((False, None), (True, (_du2_to_tc(b, B, N, K, d), scal))):
        with _force(ops, False):                       # du2 decides the kernel family of pass 2
            if du2 is None:
                _, _, du2, _ = ops.lfa_pool2_bwd_train(t["xyz"], idx, t["feat"], t["w1"], t["a1"], t["b1"],
                                                       t["w2"].t().contiguous(), a2, c2, t["ws"].t().contiguous(),
                                                       t["ws"], t["dp"])
        g1, dw2 = ops.lfa_bn2_bwd(t["xyz"], idx, t["w1"], t["a1"], t["b1"], du2, t["w2"].t().contiguous(), t["w2"], bn2, h)
        res[on] = (g1[:, :11], dw2)
    ops.check_tc_status(t["xyz"].device)
    assert rel(res[True][0], res[False][0]) < 5e-5, rel(res[True][0], res[False][0])
    assert rel(res[True][1], res[False][1]) < 5e-5, rel(res[True][1], res[False][1])


@pytest.mark.parametrize("d,K", [s for s in SHAPES if s[0] <= 128])
@pytest.mark.parametrize("B,N", SIZES)
def test_r1_moments_tc_vs_fp32(ops, d, K, B, N):
    t = _case(d, K, B, N, 3 * d + K + N)
    idx = ops.knn(t["xyz"], t["xyz"], K, idx64=False, idx32=True, dist=False)["idx32"]
    out = {}
    for on in (False, True):
        with _force(ops, on):
            out[on] = ops.lfa_moments(1, t["xyz"], idx, d, t["w1"], t["a1"], t["b1"])
    assert rel(out[True][0], out[False][0]) < 1e-5
    assert rel(out[True][1][:, 10], out[False][1][:, 10]) < 2e-6
    # what the statistics are used for: the covariance (second moments minus the outer product of the means) must
    # agree as well as the raw moments do — the tensor-core kernel accumulates CENTRED values for exactly this reason
    n = float(B * N * K)
    cov = [m.double() / n - torch.outer(s[:, 10].double(), s[:, 10].double()) / n ** 2 for m, s in (out[True], out[False])]
    assert rel(cov[0], cov[1]) < 2e-5


def test_accumulator_fold_keeps_fp32_accuracy(ops):
    """A long reduction (>= 3 folds of the first-level TMEM accumulator per group) against an fp64 product."""
    d, K, B, N = 64, 16, 4, 40000
    t = _case(d, K, B, N, 5)
    idx = ops.knn(t["xyz"], t["xyz"], K, idx64=False, idx32=True, dist=False)["idx32"]
    with _force(ops, True):
        m_r1, s_r1 = ops.lfa_moments(1, t["xyz"], idx, d, t["w1"], t["a1"], t["b1"])
    engine = importlib.import_module("3d_recognizer_b200.engine")
    nn_ = ops.knn(t["xyz"], t["xyz"], K, idx64=True, idx32=False, dist=True)
    rpe = engine.relative_position_encoding(t["xyz"].double(), nn_["idx64"], nn_["dist"].double()).reshape(-1, 10)
    r1 = torch.relu(rpe @ t["w1"].double().t() * t["a1"].double() + t["b1"].double())
    assert rel(m_r1, r1.t() @ r1) < 3e-6
    assert rel(s_r1[:, 10], r1.sum(0)) < 3e-6
