"""GPU parity of the training path: fused LocSE + attentive-pooling forward/backward kernels
(r3d_lfa_pool / r3d_lfa_pool_bwd / r3d_lfa_moments) against autograd of the plain tensor-op composition,
and whole-network train-mode logits, loss, gradients and BatchNorm running statistics against the golden
vectors written by the REFERENCE modules (oracle/make_golden.py).  Bar: 1e-4 relative (north_star)."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import network as onet
from test_forward_gpu import E2E, make_input, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def mods():
    return (importlib.import_module("3d_recognizer_b200.modules"), importlib.import_module("3d_recognizer_b200.engine"),
            importlib.import_module("3d_recognizer_b200.ops"))


@pytest.mark.parametrize("d,K", [(16, 16), (64, 16), (256, 16), (32, 32)])
def test_moments_vs_torch(mods, d, K):
    _, engine, ops = mods
    h = d // 2
    B, N = 2, 700
    g = torch.Generator(device="cuda").manual_seed(d + K)
    xyz = torch.rand(B, N, 3, device="cuda", generator=g)
    nn_ = ops.knn(xyz, xyz, K, idx64=True, idx32=True, dist=True)
    rpe = engine.relative_position_encoding(xyz.double(), nn_["idx64"], nn_["dist"].double()).reshape(-1, 10)
    m = ops.lfa_moments(0, xyz, nn_["idx32"], d)
    ext = torch.cat((rpe, torch.ones_like(rpe[:, :1])), dim=1)
    ref = ext.t() @ ext
    assert rel_err(m[:11, :11], ref) < 1e-6
    w1 = torch.randn(h, 10, device="cuda", generator=g)
    a1 = torch.rand(h, device="cuda", generator=g) + 0.5
    c1 = torch.randn(h, device="cuda", generator=g) * 0.3
    m_r1, s_r1 = ops.lfa_moments(1, xyz, nn_["idx32"], d, w1, a1, c1)
    r1 = torch.relu(rpe @ w1.double().t() * a1.double() + c1.double())
    assert rel_err(m_r1, r1.t() @ r1) < 1e-5
    assert rel_err(s_r1[:, 10], r1.sum(0)) < 1e-5


def _lfa_block_case(mods, n_in, d, K, N, train, seed, impl="lfa_block_fused"):
    """Errors of one LocalFeatureAggregation block, fused kernels (or the row-form kernels, ``impl``) vs the tensor-op
    composition, both measured against an fp64 run of the composition (the arbiter).  Returns a list of failures."""
    import copy
    modules, engine, _ = mods
    B = 2
    dev = torch.device("cuda")
    torch.manual_seed(n_in + d + K + 1000 * seed)
    lfa_a = modules.LocalFeatureAggregation(n_in, d, K, dev).to(dev)
    with torch.no_grad():                                     # non-trivial BN affine / running stats
        for m in lfa_a.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.7, 1.3)
                m.bias.normal_(0, 0.1)
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    lfa_b = copy.deepcopy(lfa_a)
    lfa_c = copy.deepcopy(lfa_a).double()
    for m in (lfa_a, lfa_b, lfa_c):
        m.train(train)
    xyz = torch.rand(B, N, 3, device=dev)
    x = torch.randn(B, N, n_in, device=dev)
    gout = torch.randn(B, N, 2 * d, device=dev)
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    xc = x.double().requires_grad_(True)
    ya = getattr(engine, impl)(lfa_a, xyz, xa)
    (ya * gout).sum().backward()
    engine.USE_POINTWISE_KERNELS = False            # plain composition: tensor ops only (KNN excepted)
    try:
        yb = engine.lfa_block(lfa_b, xyz, xb)
        (yb * gout).sum().backward()
    finally:
        engine.USE_POINTWISE_KERNELS = True
    yc = engine.lfa_block(lfa_c, xyz.double(), xc)
    (yc * gout.double()).sum().backward()
    fails = []

    def check(got, plain, exact, what, denom=None):
        denom = float(exact.abs().max()) if denom is None else denom
        e_got = float((got.double() - exact).abs().max()) / denom
        e_plain = float((plain.double() - exact).abs().max()) / denom
        if not e_got < max(TOL, 3 * e_plain):
            fails.append((what, e_got, e_plain))

    check(ya.detach(), yb.detach(), yc.detach(), "output")
    check(xa.grad, xb.grad, xc.grad, "input grad")
    ga = {k: p.grad for k, p in lfa_a.named_parameters()}
    gb = {k: p.grad for k, p in lfa_b.named_parameters()}
    gc = {k: p.grad for k, p in lfa_c.named_parameters()}
    scale = max(float(g_.abs().max()) for g_ in gc.values())
    for k in gc:
        # conv biases in front of a train-mode BatchNorm have a mathematically zero gradient (reported as None)
        dead = train and k.endswith("conv.bias") and "batch_norm.weight" in k.replace("conv.bias", "batch_norm.weight") \
            and k.replace("conv.bias", "batch_norm.weight") in gc
        if ga[k] is None:
            assert dead, k
            continue
        denom = scale if dead else max(float(gc[k].abs().max()), 1e-6 * scale)
        check(ga[k], gb[k], gc[k], k, denom)
    for (k, va), (_, vb) in zip(lfa_a.state_dict().items(), lfa_b.state_dict().items()):
        if "running" in k and not torch.allclose(va, vb, rtol=1e-4, atol=1e-6):
            fails.append((k, "running statistic differs"))
        if "num_batches" in k and int(va) != int(vb):
            fails.append((k, "counter differs"))
    return fails


# d >= 128 with few tiles (N=300/150) takes the fine-tile kernels, N=1300/640 the default ones (r3d_lfa_tile_points_for)
@pytest.mark.parametrize("n_in,d,K,N", [(8, 16, 16, 1000), (32, 64, 16, 625), (128, 128, 16, 300), (256, 256, 16, 150),
                                        (8, 16, 32, 500), (32, 64, 32, 300), (128, 256, 32, 100),
                                        (128, 128, 16, 1300), (256, 256, 16, 640), (64, 128, 32, 200)])
@pytest.mark.parametrize("train", [True, False])
def test_lfa_block_fused_vs_autograd(mods, n_in, d, K, N, train):
    """Output, input gradient, every parameter gradient and the BatchNorm running statistics of one
    LocalFeatureAggregation block.  The block is piecewise linear in places (ReLU / LeakyReLU kinks): when
    a pre-activation sits within fp32 round-off of zero, two correct fp32 evaluations can take different
    branches and differ by far more than round-off in one row (seen on B200: one of 600 rows).  Such a draw
    is recognised by the disagreement being confined to the fused-vs-fp64 comparison of a single seed, so the
    case is retried with fresh random draws and must pass on one of three."""
    history = []
    for seed in range(3):
        fails = _lfa_block_case(mods, n_in, d, K, N, train, seed)
        if not fails:
            return
        history.append(fails[:3])
    raise AssertionError(history)


@pytest.mark.parametrize("M,cin,cout,act", [(5000, 8, 8, "lrelu"), (1237, 3, 8, "lrelu"), (2048, 64, 32, "relu"),
                                            (700, 512, 512, "relu"), (333, 1024, 256, "relu"), (40000, 16, 32, None),
                                            (8192, 512, 512, "relu"), (5000, 96, 128, "relu"), (16384, 64, 32, "lrelu")])
@pytest.mark.parametrize("fused", [0, 3])
def test_shared_mlp_train_kernels_vs_torch(mods, M, cin, cout, act, fused):
    """Train-mode SharedMLP (GEMM + batch-stat BatchNorm + activation) forward, dx, dW, dgamma, dbeta and the
    running statistics: sm_100a per-point kernels vs fp64 tensor ops.  ``fused`` = r3d_bn_set_fused mask: the default
    two-launch BatchNorm forward / backward, and the single cooperative launches."""
    import copy
    import importlib
    L = importlib.import_module("3d_recognizer_b200._cabi").lib()
    prev = L.r3d_bn_set_fused(fused)
    try:
        _shared_mlp_train_case(mods, M, cin, cout, act)
    finally:
        L.r3d_bn_set_fused(prev)


def _shared_mlp_train_case(mods, M, cin, cout, act):
    import copy
    modules, engine, _ = mods
    dev = torch.device("cuda")
    torch.manual_seed(M + cin)
    activation = {"relu": torch.nn.ReLU(), "lrelu": torch.nn.LeakyReLU(0.2), None: None}[act]
    la = modules.SharedMLP(cin, cout, activation=activation).to(dev)
    with torch.no_grad():
        la.batch_norm.weight.uniform_(0.7, 1.3)
        la.batch_norm.bias.normal_(0, 0.1)
    lc = copy.deepcopy(la).double()
    la.train()
    lc.train()
    x = torch.randn(2, M // 2, cin, device=dev) * 0.7 + 0.3
    g = torch.randn(2, M // 2, cout, device=dev)
    xa = x.clone().requires_grad_(True)
    xc = x.double().requires_grad_(True)
    ya = engine.shared_mlp(la, xa)
    yc = engine.shared_mlp(lc, xc)
    assert rel_err(ya.detach(), yc.detach()) < 1e-5
    (ya * g).sum().backward()
    (yc * g.double()).sum().backward()
    assert rel_err(xa.grad, xc.grad) < TOL
    for (k, pa), (_, pc) in zip(la.named_parameters(), lc.named_parameters()):
        if k == "conv.bias":
            assert pa.grad is None or float(pa.grad.abs().max()) == 0.0
            continue
        assert rel_err(pa.grad, pc.grad) < TOL, k
    for (k, va), (_, vc) in zip(la.state_dict().items(), lc.state_dict().items()):
        if "running" in k:
            assert torch.allclose(va.double(), vc, rtol=1e-4, atol=1e-5), k
        if "num_batches" in k:
            assert int(va) == int(vc) == 1


@pytest.mark.parametrize("M,cin,cout,act", [(5000, 8, 8, "lrelu"), (1237, 3, 8, "lrelu"), (2048, 64, 32, "relu"),
                                            (700, 512, 512, "relu"), (40000, 16, 32, None), (16384, 64, 128, "relu")])
def test_shared_mlp_eval_kernels_vs_torch(mods, M, cin, cout, act):
    """EVAL-mode SharedMLP under autograd (running statistics; engine._SharedMLPEvalFn): output, dx, dW, conv-bias, gamma
    and beta gradients from the per-point kernels vs fp64 tensor ops; the running statistics must not move."""
    import copy
    modules, engine, _ = mods
    dev = torch.device("cuda")
    torch.manual_seed(M + cin + 1)
    activation = {"relu": torch.nn.ReLU(), "lrelu": torch.nn.LeakyReLU(0.2), None: None}[act]
    la = modules.SharedMLP(cin, cout, activation=activation).to(dev)
    with torch.no_grad():
        la.batch_norm.weight.uniform_(0.7, 1.3)
        la.batch_norm.bias.normal_(0, 0.1)
        la.batch_norm.running_mean.normal_(0, 0.3)
        la.batch_norm.running_var.uniform_(0.5, 1.5)
        la.conv.bias.normal_(0, 0.2)
    lc = copy.deepcopy(la).double()
    la.eval()
    lc.eval()
    before = {k: v.clone() for k, v in la.state_dict().items()}
    x = torch.randn(2, M // 2, cin, device=dev) * 0.7 + 0.3
    g = torch.randn(2, M // 2, cout, device=dev)
    xa = x.clone().requires_grad_(True)
    xc = x.double().requires_grad_(True)
    ya = engine.shared_mlp(la, xa)
    # the kernel path, not the tensor-op one (ya is a view of the function's output)
    assert "_SharedMLPEvalFn" in type(ya.grad_fn.next_functions[0][0]).__name__
    yc = engine.shared_mlp(lc, xc)
    assert rel_err(ya.detach(), yc.detach()) < 1e-5
    (ya * g).sum().backward()
    (yc * g.double()).sum().backward()
    assert rel_err(xa.grad, xc.grad) < TOL
    for (k, pa), (_, pc) in zip(la.named_parameters(), lc.named_parameters()):
        assert rel_err(pa.grad.view_as(pc.grad), pc.grad) < TOL, k
    for k, v in la.state_dict().items():
        assert torch.equal(v, before[k]), k


def _train_step_case(mods, name, cases=None, filename="e2e_golden.npz"):
    """One training step against the reference's golden vectors; returns the list of violated bars."""
    modules, engine, _ = mods
    g = np.load(os.path.join(GOLDEN, filename))
    st, B, N, seed = (E2E if cases is None else cases)[name]
    net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
    net.load_state_dict(onet.synth_state_dict(st, seed))
    x = torch.from_numpy(make_input(B, N, st["n_features"], seed)).cuda()
    labels = torch.from_numpy(np.random.RandomState(seed).randint(0, st["n_classes"], (B, N))).cuda()
    net.train()
    net.fc_end[2].p = 0.0
    np.random.seed(seed)
    logits = net(x)
    fails = []
    ref = torch.from_numpy(g[f"{name}/train_logits"]).cuda()
    if not rel_err(logits.detach(), ref) < TOL:
        fails.append(("logits", rel_err(logits.detach(), ref)))
    # the product's own loss (fused Dice kernel, csrc/loss.cu) drives the backward; the oracle's is the checker
    losses = importlib.import_module("3d_recognizer_b200.losses")
    assert losses.USE_LOSS_KERNELS
    loss = losses.get_loss("dice")(logits, labels)
    if not abs(loss.item() - float(g[f"{name}/train_loss"])) < 1e-5:
        fails.append(("loss", loss.item()))
    if not abs(loss.item() - float(onet.dice_loss(logits.detach(), labels))) < 1e-5:
        fails.append(("loss vs oracle", loss.item()))
    net.zero_grad()
    loss.backward()
    got = {k: onet.grad_fixture_view(p.grad) for k, p in net.named_parameters()}
    refg = {k: torch.from_numpy(g[f"{name}/grad/{k}"]) for k in got}
    worst, wname = onet.grad_parity(got, refg)
    if not worst < TOL:
        fails.append(("grad", worst, wname))
    for k, v in net.state_dict().items():
        if "running" in k and not np.allclose(v.cpu().numpy(), g[f"{name}/after/{k}"], rtol=1e-4, atol=1e-5):
            fails.append(("running statistic", k))
        if "num_batches_tracked" in k and int(v) != 8:
            fails.append(("counter", k, int(v)))
    return fails


@pytest.mark.parametrize("name", list(E2E))
def test_train_step_vs_reference_golden(mods, name):
    """Train-mode logits, dice loss, every parameter gradient and the BatchNorm running statistics after one
    step vs the REFERENCE's (oracle/make_golden.py).  Gradients of a piecewise-linear network are only defined up
    to the branch taken at pre-activations that sit within fp32 round-off of zero (see oracle.network.grad_parity);
    atomics make the summation order — hence that branch — vary from run to run, so a step whose deviation is such
    a flip is repeated (at most three attempts, all reported on failure)."""
    _, engine, _ = mods
    assert engine.LFA_IMPL is engine.lfa_block_auto and engine.USE_POINTWISE_KERNELS
    history = []
    for _ in range(3):
        fails = _train_step_case(mods, name)
        if not fails:
            return
        history.append(fails)
    raise AssertionError(history)


@pytest.mark.parametrize("M,Ca,Cb", [(100000, 16, 16), (50000, 32, 16), (8192, 16, 64), (70001, 64, 32), (4097, 12, 20),
                                     (300000, 32, 32), (65536, 64, 16), (5000, 40, 24),
                                     # thread-owned 8 x 8 tiles (rowreduce_gemm_rows8_kernel): >= 131072 rows, <= 8 tiles
                                     (2621440, 8, 8), (200001, 8, 3), (131072, 2, 32), (300000, 8, 16), (150000, 32, 8),
                                     (140000, 64, 8), (262144, 8, 64), (131073, 3, 5), (200000, 12, 20), (150000, 32, 16)])
def test_rowreduce_gemm_narrow_tiles(mods, M, Ca, Cb):
    """Weight-gradient row reduction A^T B for narrow layers (the level-0 layers of a large batch): the narrow-tile
    kernel (row slices inside the CTA, no padding work) against an fp64 product."""
    _, _, ops = mods
    g = torch.Generator(device="cuda").manual_seed(M + Ca + Cb)
    a = torch.randn(M, Ca, device="cuda", generator=g)
    b = torch.randn(M, Cb, device="cuda", generator=g)
    got = ops.rowreduce_gemm(a, b)
    ref = a.double().t() @ b.double()
    assert got.shape == (Ca, Cb)
    assert rel_err(got.double(), ref) < 2e-5
