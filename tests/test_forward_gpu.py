"""GPU parity of the fused inference kernels (C ABI r3d_lfa_pool / r3d_pointwise) and of the whole
eval-mode forward / Model.predict against the oracle and the golden vectors written by the REFERENCE
modules (oracle/make_golden.py).  Bar: logits within 1e-4 relative in fp32 (north_star)."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import network as onet

pytestmark = pytest.mark.gpu
TOL = 1e-4

E2E = {
    "k16_n1024": (dict(n_classes=2, n_points=1024, n_features=0, n_neighbors=16, knn="naive"), 2, 1024, 11),
    "k32_n2500": (dict(n_classes=2, n_points=2500, n_features=0, n_neighbors=32, knn="naive"), 1, 2500, 12),
    "k16_f2_c3_n1100": (dict(n_classes=3, n_points=1100, n_features=2, n_neighbors=16, knn="approximate"), 2, 1100, 13),
}


def make_input(B, N, Fe, seed):
    rng = np.random.RandomState(seed)
    x = rng.rand(B, N, 3 + Fe).astype(np.float32)
    x[..., :3] = x[..., :3] * np.array([0.78, 0.61, 0.55], np.float32) + np.array([-0.44, -0.31, 0.05], np.float32)
    return x


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.fixture(scope="module")
def mods():
    return (importlib.import_module("3d_recognizer_b200.modules"), importlib.import_module("3d_recognizer_b200.engine"),
            importlib.import_module("3d_recognizer_b200.ops"))


@pytest.mark.parametrize("B,n,ca,cb,cout,act,gather,tr", [
    (2, 1000, 3, 0, 8, "lrelu", "shared", False), (3, 517, 8, 0, 8, "lrelu", None, False),
    (2, 700, 16, 8, 32, "lrelu", None, False), (2, 300, 32, 0, 2, None, None, True),
    (2, 640, 512, 512, 256, "relu", "per_cloud", False), (1, 2500, 64, 64, 8, "relu", "per_cloud", False),
    (2, 999, 128, 0, 64, "relu", None, False), (4, 39, 512, 0, 512, "relu", None, False),
    (2, 333, 8, 0, 64, "relu", "shared", False), (1, 77, 64, 0, 32, "relu", None, False),
    (2, 100, 5, 0, 8, "lrelu", "shared", False),
    (2, 2560, 512, 512, 256, "relu", "per_cloud", False), (4, 1100, 128, 0, 512, "relu", None, False),
    (1, 5000, 64, 32, 128, "lrelu", "shared", False), (2, 4096, 32, 0, 64, None, None, False),
    # few rows, long contraction: clusters of 4 CTAs split the input channels (distributed shared memory reduction)
    (1, 312, 256, 0, 512, "relu", None, False), (2, 36, 520, 0, 64, None, None, False),
    (1, 72, 512, 512, 256, "lrelu", "per_cloud", False), (3, 100, 256, 0, 36, "relu", None, True),
    # >= 32768 rows with a long contraction or a narrow output: the tcgen05 path (pw_tc_eligible)
    (2, 16384, 512, 512, 256, "relu", "per_cloud", False), (4, 8200, 256, 0, 32, "relu", None, False),
    (1, 33000, 64, 32, 32, "lrelu", "shared", False), (3, 11000, 128, 128, 512, None, None, False),
])
def test_pointwise_vs_torch(mods, B, n, ca, cb, cout, act, gather, tr):
    _, _, ops = mods
    g = torch.Generator(device="cuda").manual_seed(ca * 1000 + cout)
    na = n if gather is None else max(n // 3, 1) if gather == "per_cloud" else n
    xa = torch.randn(B, na, ca, device="cuda", generator=g)
    xb = torch.randn(B, n, cb, device="cuda", generator=g) if cb else None
    wT = torch.randn(ca + cb, cout, device="cuda", generator=g) / (ca + cb) ** 0.5
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    sh = torch.randn(cout, device="cuda", generator=g)
    gidx = None
    if gather == "shared":
        gidx = torch.randperm(n, device="cuda", generator=g).to(torch.int32)
    elif gather == "per_cloud":
        gidx = torch.randint(0, na, (B, n), device="cuda", generator=g, dtype=torch.int32)
    got = ops.pointwise(xa, wT, sc, sh, act, 0.2, gidx=gidx, xb=xb, transpose_out=tr)
    src = xa
    if gather == "shared":
        src = xa[:, gidx.long()]
    elif gather == "per_cloud":
        src = torch.gather(xa, 1, gidx.long().unsqueeze(-1).expand(B, n, ca))
    full = torch.cat((src, xb), dim=-1) if cb else src
    ref = (full.double() @ wT.double()) * sc.double() + sh.double()
    ref = F.relu(ref) if act == "relu" else F.leaky_relu(ref, 0.2) if act == "lrelu" else ref
    if tr:
        ref = ref.transpose(1, 2)
    assert got.shape == ref.shape
    # wide layers run on tcgen05 with the 3xTF32 split (error up to ~1e-5 at C_in = 1024), the rest in plain fp32
    assert rel_err(got.double(), ref) < 2e-5
    # and the two back-ends agree with each other
    L = importlib.import_module("3d_recognizer_b200._cabi").lib()
    prev = L.r3d_pointwise_set_tensor_cores(0)
    try:
        got_cc = ops.pointwise(xa, wT, sc, sh, act, 0.2, gidx=gidx, xb=xb, transpose_out=tr)
    finally:
        L.r3d_pointwise_set_tensor_cores(prev)
    assert rel_err(got_cc.double(), ref) < 2e-6
    assert rel_err(got, got_cc) < 2e-5


@pytest.mark.parametrize("d,K", [(16, 16), (64, 16), (128, 16), (256, 16), (16, 32), (32, 32), (64, 32), (128, 32),
                                 (256, 32), (32, 16)])
@pytest.mark.parametrize("stage", [1, 2])
def test_lfa_pool_vs_torch(mods, d, K, stage):
    """One fused LocSE+pooling launch vs the same math in fp64 torch ops (engine.* composition)."""
    _, engine, ops = mods
    h = d // 2
    B, N = 2, 333 if d >= 128 else 1000
    g = torch.Generator(device="cuda").manual_seed(d + K + stage)
    xyz = torch.rand(B, N, 3, device="cuda", generator=g)
    feat = torch.randn(B, N, h, device="cuda", generator=g)
    nn_ = ops.knn(xyz, xyz, K, idx64=True, idx32=True, dist=True)
    w1 = torch.randn(h, 10, device="cuda", generator=g)
    a1 = torch.rand(h, device="cuda", generator=g) + 0.5
    b1 = torch.randn(h, device="cuda", generator=g) * 0.3
    w2 = torch.randn(h, h, device="cuda", generator=g) / h ** 0.5
    a2 = torch.rand(h, device="cuda", generator=g) + 0.5
    b2 = torch.randn(h, device="cuda", generator=g) * 0.3
    ws = torch.randn(d, d, device="cuda", generator=g) / d ** 0.5
    got = ops.lfa_pool(stage, xyz, nn_["idx32"], feat, w1, a1, b1, w2.t().contiguous() if stage == 2 else None,
                       a2 if stage == 2 else None, b2 if stage == 2 else None, ws.t().contiguous())
    D = torch.float64
    rpe = engine.relative_position_encoding(xyz.to(D), nn_["idx64"], nn_["dist"].to(D))
    r = F.relu(rpe @ w1.to(D).t() * a1.to(D) + b1.to(D))
    if stage == 2:
        r = F.relu(r @ w2.to(D).t() * a2.to(D) + b2.to(D))
    x = torch.cat((r, engine.gather_points(feat.to(D), nn_["idx64"])), dim=-1)
    ref = (F.softmax(x @ ws.to(D).t(), dim=2) * x).sum(dim=2)
    assert rel_err(got.to(D), ref) < 5e-6


@pytest.mark.parametrize("name", list(E2E))
def test_eval_forward_vs_reference_golden(mods, name):
    modules, engine, _ = mods
    g = np.load(os.path.join(GOLDEN, "e2e_golden.npz"))
    st, B, N, seed = E2E[name]
    net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
    net.load_state_dict(onet.synth_state_dict(st, seed))
    net.eval()
    x = torch.from_numpy(make_input(B, N, st["n_features"], seed)).cuda()
    np.random.seed(seed)
    with torch.no_grad():
        logits = net(x)
    ref = torch.from_numpy(g[f"{name}/eval_logits"]).cuda()
    assert logits.shape == ref.shape
    assert rel_err(logits, ref) < TOL
    # the kernel path really ran (not the autograd composition)
    np.random.seed(seed)
    perm = np.random.permutation(N)
    with torch.no_grad():
        a = engine.forward_kernels(net, x, perm)
        b = engine.forward_autograd(net, x, perm)
    assert torch.equal(a, logits)
    assert rel_err(a, b) < TOL


def test_eval_forward_larger_cloud_vs_oracle(mods):
    """N=8192, K=16, B=2: all four levels have full CTAs; checked against the oracle port on the CPU."""
    modules, _, _ = mods
    st = dict(n_classes=2, n_points=8192, n_features=0, n_neighbors=16, knn="naive")
    sd = onet.synth_state_dict(st, 5)
    net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
    net.load_state_dict(sd)
    net.eval()
    x = torch.from_numpy(make_input(2, 8192, 0, 5))
    np.random.seed(1)
    with torch.no_grad():
        got = net(x.cuda()).cpu()
    np.random.seed(1)
    ref = onet.forward({k: v.clone() for k, v in sd.items()}, st, x, training=False)
    assert rel_err(got, ref) < TOL


def test_model_predict_vs_reference_golden(tmp_path):
    """Model.load(save) round trip + Model.predict on a mock LiDAR cloud vs the reference façade's output."""
    model_mod = importlib.import_module("3d_recognizer_b200.model")
    modules = importlib.import_module("3d_recognizer_b200.modules")
    g = np.load(os.path.join(GOLDEN, "predict_golden.npz"))
    st = dict(n_classes=2, n_points=2500, n_features=0, n_neighbors=32, knn="naive")
    m = model_mod.Model(modules.RandLANetSettings(**st), weights=onet.synth_state_dict(st, 21))
    m.save(tmp_path / "ckpt")
    for ap in ("nni", "idw"):
        m2 = model_mod.Model.load(tmp_path / "ckpt", upsampling=ap)
        assert m2.settings.upsampling == ap
        np.random.seed(3)
        conf = m2.predict(g["cloud"])
        ref = g[f"conf_{ap}"]
        assert conf.shape == ref.shape
        assert np.abs(conf - ref).max() < 1e-4


@pytest.mark.parametrize("M,N,K", [(128, 32, 8), (1000, 64, 64), (777, 128, 128), (4096, 256, 256), (300, 256, 1024)])
def test_tcgen05_gemm_3xtf32_matches_fp64(M, N, K):
    """r3d_tc_gemm (tcgen05.mma kind::tf32, TMEM accumulators): the 3xTF32 split reaches fp32-level accuracy, the
    single-term variant shows plain TF32 error — i.e. the tensor-core path really ran and the split is what fixes it."""
    cabi = importlib.import_module("3d_recognizer_b200._cabi")
    L = cabi.lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    ref = A.double() @ W.double().t()
    errs = {}
    for terms in (1, 3):
        C = torch.zeros(M, N, device="cuda")
        cabi.check(L.r3d_tc_gemm(cabi.ptr(A), cabi.ptr(W), cabi.ptr(C), M, N, K, terms, cabi.stream_ptr(A.device)), "r3d_tc_gemm")
        errs[terms] = rel_err(C.double(), ref)
    assert errs[3] < 2e-5, errs
    assert 1e-5 < errs[1] < 5e-3, errs


@pytest.mark.parametrize("d,K", [(16, 16), (32, 16), (64, 16), (128, 16), (256, 16), (16, 32), (64, 32), (128, 32),
                                 (256, 32)])
@pytest.mark.parametrize("stage", [1, 2])
@pytest.mark.parametrize("B,N", [(2, 1000), (3, 37), (1, 20011)])
def test_lfa_pool_tensor_core_vs_cuda_core(mods, d, K, stage, B, N):
    """r3d_lfa_pool_tc (channel-lane kernel: score GEMM + mlp_rpe2 as split-fp16 tcgen05 MMAs) against the FP32
    CUDA-core kernel; ragged sizes (tail tiles, clouds that end inside a tile), 1 to ~150 tiles per CTA group."""
    _, _, ops = mods
    h = d // 2
    g = torch.Generator(device="cuda").manual_seed(d + K + stage)
    xyz = torch.rand(B, N, 3, device="cuda", generator=g)
    feat = torch.randn(B, N, h, device="cuda", generator=g)
    idx = ops.knn(xyz, xyz, K, idx64=False, idx32=True, dist=False)["idx32"]
    w1 = torch.randn(h, 10, device="cuda", generator=g)
    a1 = torch.rand(h, device="cuda", generator=g) + 0.5
    b1 = torch.randn(h, device="cuda", generator=g) * 0.3
    w2 = (torch.randn(h, h, device="cuda", generator=g) / h ** 0.5).contiguous()
    ws = (torch.randn(d, d, device="cuda", generator=g) / d ** 0.5).contiguous()
    s2 = stage == 2
    ref = ops.lfa_pool(stage, xyz, idx, feat, w1, a1, b1, w2.t().contiguous() if s2 else None, a1 if s2 else None,
                       b1 if s2 else None, ws.t().contiguous())
    got = ops.lfa_pool_tc(stage, xyz, idx, feat, w1, a1, b1, w2 if s2 else None, a1 if s2 else None,
                          b1 if s2 else None, ws)
    assert rel_err(got, ref) < 1e-5
    ops.check_tc_status(xyz.device)


@pytest.mark.parametrize("ca,cout", [(8, 8), (8, 64), (16, 32), (32, 64), (64, 32), (64, 8), (32, 16), (16, 16)])
@pytest.mark.parametrize("w_out_in,act", [(False, "relu"), (True, None)])
def test_pointwise_streaming_rows_kernel(mods, ca, cout, w_out_in, act):
    """pw_rows_kernel (dense narrow layers on >= 131072 rows: the level-0 layers of a large batch): values, the
    batch statistics of the train-mode epilogue, a row count that is not a multiple of the 128-row block."""
    _, _, ops = mods
    L = importlib.import_module("3d_recognizer_b200._cabi").lib()
    B, n = 2, 100003
    assert L.r3d_pointwise_plan(ca, 0, cout, B * n, 0) == 4
    g = torch.Generator(device="cuda").manual_seed(ca * 100 + cout)
    x = torch.randn(B, n, ca, device="cuda", generator=g)
    shape = (cout, ca) if w_out_in else (ca, cout)
    w = torch.randn(*shape, device="cuda", generator=g) / ca ** 0.5
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    sh = torch.randn(cout, device="cuda", generator=g)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    got = ops.pointwise(x, w, sc, sh, act, 0.0, stats=stats, w_out_in=w_out_in)
    wm = w.double().t() if w_out_in else w.double()
    ref = (x.double() @ wm) * sc.double() + sh.double()
    ref = F.relu(ref) if act == "relu" else ref
    assert rel_err(got.double(), ref) < 2e-6
    flat = ref.reshape(-1, cout)
    assert rel_err(stats[:cout], flat.sum(0)) < 1e-5 * max(1.0, float(flat.abs().sum(0).max() / flat.sum(0).abs().max()))
    assert rel_err(stats[cout:], (flat * flat).sum(0)) < 1e-5


@pytest.mark.parametrize("ca,cout", [(2, 32), (3, 64), (4, 20), (1, 32)])
@pytest.mark.parametrize("w_out_in,act", [(False, None), (True, "lrelu")])
def test_pointwise_expand_kernel(mods, ca, cout, w_out_in, act):
    """pw_expand_kernel (C_in <= 4 into a wider layer on >= 131072 rows: the input gradient of the class-logits layer)."""
    _, _, ops = mods
    L = importlib.import_module("3d_recognizer_b200._cabi").lib()
    B, n = 2, 70001
    assert L.r3d_pointwise_plan(ca, 0, cout, B * n, 0) == 5
    g = torch.Generator(device="cuda").manual_seed(ca * 100 + cout)
    x = torch.randn(B, n, ca, device="cuda", generator=g)
    w = torch.randn(*((cout, ca) if w_out_in else (ca, cout)), device="cuda", generator=g)
    sc = torch.rand(cout, device="cuda", generator=g) + 0.5
    sh = torch.randn(cout, device="cuda", generator=g)
    got = ops.pointwise(x, w, sc, sh, act, 0.1, w_out_in=w_out_in)
    ref = (x.double() @ (w.double().t() if w_out_in else w.double())) * sc.double() + sh.double()
    ref = F.leaky_relu(ref, 0.1) if act == "lrelu" else ref
    assert got.shape == (B, n, cout) and rel_err(got.double(), ref) < 1e-6
