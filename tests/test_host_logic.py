"""CPU tests of the host-side logic (no GPU): checkpoint schema, settings validation, and the
engine's tensor plumbing (point-major layouts, permutation, prefix down-sampling, decoder wiring)
checked against the golden vectors produced by the REFERENCE modules (oracle/make_golden.py).

The product has no CPU path: to exercise the host logic here the CUDA KNN entry point is replaced,
in this test only, by the oracle KNN, and the CUDA-device guard is lifted."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import network as onet
from oracle.knn import knn_exact

E2E = {
    "k16_n1024": (dict(n_classes=2, n_points=1024, n_features=0, n_neighbors=16, knn="naive"), 2, 1024, 11),
    "k32_n2500": (dict(n_classes=2, n_points=2500, n_features=0, n_neighbors=32, knn="naive"), 1, 2500, 12),
    "k16_f2_c3_n1100": (dict(n_classes=3, n_points=1100, n_features=2, n_neighbors=16, knn="approximate"), 2, 1100, 13),
}


def make_input(B, N, F, seed):
    rng = np.random.RandomState(seed)
    x = rng.rand(B, N, 3 + F).astype(np.float32)
    x[..., :3] = x[..., :3] * np.array([0.78, 0.61, 0.55], np.float32) + np.array([-0.44, -0.31, 0.05], np.float32)
    return x


@pytest.fixture()
def cpu_product(monkeypatch, r3d):
    ops = importlib.import_module("3d_recognizer_b200.ops")
    engine = importlib.import_module("3d_recognizer_b200.engine")
    modules = importlib.import_module("3d_recognizer_b200.modules")

    def knn_stub(support, query, k, idx64=True, idx32=False, dist=True, dist_sq=False):
        i, d2 = knn_exact(support.detach().numpy(), query.detach().numpy(), k)
        out = {}
        if idx64:
            out["idx64"] = torch.from_numpy(i)
        if idx32:
            out["idx32"] = torch.from_numpy(i.astype(np.int32))
        if dist:
            out["dist"] = torch.sqrt(torch.from_numpy(d2))
        if dist_sq:
            out["dist_sq"] = torch.from_numpy(d2)
        return out

    monkeypatch.setattr(ops, "knn", knn_stub)
    monkeypatch.setattr(engine, "_require_cuda", lambda t: None)
    monkeypatch.setattr(engine, "LFA_IMPL", engine.lfa_block)      # tensor-op composition instead of the fused kernels
    monkeypatch.setattr(engine, "forward_kernels", engine.forward_autograd)
    return modules


def test_settings_validation(r3d):
    m = importlib.import_module("3d_recognizer_b200.modules")
    s = m.RandLANetSettings(n_classes=2)
    assert (s.n_points, s.n_neighbors, s.decimation, s.layer_sizes, s.knn, s.upsampling) == \
        (10000, 32, 4, [16, 64, 128, 256], "approximate", "nni")
    with pytest.raises(AssertionError):
        m.RandLANetSettings(n_classes=2, knn="ball")
    with pytest.raises(AssertionError):
        m.RandLANetSettings(n_classes=2, upsampling="cubic")
    s.update(n_points=5, bogus=1)
    assert s.n_points == 5 and not hasattr(s, "bogus")


@pytest.mark.parametrize("name", list(E2E))
def test_state_dict_schema_matches_reference(r3d, name):
    m = importlib.import_module("3d_recognizer_b200.modules")
    st = E2E[name][0]
    net = m.RandLANet(m.RandLANetSettings(**st), torch.device("cpu"))
    sd = net.state_dict()
    schema = onet.state_dict_schema(st)        # verified == the reference's state_dict in make_golden.py
    assert list(sd.keys()) == list(schema.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == schema[k][0] and v.dtype == schema[k][1], k
    if st["n_features"] == 0 and st["n_classes"] == 2:
        assert sum(p.numel() for p in net.parameters()) == 1322666
        assert len(sd) == 262 and len(list(net.parameters())) == 154


def test_forward_asserts(cpu_product):
    m = cpu_product
    net = m.RandLANet(m.RandLANetSettings(n_classes=2, n_neighbors=16), torch.device("cpu"))
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 1024, 4))
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 1023, 3))          # min = max(16*64, 2*256) = 1024


@pytest.mark.parametrize("name", list(E2E))
def test_engine_host_logic_vs_reference_golden(cpu_product, name):
    m = cpu_product
    g = np.load(os.path.join(GOLDEN, "e2e_golden.npz"))
    st, B, N, seed = E2E[name]
    net = m.RandLANet(m.RandLANetSettings(**st), torch.device("cpu"))
    net.load_state_dict(onet.synth_state_dict(st, seed))
    x = torch.from_numpy(make_input(B, N, st["n_features"], seed))
    labels = torch.from_numpy(np.random.RandomState(seed).randint(0, st["n_classes"], (B, N)))

    net.eval()
    np.random.seed(seed)
    with torch.no_grad():
        logits = net(x)
    ref = torch.from_numpy(g[f"{name}/eval_logits"])
    assert logits.shape == ref.shape
    assert (logits - ref).abs().max() / ref.abs().max() < 1e-4

    net.train()
    net.fc_end[2].p = 0.0
    np.random.seed(seed)
    logits = net(x)
    ref = torch.from_numpy(g[f"{name}/train_logits"])
    assert (logits - ref).abs().max() / ref.abs().max() < 1e-4
    loss = onet.dice_loss(logits, labels)
    assert abs(loss.item() - float(g[f"{name}/train_loss"])) < 1e-5
    net.zero_grad()
    loss.backward()
    got = {k: onet.grad_fixture_view(p.grad) for k, p in net.named_parameters()}
    refg = {k: torch.from_numpy(g[f"{name}/grad/{k}"]) for k in got}
    worst, wname = onet.grad_parity(got, refg)
    assert worst < 1e-4, (worst, wname)
    for k, v in net.state_dict().items():
        if "running" in k:
            assert np.allclose(v.numpy(), g[f"{name}/after/{k}"], rtol=1e-4, atol=1e-5), k
        if "num_batches_tracked" in k:
            assert int(v) == 8


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no file of the product package may import or execute it."""
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d_recognizer_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/" not in src and "knn_oracle" not in src, f


def test_cabi_exports_every_declared_symbol(built_lib):
    """include/r3d_b200.h <-> lib/libr3d_b200.so <-> _cabi.SIGNATURES agree (no compute calls)."""
    import ctypes
    import re
    cabi = importlib.import_module("3d_recognizer_b200._cabi")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "r3d_b200.h")).read()
    declared = set(re.findall(r"\b(r3d_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(cabi.SIGNATURES), declared ^ set(cabi.SIGNATURES)
    handle = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(handle, name), name
    assert cabi.lib().r3d_abi_version() == 1
    assert cabi.lib().r3d_error_string(-2).decode().startswith("Not enough points")


def test_sample_points_matches_reference_protocol():
    """preprocessing.sample_points: same indices as the oracle restatement of preprocessing.py:35-62, global
    RNG state untouched by consistent draws."""
    pre = importlib.import_module("3d_recognizer_b200.preprocessing")
    np.random.seed(123)
    before = np.random.get_state()[1].copy()
    a = pre.sample_points(140801, 2500, consistent=True)
    assert list(a[:5]) == [110048, 47731, 136733, 49407, 61773]          # SURVEY.md §8c probe of the reference
    assert np.array_equal(np.random.get_state()[1], before)
    b = onet.sample_points(140801, 2500, consistent=True)
    assert np.array_equal(a, b)
    np.random.seed(5)
    c = pre.sample_points(30, 2500, consistent=False)                     # up-sampling with duplicates (predict.py:22)
    np.random.seed(5)
    assert np.array_equal(c, onet.sample_points(30, 2500, consistent=False))
    assert c.shape == (2500,) and c.max() < 30


def test_zero_pool_hands_out_fresh_zero_slices(r3d):
    """ops._ZeroPool (scratch for the kernels' atomics): every slice is zero-filled, aligned, disjoint from all earlier
    ones for as long as they are referenced, sized from the previous step, and large requests bypass the pool."""
    ops = importlib.import_module("3d_recognizer_b200.ops")
    pool = ops._ZeroPool()
    a = pool.take((3, 5), torch.float64, "cpu")
    b = pool.take(7, torch.float32, "cpu")
    assert a.shape == (3, 5) and a.dtype == torch.float64 and float(a.abs().sum()) == 0.0
    assert b.shape == (7,) and float(b.abs().sum()) == 0.0
    assert a.data_ptr() % 8 == 0 and (b.data_ptr() - a.data_ptr()) % 256 == 0
    a.fill_(1.0)
    b.fill_(2.0)
    c = pool.take(1000, torch.float32, "cpu")
    assert float(c.abs().sum()) == 0.0 and float(a.sum()) == 15.0 and float(b.sum()) == 14.0    # no aliasing
    need = pool.need
    pool.begin()                                                    # step boundary: new block, sized by the last step
    assert pool.hint >= need and pool.block is None
    d = pool.take((3, 5), torch.float64, "cpu")
    assert float(d.abs().sum()) == 0.0 and float(a.sum()) == 15.0   # the previous step's buffers stay valid
    big = pool.take(pool.MAX_BYTES // 4 + 1, torch.float32, "cpu")  # too large: its own allocation
    assert float(big.abs().sum()) == 0.0 and big.numel() == pool.MAX_BYTES // 4 + 1
    grown = [pool.take(1 << 16, torch.float32, "cpu") for _ in range(8)]   # outgrows the hint: further blocks
    assert all(float(t.abs().sum()) == 0.0 for t in grown)
    assert len({t.data_ptr() for t in grown}) == 8


def test_knn_and_pointwise_plans_are_pure_functions_of_the_shape(built_lib):
    """r3d_knn_plan / r3d_pointwise_plan / r3d_lfa_tile_points_for (host-only dispatch rules, no GPU needed)."""
    L = importlib.import_module("3d_recognizer_b200._cabi").lib()
    assert L.r3d_knn_set_algorithm(-1) == 0
    assert L.r3d_knn_plan(8, 2500, 2500, 16) == 2                      # uniform grid from 2048 support points
    assert L.r3d_knn_plan(8, 625, 625, 16) in (2, 3) and L.r3d_knn_plan(8, 156, 156, 16) == 3
    assert L.r3d_knn_plan(8, 625, 2500, 1) == 3                        # decoder 1-NN on a small support
    prev = L.r3d_knn_set_algorithm(1)
    try:
        assert L.r3d_knn_plan(8, 156, 156, 16) == 1
    finally:
        L.r3d_knn_set_algorithm(prev)
    assert L.r3d_pointwise_plan(3, 0, 8, 20000, 0) == 0                 # thin layer: weights in shared memory
    assert L.r3d_pointwise_plan(64, 0, 128, 5000, 0) == 2               # FP32 GEMM
    assert L.r3d_pointwise_plan(5, 0, 64, 5000, 0) == 1                 # channels not a multiple of 4
    assert L.r3d_pointwise_plan(512, 0, 256, 1 << 20, 0) == 3           # tcgen05: many rows, long contraction
    assert L.r3d_pointwise_plan(64, 0, 128, 1 << 20, 0) == 2            # mid-size layers stay on FP32
    assert L.r3d_lfa_tile_points(16, 128) == 8 and L.r3d_lfa_tile_points_for(16, 128, 8, 156) == 4
    assert L.r3d_lfa_tile_points_for(16, 128, 64, 2560) == 8 and L.r3d_lfa_tile_points_for(16, 256, 8, 39) == 4
    assert L.r3d_bn_set_fused(-1) == 0


def test_bench_roofline_picks_the_dominant_kernel_family():
    """bench.py: per-kernel event samples -> table (event-pair overhead removed) -> roofline entry of the kernel with
    the largest summed time, the two halves of an LFA block counted as one kernel, bound = the slower of FP32 / HBM."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.kernel_family("lfa_pool2_bwd[N=156,d=128]") == "lfa_pool_bwd"
    assert bench.kernel_family("lfa_pool1[N=39,d=256]") == "lfa_pool"
    assert bench.kernel_family("pw_gemm_fast[M=312,512->256]") == "pw_gemm_fast"

    class Ev:                                          # stands in for a pair of CUDA events
        def __init__(self, ms):
            self.ms = ms

        def elapsed_time(self, other):
            return other.ms - self.ms

    timers = {
        "lfa_pool1_bwd[N=156,d=128]": [(Ev(0.0), Ev(0.103), dict(flops=2.0e9, bytes=6.0e6))],
        "lfa_pool2_bwd[N=156,d=128]": [(Ev(0.0), Ev(0.113), dict(flops=2.2e9, bytes=1.1e7))],
        "bn_apply[M=20000,C=8]": [(Ev(0.0), Ev(0.009), dict(flops=6.4e5, bytes=1.3e6)) for _ in range(4)],
    }
    tab = bench.kernel_table(timers, overhead_ms=0.003)
    assert abs(tab["lfa_pool1_bwd[N=156,d=128]"]["ms_avg"] - 0.100) < 1e-9 and tab["bn_apply[M=20000,C=8]"]["launches"] == 4
    roof = bench.roofline_of(tab, dict(hbm_gbs=6548.5, source="test"), dict(ffma=71.0, ffma2=74.0))
    assert roof["kernel"] == "lfa_pool_bwd" and roof["launches"] == 2 and roof["bound"] == "fp32"
    assert abs(roof["achieved"] - 4.2e9 / 0.210e-3 * 1e-12) < 1e-6 and abs(roof["frac"] - roof["achieved"] / 71.0) < 1e-12
    # a memory-bound table: the bound flips to HBM
    tab2 = bench.kernel_table({"bn_apply[M=2000000,C=64]": [(Ev(0.0), Ev(0.5), dict(flops=5.1e8, bytes=1.0e9))]})
    roof2 = bench.roofline_of(tab2, dict(hbm_gbs=6548.5, source="test"), dict(ffma=71.0, ffma2=74.0))
    assert roof2["bound"] == "hbm" and abs(roof2["achieved"] - 2000.0) < 1e-6


def test_bench_line_is_short_and_round_trips():
    """The driver keeps only a short tail of stdout: the bench line must stay under 1 200 characters whatever the size
    of the per-kernel table, parse as JSON and carry every contract key; the bulky record goes to the side file."""
    import argparse
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test2", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    tab = {f"pw_gemm_fast[M={m},{a}->{b}]": dict(launches=3, ms_total=0.1, ms_avg=0.033, flops=1e9, bytes=1e7,
                                                 tflops=3.0, frac_fp32_peak=0.04, gbs=10.0, frac_hbm_peak=0.01)
           for m in range(40) for a in (64, 128) for b in (64, 128, 256)}                 # 240 rows ≈ 50 KB
    roof = dict(kernel="lfa_pool_bwd", launches=8, ms_avg=7.7, traffic=123456789.0, shapes=sorted(tab),
                algorithmic_flops_per_launch=1.0e11, algorithmic_bytes_per_launch=1.0e9, bound="fp32",
                achieved=28.612345678, peak=70.961234, unit="TFLOP/s", frac=0.40321234, peak_source="x" * 300)
    wl = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert bench.DEFAULT_WORKLOAD == "train40960" and wl["scaling"] == "strong" and wl["global_batch"] == 64
    args = argparse.Namespace(steps=20, warmup=5)
    line, side = bench.compose_line(
        metric="train_step_points_per_sec", value=17912345.678, unit="points/s", world=8, args=args,
        ms_per_step=146.31234567, wl=wl, name="train40960", n=40960, k=16, gbatch=64, batch=8, e2e_value=17812345.6,
        h2d=52428800, d2h=4, launches=6780, clocks=dict(sm_mhz=1965.0, sm_max_mhz=1965.0, reasons=["sw_power_cap"],
                                                        samples=33),
        roof=roof, cpu_base=dict(value=111500.123, unit="points/s", cores=16, kind="port", sample="y" * 400), tab=tab,
        fp32_peak=dict(ffma=70.9, ffma2=74.0), wall=3.2, graphed=True, eager_ms_per_step=150.0, extras={"a": 1})
    assert len(line) < bench.MAX_LINE_CHARS <= 1200 and "\n" not in line
    d = json.loads(line)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["value"] > 0 and d["n_gpus"] == 8 and d["config"]["workload"] == "train40960"
    assert set(d["e2e"]) == {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind"} <= set(d["cpu_baseline"])
    assert "kernels" not in d and len(side["kernels"]) == 240 and side["roofline"]["shapes"]
    assert json.loads(json.dumps(side))["extras"] == {"a": 1}


def test_reference_written_checkpoint_matches_the_schema():
    """tests/golden/ref_checkpoint.zip (written by the reference's Model.save): archive members, config keys and every
    tensor name / shape / dtype equal the schema the product's RandLANet builds (no GPU needed)."""
    import io
    import json
    import zipfile
    from oracle import network as onet
    modules = importlib.import_module("3d_recognizer_b200.modules")
    with zipfile.ZipFile(os.path.join(GOLDEN, "ref_checkpoint.zip")) as z:
        assert sorted(z.namelist()) == ["config", "model"]
        config = json.loads(z.read("config"))
        sd = torch.load(io.BytesIO(z.read("model")), map_location="cpu")
    st = modules.RandLANetSettings(**config)
    assert st.n_points == 2500 and st.n_neighbors == 32 and st.layer_sizes == [16, 64, 128, 256]
    schema = onet.state_dict_schema(config)
    assert list(sd) == list(schema)
    for k, (shape, dtype) in schema.items():
        assert tuple(sd[k].shape) == shape and sd[k].dtype == dtype, k


def test_settings_outside_the_fused_kernels():
    """Any n_neighbors in 1..64 and any layer size that is a multiple of 8 constructs (the fused kernels cover
    {16,...,256} x {16,32}; the rest runs in row form, tests/test_lfa_rows_gpu.py); what no kernel serves is refused when
    the network is built, with the supported sets in the message — not at the first forward (advisor finding, round 1)."""
    modules = importlib.import_module("3d_recognizer_b200.modules")
    dev = torch.device("cpu")
    base = dict(n_classes=2, n_points=4096, n_features=0)
    for kw in (dict(n_neighbors=24), dict(layer_sizes=[16, 64, 96, 256]), dict(n_neighbors=8), dict(layer_sizes=[8, 24])):
        modules.RandLANet(modules.RandLANetSettings(**dict(base, **kw)), dev)
    for kw in (dict(n_neighbors=65), dict(n_neighbors=0), dict(layer_sizes=[16, 20, 64])):
        with pytest.raises(ValueError, match="multiples of 8"):
            modules.RandLANet(modules.RandLANetSettings(**dict(base, **kw)), dev)
    modules.RandLANet(modules.RandLANetSettings(n_neighbors=32, **base), dev)


ROWS = {
    "k8_sizes_8_24_40_n1024": (dict(n_classes=2, n_points=1024, n_features=0, n_neighbors=8, layer_sizes=[8, 24, 40],
                                    knn="naive"), 2, 1024, 131),
    "k20_n1600": (dict(n_classes=2, n_points=1600, n_features=0, n_neighbors=20, knn="naive"), 2, 1600, 32),
    "k16_sizes_16_48_96_256_n2048": (dict(n_classes=2, n_points=2048, n_features=0, n_neighbors=16,
                                          layer_sizes=[16, 48, 96, 256], knn="naive"), 1, 2048, 34),
}


@pytest.mark.parametrize("name", list(ROWS))
def test_oracle_vs_reference_golden_outside_fused_shapes(name):
    """The oracle port against vectors written by the reference's own modules for settings outside the fused kernels'
    template lists (oracle/make_golden.py E2E_ROWS): eval logits, train logits and loss — the checker of
    tests/test_lfa_rows_gpu.py is pinned on these shapes too."""
    g = np.load(os.path.join(GOLDEN, "e2e_rows_golden.npz"))
    st, B, N, seed = ROWS[name]
    sd = onet.synth_state_dict(st, seed)
    x = torch.from_numpy(make_input(B, N, 0, seed))
    labels = torch.from_numpy(np.random.RandomState(seed).randint(0, 2, (B, N)))
    perm = g[f"{name}/perm"].astype(np.int64)
    with torch.no_grad():
        ev = onet.forward({k: v.clone() for k, v in sd.items()}, st, x, perm)
        tr = onet.forward({k: v.clone() for k, v in sd.items()}, st, x, perm, training=True, dropout_p=0.0)
    for got, key in ((ev, "eval_logits"), (tr, "train_logits")):
        ref = torch.from_numpy(g[f"{name}/{key}"])
        assert float((got - ref).abs().max() / ref.abs().max()) < 1e-5, key
    assert abs(float(onet.dice_loss(tr, labels)) - float(g[f"{name}/train_loss"])) < 1e-5
