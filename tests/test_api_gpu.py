"""GPU tests of the reference-facing module API beyond the golden cases: KNN / UpSampler modules, error
behaviour, a training step on a configuration the goldens do not cover (K=32, point features, 3 classes)
against the oracle port run on the CPU, and CUDA-graph replay of the training step against eager execution."""
import copy
import importlib

import numpy as np
import pytest
import torch

from oracle import network as onet
from oracle.knn import knn_exact
from test_forward_gpu import make_input, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def mods():
    return (importlib.import_module("3d_recognizer_b200.modules"), importlib.import_module("3d_recognizer_b200.engine"),
            importlib.import_module("3d_recognizer_b200.model"))


def test_knn_module_contract(mods):
    """KNN.forward (modules.py:107-150): int64 indices, NON-squared distances, any back-end name, ValueError
    on an unknown one; inputs may live on the CPU (the reference moves them itself)."""
    modules, _, _ = mods
    dev = torch.device("cuda")
    knn = modules.KNN(dev)
    rng = np.random.RandomState(0)
    s, q = rng.rand(2, 900, 3).astype(np.float32), rng.rand(2, 300, 3).astype(np.float32)
    oi, od = knn_exact(s, q, 8)
    for approach in ("kdtree", "approximate", "naive"):
        idx, dist = knn(torch.from_numpy(s), torch.from_numpy(q), 8, approach)
        assert idx.dtype == torch.int64 and dist.dtype == torch.float32 and idx.is_cuda
        assert np.array_equal(idx.cpu().numpy(), oi)
        assert np.array_equal(dist.cpu().numpy(), np.sqrt(od))
    with pytest.raises(ValueError):
        knn(torch.from_numpy(s), torch.from_numpy(q), 8, "ball_tree")


@pytest.mark.parametrize("approach", ["none", "nni", "nna", "idw", "isdw"])
def test_upsampler_vs_oracle(mods, approach):
    """UpSampler (modules.py:328-456) on (B,F,N1,1) features: all modes vs the oracle restatement."""
    modules, _, _ = mods
    rng = np.random.RandomState(4)
    feat = torch.from_numpy(rng.randn(2, 5, 700, 1).astype(np.float32))
    xyz = torch.from_numpy(rng.rand(2, 700, 3).astype(np.float32))
    xyz_up = torch.from_numpy(rng.rand(2, 3000, 3).astype(np.float32))
    up = modules.UpSampler(approach, torch.device("cuda"))
    got = up(feat.cuda(), xyz.cuda(), xyz_up.cuda()).cpu()
    ref = onet.upsample(approach, feat, xyz, xyz_up)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-5
    with pytest.raises(ValueError):
        modules.UpSampler("cubic", torch.device("cuda"))(feat.cuda(), xyz.cuda(), xyz_up.cuda())


def test_forward_asserts_on_gpu(mods):
    modules, _, _ = mods
    net = modules.RandLANet(modules.RandLANetSettings(n_classes=2, n_neighbors=16), torch.device("cuda"))
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 1024, 4, device="cuda"))            # dim != 3 + F
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 1023, 3, device="cuda"))            # fewer than max(16*64, 2*256) points
    assert net.device.type == "cuda" and net.settings.n_neighbors == 16


def test_train_step_k32_features_vs_oracle_port(mods):
    """K=32, 2 point features, 3 classes, N=2560, B=2 — not among the golden cases: logits, loss and running statistics
    of one training step vs the oracle port (pinned to the reference) on the CPU, one run, strict tolerance.
    Gradients: every tensor within 2 % relative L2 and >= 99 % of the entries within 3e-3 of their tensor's maximum.
    The bottleneck BatchNorm of this configuration normalises over 20 rows, which amplifies fp32 round-off enough that
    ReLU branches flip between any two evaluations — the oracle port's own runs included (its threaded reductions are
    not bit-reproducible); the entry-by-entry bar with pinned branches is held in tests/test_kink_pinned_gpu.py on
    configurations whose batch statistics are not degenerate."""
    modules, _, _ = mods
    st = dict(n_classes=3, n_points=2560, n_features=2, n_neighbors=32, knn="kdtree")
    sd = onet.synth_state_dict(st, 31)
    x = torch.from_numpy(make_input(2, 2560, 2, 31))
    labels = torch.from_numpy(np.random.RandomState(31).randint(0, 3, (2, 2560)))
    sd_ref = {k: v.clone() for k, v in sd.items()}
    leaves = {}
    for k, v in sd_ref.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
            leaves[k] = v
    np.random.seed(77)
    ref_logits = onet.forward(sd_ref, st, x, training=True, dropout_p=0.0)
    ref_loss = onet.dice_loss(ref_logits, labels)
    ref_loss.backward()
    ref_logits, ref_loss = ref_logits.detach(), ref_loss.detach()
    net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
    net.load_state_dict(sd)
    net.train()
    net.fc_end[2].p = 0.0
    np.random.seed(77)
    logits = net(x.cuda())
    loss = onet.dice_loss(logits, labels.cuda())
    loss.backward()
    assert rel_err(logits.detach().cpu(), ref_logits) < TOL
    assert abs(loss.item() - ref_loss.item()) < 1e-5
    got = {k: p.grad for k, p in net.named_parameters()}
    refg = {k: v.grad for k, v in leaves.items()}
    worst_l2, wname = onet.grad_parity_l2(got, refg)
    assert worst_l2 < 2e-2, (worst_l2, wname)
    assert onet.grad_parity_fraction(got, refg, 3e-3) >= 0.99
    for k, v in net.state_dict().items():
        if "running" in k:
            assert torch.allclose(v.cpu(), sd_ref[k].detach(), rtol=1e-4, atol=1e-5), k


def test_train_step_16k_vs_oracle_port(mods):
    """N=16384, B=1, K=16 (four times the golden size; the level-0 KNN runs on the uniform grid, the per-point layers
    on the large-row kernels): logits, loss and gradients of one step vs the oracle port on the CPU.

    Beyond the golden sizes a flipped activation branch (oracle.network.grad_parity) is the rule, not the exception:
    of the 4 M pre-activations of this step a few sit within fp32 round-off of zero (measured with
    tools/kink_flip_probe.py: one element of the level-2 block output at |s| = 8e-8 takes the other LeakyReLU branch than
    the fp64 evaluation), and ONE such flip moves every gradient upstream of it by 1e-4 .. 5e-4 in relative L2
    (BatchNorm's batch means spread it over the whole channel).  A defect, in contrast, shows on every input.  So:
    every seed must keep logits and loss at the strict tolerance and every gradient tensor within 2 % relative L2;
    and at least one of up to five seeds must have >= 99.5 % of all 1.3 M gradient entries within 1e-4."""
    modules, _, _ = mods
    st = dict(n_classes=2, n_points=16384, n_features=0, n_neighbors=16, knn="kdtree")
    history = []
    for seed in (41, 42, 43, 44, 45):
        sd = onet.synth_state_dict(st, seed)
        x = torch.from_numpy(make_input(1, 16384, 0, seed))
        labels = torch.from_numpy(np.random.RandomState(seed).randint(0, 2, (1, 16384)))
        sd_ref = {k: v.clone() for k, v in sd.items()}
        leaves = {}
        for k, v in sd_ref.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
                leaves[k] = v
        np.random.seed(78)
        ref_logits = onet.forward(sd_ref, st, x, training=True, dropout_p=0.0)
        ref_loss = onet.dice_loss(ref_logits, labels)
        ref_loss.backward()
        net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
        net.load_state_dict(sd)
        net.train()
        net.fc_end[2].p = 0.0
        np.random.seed(78)
        logits = net(x.cuda())
        loss = onet.dice_loss(logits, labels.cuda())
        loss.backward()
        assert rel_err(logits.detach().cpu(), ref_logits.detach()) < TOL, seed
        assert abs(loss.item() - ref_loss.item()) < 1e-5, seed
        got = {k: p.grad for k, p in net.named_parameters()}
        refg = {k: v.grad for k, v in leaves.items()}
        worst_l2, wname_l2 = onet.grad_parity_l2(got, refg)
        assert worst_l2 < 2e-2, (seed, worst_l2, wname_l2)
        frac = onet.grad_parity_fraction(got, refg, TOL)
        if frac >= 0.995:
            return
        history.append((seed, frac, worst_l2, wname_l2))
    raise AssertionError(("no seed met the strict gradient bar", history))


@pytest.mark.parametrize("extra", [{}, dict(n_neighbors=8, layer_sizes=[16, 24, 64])], ids=["fused", "row_form"])
def test_graphed_train_step_matches_eager(mods, extra):
    """GraphedTrainStep (CUDA-graph replay) follows the same loss trajectory as eager Model.train_step from the
    same weights, data and numpy seed (Dropout off so that both paths are deterministic functions of those).
    ``row_form``: settings outside the fused kernels' template lists (csrc/lfa_rows.cu) capture and replay too."""
    modules, _, model_mod = mods
    syn = importlib.import_module("3d_recognizer_b200.synthetic")
    st = modules.RandLANetSettings(**dict(dict(n_classes=2, n_points=1024, n_features=0, n_neighbors=16, knn="naive"), **extra))
    torch.manual_seed(3)
    a = model_mod.Model(st)
    b = model_mod.Model(st, weights=copy.deepcopy(a.module.state_dict()))
    for m in (a, b):
        m.module.fc_end[2].p = 0.0
    x, y = syn.fingertip_batch(5, 4, 1024, n_raw=20000)
    x, y = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    opt_a = a.make_optimizer(1e-3)
    opt_b = b.make_optimizer(1e-3, capturable=True)
    np.random.seed(11)
    gstep = model_mod.GraphedTrainStep(b, opt_b, (4, 1024, 3), warmup=0)
    np.random.seed(12)
    eager = [float(a.train_step(x, y, opt_a)) for _ in range(4)]
    np.random.seed(12)
    graphed = [float(gstep(x, y)) for _ in range(4)]
    assert np.allclose(eager, graphed, rtol=2e-3, atol=1e-5), (eager, graphed)
    pa = torch.cat([p.detach().flatten() for p in a.module.parameters()])
    pb = torch.cat([p.detach().flatten() for p in b.module.parameters()])
    # Adam normalises every gradient to a step of about lr, also where the gradient is pure round-off, so after four
    # steps two correct runs may differ by up to ~2 * 4 * lr in such parameters; the loss trajectory above is the check
    assert float((pa - pb).abs().max()) < 2 * 4 * 1e-3 + 1e-4
    # the reference's StepLR (trainer.py:82) must reach the replayed optimiser: lr -> 0 freezes the parameters
    sched = torch.optim.lr_scheduler.StepLR(opt_b, step_size=1, gamma=0.0)
    sched.step()
    assert float(opt_b.param_groups[0]["lr"]) == 0.0
    before = torch.cat([p.detach().flatten() for p in b.module.parameters()]).clone()
    gstep(x, y)
    after = torch.cat([p.detach().flatten() for p in b.module.parameters()])
    assert torch.equal(before, after)


@pytest.mark.parametrize("bound", [False, True])
def test_flat_adam_matches_torch_adam(mods, bound):
    """model.FlatAdam (one fused launch over a flat buffer the parameters are views of) == torch.optim.Adam on the
    same gradients, also with the data-parallel flat gradient buffer bound, and with parameters that get no gradient."""
    modules, _, model_mod = mods
    parallel = importlib.import_module("3d_recognizer_b200.parallel")
    torch.manual_seed(1)
    net_a = modules.SharedMLP(8, 16, activation=torch.nn.ReLU()).cuda()
    net_b = copy.deepcopy(net_a)
    opt_a = torch.optim.Adam(net_a.parameters(), lr=3e-3)
    opt_b = model_mod.FlatAdam(net_b, 3e-3, capturable=False)
    flat = parallel.FlatGradients(net_b) if bound else None
    if bound:
        opt_b.bind_flat_gradients(flat.flat)
    assert all(pb.data_ptr() != pa.data_ptr() for pa, pb in zip(net_a.parameters(), net_b.parameters()))
    for step in range(3):
        g = torch.Generator(device="cuda").manual_seed(step)
        grads = [torch.randn(p.shape, device="cuda", generator=g) for p in net_a.parameters()]
        if bound:
            flat.zero()
        else:
            opt_b.zero_grad()
        opt_a.zero_grad()
        for i, (pa, pb, gr) in enumerate(zip(net_a.parameters(), net_b.parameters(), grads)):
            if i == 1:
                continue                         # the conv bias gets no gradient (None): must stay untouched
            pa.grad = gr.clone()
            pb.grad = gr.clone()
        if bound:
            flat.rebind()
        opt_a.step()
        opt_b.step()
    for (k, pa), (_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-7), k
    sd = net_b.state_dict()                          # checkpoints still see ordinary tensors
    net_c = modules.SharedMLP(8, 16, activation=torch.nn.ReLU()).cuda()
    net_c.load_state_dict(sd)
    assert all(torch.equal(pb, pc) for pb, pc in zip(net_b.parameters(), net_c.parameters()))


def test_eval_forward_65536_vs_oracle(mods):
    """BASELINE config 3's middle size, one cloud of 65 536 points (uniform-grid KNN at level 0, tensor-core LFA kernels
    at levels 1-3, tcgen05 per-point layers): eval logits against the oracle port on the CPU."""
    modules, _, _ = mods
    st = dict(n_classes=2, n_points=65536, n_features=0, n_neighbors=16, knn="kdtree")
    sd = onet.synth_state_dict(st, 5)
    net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
    net.load_state_dict(sd)
    net.eval()
    x = torch.from_numpy(make_input(1, 65536, 0, 5))
    np.random.seed(1)
    with torch.no_grad():
        got = net(x.cuda()).cpu()
    np.random.seed(1)
    with torch.no_grad():
        ref = onet.forward({k: v.clone() for k, v in sd.items()}, st, x, training=False)
    assert rel_err(got, ref) < TOL
    importlib.import_module("3d_recognizer_b200.ops").check_tc_status(torch.device("cuda"))


def test_train_step_40960x2_vs_oracle_port(mods):
    """BASELINE config 4's cloud size (40 960 points, two clouds): train-mode logits, loss and BatchNorm running
    statistics at the strict bar, one run, no retries.  Gradients: every tensor within 1 % relative L2 and >= 99.9 % of
    the 1.3 M entries within 3e-3 of their tensor's maximum — at this size ~130 pre-activations sit within fp32
    round-off of a ReLU kink, a dozen of them end up on the other branch than in the oracle's own (not bit-reproducible)
    run, and each such toggle moves whole tensors by 1e-4..1e-3; pinning them one oracle run at a time
    (tests/test_kink_pinned_gpu.py does exactly that up to 16 384 points, where EVERY entry then agrees within 1e-4)
    would take a quarter of an hour here."""
    modules, _, _ = mods
    N, B, seed = 40960, 2, 7
    st = dict(n_classes=2, n_points=N, n_features=0, n_neighbors=16, knn="kdtree")
    sd = onet.synth_state_dict(st, seed)
    x = torch.from_numpy(make_input(B, N, 0, seed))
    labels = torch.from_numpy(np.random.RandomState(seed).randint(0, 2, (B, N)))
    sd_ref = {k: v.clone() for k, v in sd.items()}
    leaves = {}
    for k, v in sd_ref.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
            leaves[k] = v
    np.random.seed(78)
    ref_logits = onet.forward(sd_ref, st, x, training=True, dropout_p=0.0)
    ref_loss = onet.dice_loss(ref_logits, labels)
    ref_loss.backward()
    net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
    net.load_state_dict(sd)
    net.train()
    net.fc_end[2].p = 0.0
    np.random.seed(78)
    logits = net(x.cuda())
    loss = onet.dice_loss(logits, labels.cuda())
    loss.backward()
    assert rel_err(logits.detach().cpu(), ref_logits.detach()) < TOL
    assert abs(loss.item() - ref_loss.item()) < 1e-5
    for k, v in net.state_dict().items():
        if "running" in k:
            assert torch.allclose(v.cpu(), sd_ref[k].detach(), rtol=1e-4, atol=1e-5), k
    got = {k: p.grad for k, p in net.named_parameters()}
    refg = {k: v.grad for k, v in leaves.items()}
    worst_l2, wname = onet.grad_parity_l2(got, refg)
    fracs = {t: onet.grad_parity_fraction(got, refg, t) for t in (1e-4, 3e-4, 1e-3, 3e-3)}
    print(f"40960x2: worst tensor rel-L2 {worst_l2:.2e} ({wname}); fraction of entries within tol: {fracs}")
    assert worst_l2 < 1e-2, (worst_l2, wname)
    assert fracs[3e-3] >= 0.999, fracs
    importlib.import_module("3d_recognizer_b200.ops").check_tc_status(torch.device("cuda"))


def test_graphed_predict_matches_eager():
    """Model.predict replays the network's eval forward from a CUDA graph (model.GraphedEvalForward): same confidences
    as the eager launches on the same numpy stream, for frames of different sizes (the graph is keyed by the sampled
    shape, not by the frame), and the graph is rebuilt when a weight changes."""
    model_mod = importlib.import_module("3d_recognizer_b200.model")
    modules = importlib.import_module("3d_recognizer_b200.modules")
    st = dict(n_classes=2, n_points=2500, n_features=0, n_neighbors=32, knn="naive")
    m = model_mod.Model(modules.RandLANetSettings(**st), weights=onet.synth_state_dict(st, 21))
    rng = np.random.RandomState(4)

    def both(frame):
        m.use_cuda_graphs = True
        np.random.seed(11)
        a = m.predict(frame)
        m.use_cuda_graphs = False
        np.random.seed(11)
        b = m.predict(frame)
        m.use_cuda_graphs = True
        return a, b

    for n in (30, 7000, 50000, 7000):                     # 30: the warm-up call of predict.py:22-24
        a, b = both(rng.rand(n, 3).astype(np.float32))
        assert a.shape == (2, n) and np.array_equal(a, b)
    assert len(m._eval_graphs) == 1
    first = next(iter(m._eval_graphs.values()))
    with torch.no_grad():
        m.module.fc_end[3].conv.bias.add_(0.25)           # a training step would do the same to every parameter
    a, b = both(rng.rand(9000, 3).astype(np.float32))
    assert np.array_equal(a, b)
    assert next(iter(m._eval_graphs.values())) is not first


def test_model_infer_matches_module_forward():
    """Model.infer (eval forward replayed from a CUDA graph per input shape) == module(x) on the same numpy stream."""
    model_mod = importlib.import_module("3d_recognizer_b200.model")
    modules = importlib.import_module("3d_recognizer_b200.modules")
    st = dict(n_classes=2, n_points=4096, n_features=0, n_neighbors=16, knn="naive")
    m = model_mod.Model(modules.RandLANetSettings(**st), weights=onet.synth_state_dict(st, 7))
    for B, N in ((2, 4096), (3, 2048), (2, 4096)):
        x = torch.from_numpy(make_input(B, N, 0, B + N)).cuda()
        np.random.seed(5)
        a = m.infer(x)
        np.random.seed(5)
        with torch.no_grad():
            b = m.module(x)
        assert a.shape == (B, 2, N) and torch.equal(a, b)
    assert len(m._eval_graphs) == 2


@pytest.mark.parametrize("n_in,d,K", [(8, 16, 16), (16, 24, 8)], ids=["fused", "row_form"])
def test_lfa_module_forward_runs_the_kernels(mods, n_in, d, K):
    """LocalFeatureAggregation.forward — the reference's sub-module call signature, (B,n_in,N,1) in, (B,2d,N,1) out
    (modules.py:298-325) — runs the sm_100a kernels for fp32 CUDA tensors (fused where built, row form elsewhere) and
    agrees with the tensor-op composition; an unknown KNN approach raises like the reference."""
    modules, engine, _ = mods
    dev = torch.device("cuda")
    torch.manual_seed(n_in + d)
    lfa = modules.LocalFeatureAggregation(n_in, d, K, dev).to(dev).eval()
    with torch.no_grad():
        for m in lfa.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    xyz = torch.rand(2, 700, 3, device=dev)
    x = torch.randn(2, n_in, 700, 1, device=dev)
    with torch.no_grad():
        y = lfa(xyz, x, "naive")
        ref = engine.lfa_block(lfa.double(), xyz.double(), x.double().squeeze(-1).transpose(1, 2)).transpose(1, 2).unsqueeze(-1)
    assert y.shape == (2, 2 * d, 700, 1)
    assert rel_err(y, ref.float()) < 1e-5
    with pytest.raises(ValueError):
        lfa.float()(xyz, x, "octree")
