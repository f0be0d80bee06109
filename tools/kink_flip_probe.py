"""Kink-flip probe for the 16 384-point training step (tests/test_api_gpu.py::test_train_step_16k_vs_oracle_port):
per LFA block, the number of output pre-activations whose LeakyReLU branch differs between the fused fp32 kernels, the
fp32 tensor-op composition and an fp64 evaluation of the same block on the same inputs.
usage: python tools/kink_flip_probe.py"""
import importlib, sys, os, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from oracle import network as onet
modules = importlib.import_module("3d_recognizer_b200.modules")
L = importlib.import_module("3d_recognizer_b200._cabi").lib()
from test_api_gpu import make_input
st = dict(n_classes=2, n_points=16384, n_features=0, n_neighbors=16, knn="kdtree")
sd = onet.synth_state_dict(st, 41)
x = torch.from_numpy(make_input(1, 16384, 0, 41))
labels = torch.from_numpy(np.random.RandomState(41).randint(0, 2, (1, 16384)))
sd_ref = {k: v.clone() for k, v in sd.items()}
leaves = {}
for k, v in sd_ref.items():
    if v.is_floating_point() and "running" not in k:
        v.requires_grad_(True); leaves[k] = v
np.random.seed(78)
ref_logits = onet.forward(sd_ref, st, x, training=True, dropout_p=0.0)
onet.dice_loss(ref_logits, labels).backward()
import time
t0=time.time()
sd64 = {k: (v.clone().double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
leaves64 = {}
for k, v in sd64.items():
    if v.is_floating_point() and "running" not in k:
        v.requires_grad_(True); leaves64[k] = v
np.random.seed(78)
l64 = onet.forward(sd64, st, x.double(), training=True, dropout_p=0.0)
onet.dice_loss(l64, labels).backward()
print("fp64 arbiter", time.time()-t0, "s")
def l2(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return float((a.double()-b.double()).norm()/max(float(b.double().norm()),1e-30))
engine = importlib.import_module("3d_recognizer_b200.engine")
import copy
net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
net.load_state_dict(sd); net.train(); net.fc_end[2].p = 0.0
rec = []
def wrapped(l, xyz, feat):
    y = engine.lfa_block_fused(l, xyz, feat); y.retain_grad(); rec.append((l, xyz.detach().clone(), feat.detach().clone(), y)); return y
engine.LFA_IMPL = wrapped
np.random.seed(78)
logits = net(x.cuda())
onet.dice_loss(logits, labels.cuda()).backward()
real_lrelu = F_leaky = torch.nn.functional.leaky_relu
pre = {}
def spy(x, slope=0.01, *a, **k):
    pre.setdefault(cur[0], []).append(x.detach().double().cpu())
    return real_lrelu(x, slope, *a, **k)
cur = [None]
import torch.nn.functional as F
engine.F.leaky_relu = spy
for lvl, (l, xyz, feat, y) in enumerate(rec):
    for name, fn, dt in (("fused", engine.lfa_block_fused, torch.float32), ("ops", engine.lfa_block, torch.float32), ("fp64", engine.lfa_block, torch.float64)):
        cur[0] = (lvl, name)
        lc = copy.deepcopy(l).to(dt)
        with torch.no_grad():
            fn(lc, xyz.to(dt), feat.detach().clone().to(dt))
    for name in ("fused", "ops"):
        msg = []
        a, b = pre[(lvl, name)][-1], pre[(lvl, "fp64")][-1]
        mism = (a > 0) != (b > 0)
        msg.append((int(mism.sum()), float(b[mism].abs().max()) if mism.any() else 0.0, float((a - b).abs().max()), a.numel()))
        print("level", lvl, name, "leaky_relu calls (index, sign mismatches vs fp64, max |s64| at mismatch, max |s - s64|):", msg)
