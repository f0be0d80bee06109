"""Kink-flip probe for the 16 384-point training step (tests/test_api_gpu.py::test_train_step_16k_vs_oracle_port).

Three measurements on one seeded step (N=16384, B=1, K=16, synthetic weights):
 1. whole-step parameter gradients of the fused-kernel path and of the fp32 CPU port, each against an fp64 evaluation
    of the port (relative L2 per tensor: median and maximum);
 2. per LFA block, on the block's real inputs: forward error of the fused kernels and of the fp32 tensor-op composition
    against an fp64 run of the composition;
 3. per LFA block: the number of output pre-activations whose LeakyReLU branch differs from the fp64 run, and how close
    to zero those pre-activations are.
A single flipped element (|s| ~ 1e-7) moves every gradient upstream of it by 1e-4 .. 5e-4 in relative L2 while all
forward values agree to fp32 round-off: that is what the test's multi-seed criterion allows for.
usage: python tools/kink_flip_probe.py [seed]   (output: profiles/r01_kink_flip_probe.txt)"""
import copy
import importlib
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import network as onet           # noqa: E402  (tools may use the oracle as a checker)
from test_api_gpu import make_input          # noqa: E402

modules = importlib.import_module("3d_recognizer_b200.modules")
engine = importlib.import_module("3d_recognizer_b200.engine")


def rel_l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def port_gradients(sd, st, x, labels, dtype):
    sd = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    leaves = {}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
            leaves[k] = v
    np.random.seed(78)
    logits = onet.forward(sd, st, x.to(dtype), training=True, dropout_p=0.0)
    onet.dice_loss(logits, labels).backward()
    return logits.detach(), {k: v.grad for k, v in leaves.items()}


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 41
    st = dict(n_classes=2, n_points=16384, n_features=0, n_neighbors=16, knn="kdtree")
    sd = onet.synth_state_dict(st, seed)
    x = torch.from_numpy(make_input(1, 16384, 0, seed))
    labels = torch.from_numpy(np.random.RandomState(seed).randint(0, 2, (1, 16384)))
    logits32, g32 = port_gradients(sd, st, x, labels, torch.float32)
    logits64, g64 = port_gradients(sd, st, x, labels, torch.float64)

    # ---- 1. whole step on the fused kernels, recording every LFA block's inputs and output
    rec = []

    def recording(lfa, xyz, feat, *extra):
        y = engine.lfa_block_fused(lfa, xyz, feat, *extra)
        rec.append((lfa, xyz.detach().clone(), feat.detach().clone()))
        return y

    engine.LFA_IMPL = recording
    net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
    net.load_state_dict(sd)
    net.train()
    net.fc_end[2].p = 0.0
    np.random.seed(78)
    logits = net(x.cuda())
    onet.dice_loss(logits, labels.cuda()).backward()
    engine.LFA_IMPL = engine.lfa_block_fused
    ours = {k: rel_l2(p.grad, g64[k]) for k, p in net.named_parameters() if p.grad is not None}
    port = {k: rel_l2(g32[k], g64[k]) for k in ours}
    print(f"logits vs fp64: kernels {rel_l2(logits, logits64):.2e}, fp32 port {rel_l2(logits32, logits64):.2e}")
    print(f"gradients vs fp64 (relative L2 per tensor): kernels median {statistics.median(ours.values()):.2e} "
          f"max {max(ours.values()):.2e};  fp32 port median {statistics.median(port.values()):.2e} "
          f"max {max(port.values()):.2e}")

    # ---- 2./3. per block, on its real inputs
    pre = {}
    tag = [None]
    real = torch.nn.functional.leaky_relu

    def spy(t, slope=0.01, *a, **k):
        pre.setdefault(tag[0], []).append(t.detach().double().cpu())
        return real(t, slope, *a, **k)

    engine.F.leaky_relu = spy
    try:
        for lvl, (lfa, xyz, feat) in enumerate(rec):
            outs = {}
            for name, fn, dt in (("fused", engine.lfa_block_fused, torch.float32),
                                 ("ops", engine.lfa_block, torch.float32), ("fp64", engine.lfa_block, torch.float64)):
                tag[0] = (lvl, name)
                with torch.no_grad():
                    outs[name] = fn(copy.deepcopy(lfa).to(dt), xyz.to(dt), feat.to(dt))
            line = [f"level {lvl} N={xyz.shape[1]}"]
            for name in ("fused", "ops"):
                s, s64 = pre[(lvl, name)][-1], pre[(lvl, "fp64")][-1]      # the block's final LeakyReLU input
                flips = (s > 0) != (s64 > 0)
                line.append(f"{name}: forward rel-L2 {rel_l2(outs[name], outs['fp64']):.2e}, branch flips "
                            f"{int(flips.sum())} of {s.numel()}"
                            + (f" (max |s| there {float(s64[flips].abs().max()):.1e})" if flips.any() else ""))
            print(" | ".join(line))
    finally:
        engine.F.leaky_relu = real


if __name__ == "__main__":
    main()
