import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import network as onet
from test_forward_gpu import E2E, make_input
modules = importlib.import_module("3d_recognizer_b200.modules"); engine = importlib.import_module("3d_recognizer_b200.engine")
name = sys.argv[1] if len(sys.argv) > 1 else "k16_n1024"
g = np.load("tests/golden/e2e_golden.npz")
st, B, N, seed = E2E[name]
res = {}
for use in (True, False):
    engine.USE_POINTWISE_KERNELS = use
    net = modules.RandLANet(modules.RandLANetSettings(**st), torch.device("cuda"))
    net.load_state_dict(onet.synth_state_dict(st, seed))
    x = torch.from_numpy(make_input(B, N, st["n_features"], seed)).cuda()
    labels = torch.from_numpy(np.random.RandomState(seed).randint(0, st["n_classes"], (B, N))).cuda()
    net.train(); net.fc_end[2].p = 0.0
    np.random.seed(seed)
    logits = net(x)
    loss = onet.dice_loss(logits, labels); net.zero_grad(); loss.backward()
    res[use] = {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()}
    got = {k: onet.grad_fixture_view(p.grad) for k, p in net.named_parameters()}
    refg = {k: torch.from_numpy(g[f"{name}/grad/{k}"]) for k in got}
    print("kernels" if use else "torch  ", "worst", onet.grad_parity(got, refg))
k = "encoder.2.shortcut.batch_norm.bias"
a, b = res[True][k], res[False][k]
d = (a - b).abs()
print("max|ref|", float(b.abs().max()), "n elems off > 1e-4 rel:", int((d > 1e-4 * b.abs().max()).sum()), "of", d.numel(), "argmax", int(d.argmax()), float(d.max()))
srt = torch.sort(d, descending=True).values[:5]; print("top diffs", srt.tolist())
