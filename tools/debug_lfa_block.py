import importlib, sys, os, copy
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
engine = importlib.import_module("3d_recognizer_b200.engine")
ops = importlib.import_module("3d_recognizer_b200.ops")
modules = importlib.import_module("3d_recognizer_b200.modules")
E = engine

def fused(lfa, xyz, feat, keep):
    K = lfa._n_neighbors
    idx = ops.knn(xyz, xyz, K, idx64=False, idx32=True, dist=False)["idx32"]
    f = E.shared_mlp(lfa.mlp1, feat); keep["f"] = f
    w1 = lfa.mlp_rpe1.conv.weight.view(-1, 10)
    w2 = lfa.mlp_rpe2.conv.weight.view(w1.shape[0], w1.shape[0])
    d = 2 * w1.shape[0]
    m = ops.lfa_moments(0, xyz, idx, d)
    count = float(xyz.shape[0] * xyz.shape[1] * K)
    mu = m[10, :10] / count
    cov = m[:10, :10] / count - torch.outer(mu, mu)
    a1, c1 = E._bn_affine_from_moments(lfa.mlp_rpe1, mu, cov, count); keep["a1"] = a1; keep["c1"] = c1
    ws1, ws2 = lfa.pool1.score_fn[0].weight, lfa.pool2.score_fn[0].weight
    pooled1 = E._LfaPoolFn.apply(1, xyz, idx, f, w1, a1, c1, None, None, None, ws1); keep["pooled1"] = pooled1
    p1 = E.shared_mlp(lfa.pool1.mlp, pooled1); keep["p1"] = p1
    s_r1, m_r1 = E._R1MomentsFn.apply(xyz, idx, w1, a1, c1)
    mu_r = s_r1 / count
    cov_r = m_r1 / count - torch.outer(mu_r, mu_r)
    a2, c2 = E._bn_affine_from_moments(lfa.mlp_rpe2, mu_r, cov_r, count); keep["a2"] = a2; keep["c2"] = c2
    pooled2 = E._LfaPoolFn.apply(2, xyz, idx, p1, w1, a1, c1, w2, a2, c2, ws2); keep["pooled2"] = pooled2
    p2 = E.shared_mlp(lfa.pool2.mlp, pooled2); keep["p2"] = p2
    return F.leaky_relu(E.shared_mlp(lfa.mlp2, p2) + E.shared_mlp(lfa.shortcut, feat), 0.01)

def plain(lfa, xyz, feat, keep):
    nn_ = ops.knn(xyz, xyz, lfa._n_neighbors, idx64=True, dist=True)
    idx, dist = nn_["idx64"], nn_["dist"]
    f = E.shared_mlp(lfa.mlp1, feat); keep["f"] = f
    z1 = F.linear(E.relative_position_encoding(xyz, idx, dist), E.conv_weight_2d(lfa.mlp_rpe1), lfa.mlp_rpe1.conv.bias)
    r1 = F.relu(E.batch_norm_lastdim(lfa.mlp_rpe1.batch_norm, z1))
    pooled1 = (F.softmax(F.linear(torch.cat((r1, E.gather_points(f, idx)), dim=-1), lfa.pool1.score_fn[0].weight), dim=2) * torch.cat((r1, E.gather_points(f, idx)), dim=-1)).sum(2); keep["pooled1"] = pooled1
    p1 = E.shared_mlp(lfa.pool1.mlp, pooled1); keep["p1"] = p1
    r2 = E.shared_mlp(lfa.mlp_rpe2, r1)
    x2 = torch.cat((r2, E.gather_points(p1, idx)), dim=-1)
    pooled2 = (F.softmax(F.linear(x2, lfa.pool2.score_fn[0].weight), dim=2) * x2).sum(2); keep["pooled2"] = pooled2
    p2 = E.shared_mlp(lfa.pool2.mlp, pooled2); keep["p2"] = p2
    return F.leaky_relu(E.shared_mlp(lfa.mlp2, p2) + E.shared_mlp(lfa.shortcut, feat), 0.01)

n_in, d, K, N, B = 128, 128, 16, 300, 2
dev = torch.device("cuda")
torch.manual_seed(n_in + d + K)
lfa_a = modules.LocalFeatureAggregation(n_in, d, K, dev).to(dev)
with torch.no_grad():
    for m in lfa_a.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.uniform_(0.7, 1.3); m.bias.normal_(0, 0.1); m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
lfa_c = copy.deepcopy(lfa_a).double()
lfa_a.train(); lfa_c.train()
xyz = torch.rand(B, N, 3, device=dev); x = torch.randn(B, N, n_in, device=dev); gout = torch.randn(B, N, 2 * d, device=dev)
xa = x.clone().requires_grad_(True); xc = x.double().requires_grad_(True)
ka, kc = {}, {}
ya = fused(lfa_a, xyz, xa, ka); yc = plain(lfa_c, xyz.double(), xc, kc)
for k in ka: ka[k].retain_grad()
for k in kc: kc[k].retain_grad()
(ya * gout).sum().backward(); (yc * gout.double()).sum().backward()
rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
print("out", rel(ya.detach(), yc.detach()))
for k in kc:
    print(k, "val", rel(ka[k].detach(), kc[k].detach()), "grad", rel(ka[k].grad, kc[k].grad))
print("x grad", rel(xa.grad, xc.grad))
for (k, pa), (_, pc) in zip(lfa_a.named_parameters(), lfa_c.named_parameters()):
    print(k, rel(pa.grad, pc.grad) if pc.grad.abs().max() > 0 else "zero")

print("---- tail-only check")
for name, fn, lfa, xin, dt in (("fused", fused, lfa_a, x, torch.float32), ("plain32", plain, copy.deepcopy(lfa_a), x, torch.float32)):
    lfa.train()
    xi = xin.clone().requires_grad_(True)
    keep = {}
    y = fn(lfa, xyz, xi, keep)
    gp2, = torch.autograd.grad(y, keep["p2"], gout, retain_graph=True)
    # independent recomputation of the tail from detached tensors, fp64
    l64 = copy.deepcopy(lfa).double()
    p2d = keep["p2"].detach().double().requires_grad_(True)
    y2 = F.leaky_relu(E.shared_mlp(l64.mlp2, p2d) + E.shared_mlp(l64.shortcut, xin.double()), 0.01)
    gref, = torch.autograd.grad(y2, p2d, gout.double())
    print(name, "p2.grad vs fp64 tail:", rel(gp2, gref), " y vs y2:", rel(y.detach(), y2.detach()))

print("---- replay check")
lfa = copy.deepcopy(lfa_a); lfa.train()
xi = x.clone().requires_grad_(True)
keep = {}
y = fused(lfa, xyz, xi, keep)
torch.cuda.synchronize()
gp2, = torch.autograd.grad(y, keep["p2"], gout, retain_graph=True)
gp2b, = torch.autograd.grad(y, keep["p2"], gout, retain_graph=True)
print("same twice:", rel(gp2, gp2b.double()))
p2r = keep["p2"].detach().clone().requires_grad_(True)
yr = F.leaky_relu(E.shared_mlp(lfa.mlp2, p2r) + E.shared_mlp(lfa.shortcut, x), 0.01)
gr, = torch.autograd.grad(yr, p2r, gout)
l64 = copy.deepcopy(lfa).double()
p2d = keep["p2"].detach().double().requires_grad_(True)
y2 = F.leaky_relu(E.shared_mlp(l64.mlp2, p2d) + E.shared_mlp(l64.shortcut, x.double()), 0.01)
gref, = torch.autograd.grad(y2, p2d, gout.double())
print("orig graph vs fp64:", rel(gp2, gref), " replayed fp32 tail vs fp64:", rel(gr, gref))
print("p2 strides", keep["p2"].stride(), keep["p2"].is_contiguous(), "pooled2", keep["pooled2"].stride())
# which element is off
diff = (gp2.double() - gref).abs()
i = diff.argmax()
print("worst at", np.unravel_index(int(i), tuple(diff.shape)) if False else int(i), float(diff.max()), float(gref.abs().max()))
bad_rows = (diff.amax(dim=2) > 1e-3 * gref.abs().max()).nonzero()
print("rows with error:", bad_rows.shape[0], "of", diff.shape[0] * diff.shape[1], bad_rows[:10].tolist())
bad_cols = (diff.amax(dim=(0, 1)) > 1e-3 * gref.abs().max()).nonzero()
print("cols with error:", bad_cols.shape[0], "of", diff.shape[2])
