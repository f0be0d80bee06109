"""Debug helper: _LfaPoolFn forward/backward vs fp64 torch composition, per output, repeated."""
import importlib, sys, os
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
engine = importlib.import_module("3d_recognizer_b200.engine")
ops = importlib.import_module("3d_recognizer_b200.ops")

def run(d, K, N, stage, B=2, seed=0):
    h = d // 2
    g = torch.Generator(device="cuda").manual_seed(seed)
    xyz = torch.rand(B, N, 3, device="cuda", generator=g)
    feat = torch.randn(B, N, h, device="cuda", generator=g)
    nn_ = ops.knn(xyz, xyz, K, idx64=True, idx32=True, dist=True)
    P = dict(w1=torch.randn(h, 10, device="cuda", generator=g), a1=torch.rand(h, device="cuda", generator=g) + 0.5,
             c1=torch.randn(h, device="cuda", generator=g) * 0.3, w2=torch.randn(h, h, device="cuda", generator=g) / h ** 0.5,
             a2=torch.rand(h, device="cuda", generator=g) + 0.5, c2=torch.randn(h, device="cuda", generator=g) * 0.3,
             ws=torch.randn(d, d, device="cuda", generator=g) / d ** 0.5)
    gout = torch.randn(B, N, d, device="cuda", generator=g)
    def leaves(dt):
        L = {k: v.to(dt).clone().requires_grad_(True) for k, v in P.items()}
        L["feat"] = feat.to(dt).clone().requires_grad_(True)
        return L
    A = leaves(torch.float32)
    out = engine._LfaPoolFn.apply(stage, xyz, nn_["idx32"], A["feat"], A["w1"], A["a1"], A["c1"],
                                  A["w2"] if stage == 2 else None, A["a2"] if stage == 2 else None,
                                  A["c2"] if stage == 2 else None, A["ws"])
    (out * gout).sum().backward()
    R = leaves(torch.float64)
    rpe = engine.relative_position_encoding(xyz.double(), nn_["idx64"], nn_["dist"].double())
    r = F.relu(rpe @ R["w1"].t() * R["a1"] + R["c1"])
    if stage == 2:
        r = F.relu(r @ R["w2"].t() * R["a2"] + R["c2"])
    x = torch.cat((r, engine.gather_points(R["feat"], nn_["idx64"])), dim=-1)
    ref = (F.softmax(x @ R["ws"].t(), dim=2) * x).sum(dim=2)
    (ref * gout.double()).sum().backward()
    res = {"out": float((out.double() - ref).abs().max() / ref.abs().max())}
    for k in R:
        if R[k].grad is None:
            continue
        res[k] = float((A[k].grad.double() - R[k].grad).abs().max() / R[k].grad.abs().max())
    return res

if __name__ == "__main__":
    for (d, K, N) in [(128, 16, 300), (128, 16, 296), (256, 16, 150), (64, 16, 625), (16, 16, 1000), (128, 32, 100)]:
        for stage in (1, 2):
            for rep in range(2):
                r = run(d, K, N, stage, seed=rep)
                bad = {k: f"{v:.1e}" for k, v in r.items() if v > 1e-4}
                print(d, K, N, "stage", stage, "rep", rep, "max", f"{max(r.values()):.1e}", "BAD" if bad else "ok", bad, flush=True)
