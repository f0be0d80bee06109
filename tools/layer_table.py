"""Per-layer device times of one eager training step (CUDA events from the C-ABI wrappers).
usage: python tools/layer_table.py [n_points] [batch]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cabi = importlib.import_module("3d_recognizer_b200._cabi"); modules = importlib.import_module("3d_recognizer_b200.modules")
model_mod = importlib.import_module("3d_recognizer_b200.model"); syn = importlib.import_module("3d_recognizer_b200.synthetic")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
st = modules.RandLANetSettings(n_classes=2, n_points=n, n_features=0, n_neighbors=16, knn="naive")
torch.manual_seed(0)
m = model_mod.Model(st)
x, y = syn.fingertip_batch(0, min(B, 4), n, n_raw=max(150000, 2 * n))
x = torch.from_numpy(np.tile(x, (B // x.shape[0], 1, 1))).cuda(); y = torch.from_numpy(np.tile(y, (B // y.shape[0], 1))).cuda()
opt = m.make_optimizer()
for _ in range(2): m.train_step(x, y, opt)
torch.cuda.synchronize()
cabi.TIMER_SHAPES = True; cabi.KERNEL_TIMERS = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.train_step(x, y, opt); e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
rows = []
for k, lst in cabi.KERNEL_TIMERS.items():
    ms = sum(a.elapsed_time(b) for a, b, _ in lst); fl = sum(w["flops"] for _, _, w in lst); by = sum(w["bytes"] for _, _, w in lst)
    rows.append((ms, k, len(lst), fl / ms * 1e-9, by / ms * 1e-6))
print(f"step {tot:.2f} ms; instrumented {sum(r[0] for r in rows):.2f} ms")
for ms, k, c, tf, gb in sorted(rows, reverse=True)[:40]:
    print(f"{k:<44s} {ms:8.3f} ms  n={c}  {tf:7.2f} TFLOP/s  {gb:8.1f} GB/s")
