"""Per-kernel device times of one eval forward (CUDA events from the C-ABI wrappers).
usage: python tools/layer_table_infer.py [n_points] [batch]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cabi = importlib.import_module("3d_recognizer_b200._cabi"); modules = importlib.import_module("3d_recognizer_b200.modules")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
st = modules.RandLANetSettings(n_classes=2, n_points=n, n_features=0, n_neighbors=16, knn="naive")
torch.manual_seed(0)
net = modules.RandLANet(st, torch.device("cuda")).eval()
x = torch.rand(B, n, 3, device="cuda")
with torch.no_grad():
    for _ in range(2): net(x)
    torch.cuda.synchronize()
    cabi.KERNEL_TIMERS = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net(x); e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
rows = []
for k, lst in cabi.KERNEL_TIMERS.items():
    ms = sum(a.elapsed_time(b) for a, b, _ in lst); fl = sum(w["flops"] for _, _, w in lst); by = sum(w["bytes"] for _, _, w in lst)
    rows.append((ms, k, len(lst), fl / ms * 1e-9, by / ms * 1e-6))
print(f"forward {tot:.2f} ms; instrumented {sum(r[0] for r in rows):.2f} ms")
for ms, k, c, tf, gb in sorted(rows, reverse=True)[:32]:
    print(f"{k:<44s} {ms:8.3f} ms  n={c}  {tf:7.2f} TFLOP/s  {gb:8.1f} GB/s")
