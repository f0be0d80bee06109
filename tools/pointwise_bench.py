"""Per-point layer micro-benchmark: tcgen05 3xTF32 kernel vs the FP32 CUDA-core kernels over row counts (20 launches
captured in a CUDA graph per measurement).  usage: python tools/pointwise_bench.py"""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from knn_small_bench import graph_time  # noqa: E402

ops = importlib.import_module("3d_recognizer_b200.ops")
L = importlib.import_module("3d_recognizer_b200._cabi").lib()

if __name__ == "__main__":
    g = torch.Generator(device="cuda").manual_seed(0)
    for M in (1248, 5000, 20000, 40960, 163840, 655360):
        for cin, cout in ((32, 32), (64, 128), (256, 32), (128, 256), (512, 256)):
            x = torch.randn(1, M, cin, device="cuda", generator=g)
            w = torch.randn(cout, cin, device="cuda", generator=g)
            stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
            row = dict(M=M, cin=cin, cout=cout)
            for name, tc in (("fp32_us", 0), ("tc_us", 2)):
                L.r3d_pointwise_set_tensor_cores(tc)
                row[name] = round(1e3 * graph_time(lambda: ops.pointwise(x, w, stats=stats, w_out_in=True), reps=10), 2)
            L.r3d_pointwise_set_tensor_cores(1)
            print(json.dumps(row), flush=True)
