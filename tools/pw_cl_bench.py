"""Device-time comparison of the tensor-core per-point kernels (r3d_pc_gemm / r3d_pc_wgrad) with the FP32 kernels they
replace, on the layer shapes of the train40960 step.  Usage: python tools/pw_cl_bench.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("3d_recognizer_b200.ops")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


GEMM = [(655360, 32, 128), (655360, 64, 64), (655360, 64, 128), (655360, 128, 128), (655360, 128, 64), (655360, 32, 256),
        (163840, 128, 256), (163840, 128, 512), (163840, 128, 128), (163840, 64, 128), (40960, 128, 256),
        (2621440, 32, 64), (2621440, 64, 32), (2621440, 16, 32), (40960, 256, 512), (40960, 1024, 256), (655360, 256, 32),
        (163840, 512, 128), (40960, 512, 256), (40960, 256, 256), (163840, 256, 128), (40960, 256, 1024), (10240, 512, 512)]
WGRAD = [(40960, 256, 1024), (163840, 256, 128), (40960, 512, 256), (163840, 128, 512), (655360, 32, 256),
         (655360, 128, 64), (655360, 128, 32), (40960, 128, 256), (10240, 512, 512), (163840, 128, 128),
         (655360, 64, 64), (2621440, 32, 64), (2621440, 8, 64), (2621440, 16, 16)]

# one rank's shard of train40960 on 8 GPUs (8 clouds): the layers whose dispatch decides the strong-scaling tail
SHARD_GEMM = [(5120, 256, 512), (5120, 256, 1024), (5120, 1024, 256), (20480, 256, 128), (5120, 512, 256), (20480, 512, 128),
              (1280, 512, 512), (5120, 256, 256), (5120, 128, 256), (5120, 256, 128), (20480, 128, 128), (20480, 64, 128)]
SHARD_WGRAD = [(5120, 512, 256), (5120, 256, 1024), (5120, 128, 256), (1280, 512, 512), (5120, 256, 256), (20480, 256, 128),
               (327680, 64, 8), (327680, 32, 16)]

if __name__ == "__main__":
    if "--shard" in sys.argv:
        GEMM, WGRAD = SHARD_GEMM, SHARD_WGRAD
        L = importlib.import_module("3d_recognizer_b200._cabi").lib()
        print("fwd GEMM, forced round-1 3xTF32 kernel (r3d_pointwise_set_tensor_cores(2)) vs default dispatch")
        for M, cin, cout in GEMM:
            x = torch.randn(M, cin, device="cuda")
            wT = (torch.randn(cout, cin, device="cuda") * 0.1).t().contiguous()
            stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
            t0 = timeit(lambda: ops.pointwise(x.unsqueeze(0), wT, stats=stats))
            prev = L.r3d_pointwise_set_tensor_cores(2)
            t1 = timeit(lambda: ops.pointwise(x.unsqueeze(0), wT, stats=stats))
            L.r3d_pointwise_set_tensor_cores(prev)
            print(f"M={M:8d} {cin:4d}->{cout:4d}   default {t0:8.3f}   pw_tc {t1:8.3f}")
    print("fwd / dgrad GEMM              fp32 ms   tc ms   tc TFLOP/s   tc GB/s")
    for M, cin, cout in GEMM:
        x = torch.randn(M, cin, device="cuda")
        w = torch.randn(cout, cin, device="cuda") * 0.1
        wT = w.t().contiguous()
        stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
        t0 = timeit(lambda: ops.pointwise(x.unsqueeze(0), wT, stats=stats))
        t1 = timeit(lambda: ops.pc_gemm(x, w, stats=stats))
        print(f"M={M:8d} {cin:4d}->{cout:4d}   {t0:8.3f} {t1:8.3f} {2e-9 * M * cin * cout / t1:10.1f} {4e-6 * M * (cin + cout) / t1:10.0f}")
    print("wgrad                         fp32 ms   tc ms   tc TFLOP/s   tc GB/s   absmax ms")
    for M, ca, cb in WGRAD:
        a = torch.randn(M, ca, device="cuda") * 1e-3
        b = torch.randn(M, cb, device="cuda")
        sa, sb = a.abs().max().reshape(1), b.abs().max().reshape(1)
        t0 = timeit(lambda: ops.rowreduce_gemm(a, b))
        t1 = timeit(lambda: ops.pc_wgrad(a, b, sa, sb))
        t2 = timeit(lambda: ops.pc_wgrad(a, b)) - t1
        print(f"M={M:8d} {ca:4d}x{cb:4d}    {t0:8.3f} {t1:8.3f} {2e-9 * M * ca * cb / t1:10.1f} {4e-6 * M * (ca + cb) / t1:10.0f} {t2:8.3f}")
