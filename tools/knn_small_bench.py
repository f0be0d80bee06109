"""Small-cloud KNN timing (the down-sampled levels and the decoder of a 2 500-point cloud): 20 launches captured
in one CUDA graph so that python launch overhead stays out of the number.  usage: python tools/knn_small_bench.py"""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("3d_recognizer_b200.ops")
L = importlib.import_module("3d_recognizer_b200._cabi").lib()


def graph_time(fn, reps=20, iters=10):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * iters)


if __name__ == "__main__":
    gen = torch.Generator(device="cuda").manual_seed(0)
    for B, Ns, Nq, K in [(8, 625, 625, 16), (8, 156, 156, 16), (8, 39, 39, 16), (8, 625, 2500, 1), (8, 156, 625, 1),
                         (8, 39, 156, 1), (8, 2000, 2000, 16), (8, 625, 625, 32), (64, 640, 640, 16), (64, 160, 160, 16), (8, 2500, 2500, 16),
                         (8, 2500, 2500, 32), (64, 40960, 40960, 16), (64, 10240, 40960, 1), (1, 1 << 20, 1 << 20, 16),
                         (1, 1 << 20, 1 << 20, 32), (1, 1 << 18, 1 << 18, 64), (4, 40960, 40960, 16), (2, 40960, 40960, 16)]:
        s = torch.rand(B, Ns, 3, device="cuda", generator=gen)
        q = s if Ns == Nq else torch.rand(B, Nq, 3, device="cuda", generator=gen)
        row = dict(B=B, Ns=Ns, Nq=Nq, K=K)
        reps = 20 if B * Nq < 100000 else 2
        for name, algo in (("auto", 0), ("tiled", 1), ("grid", 2), ("grid_thread", 4)):
            if algo == 1 and B * Ns * Nq > 2e11:      # the exhaustive scan of a 1M x 1M search takes ~0.3 s
                continue
            L.r3d_knn_set_algorithm(algo)
            row[name + "_us"] = round(1e3 * graph_time(lambda: ops.knn(s, q, K, idx64=False, idx32=True, dist=True), reps=reps), 2)
        L.r3d_knn_set_algorithm(0)
        print(json.dumps(row), flush=True)
