"""KNN micro-benchmark (BASELINE config 5 and the config-D level-0 shape): CUDA-event timing of
r3d_knn for each variant.  usage: python tools/knn_bench.py [--quick]"""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("3d_recognizer_b200.ops")
L = importlib.import_module("3d_recognizer_b200._cabi").lib()


def time_knn(B, Ns, Nq, K, variant, iters=3, algo=1):
    g = torch.Generator(device="cuda").manual_seed(0)
    s = torch.rand(B, Ns, 3, device="cuda", generator=g)
    q = s if Ns == Nq else torch.rand(B, Nq, 3, device="cuda", generator=g)
    L.r3d_knn_set_variant(variant)
    L.r3d_knn_set_algorithm(algo)
    ops.knn(s, q, K, idx64=False, idx32=True, dist=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops.knn(s, q, K, idx64=False, idx32=True, dist=True)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    pairs = B * Ns * Nq
    return dict(B=B, Ns=Ns, Nq=Nq, K=K, algo=("auto", "brute", "grid")[algo], variant=variant, ms=ms, queries_per_s=B * Nq / ms * 1e3,
                pairs_per_s=pairs / ms * 1e3, tflops_8=8 * pairs / ms * 1e-9)


if __name__ == "__main__":
    if "--shape" in sys.argv:      # --shape B,Ns,Nq,K [--variant V] [--iters I]: one shape (ncu captures)
        shp = tuple(int(v) for v in sys.argv[sys.argv.index("--shape") + 1].split(","))
        var = int(sys.argv[sys.argv.index("--variant") + 1]) if "--variant" in sys.argv else 2
        its = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 1
        algo = int(sys.argv[sys.argv.index("--algo") + 1]) if "--algo" in sys.argv else 1
        print(json.dumps(time_knn(*shp, variant=var, iters=its, algo=algo)), flush=True)
        sys.exit(0)
    quick = "--quick" in sys.argv
    shapes = [(64, 40960, 40960, 16), (1, 1 << 20, 1 << 20, 16), (1, 1 << 20, 1 << 20, 32), (8, 2500, 2500, 16),
              (8, 625, 2500, 1), (64, 10240, 40960, 1)]
    if quick:
        shapes = [(8, 40960, 40960, 16), (8, 2500, 2500, 16)]
    if "--sweep" in sys.argv:       # brute force vs grid crossover
        shapes = [(8, 2500, 2500, 16), (8, 4096, 4096, 16), (8, 8192, 8192, 16), (8, 16384, 16384, 16),
                  (64, 40960, 40960, 16), (8, 65536, 65536, 16), (4, 262144, 262144, 16), (1, 1 << 20, 1 << 20, 16),
                  (1, 1 << 20, 1 << 20, 32), (8, 2500, 2500, 32), (8, 625, 2500, 1), (64, 10240, 40960, 1),
                  (32, 4096, 16384, 1), (32, 65536, 262144, 1)]
        for shp in shapes:
            for algo in (1, 2):
                if algo == 1 and shp[1] * shp[2] * shp[0] > 3e12:
                    continue
                print(json.dumps(time_knn(*shp, variant=2, algo=algo)), flush=True)
        sys.exit(0)
    for shp in shapes:
        for v in (0, 1, 2):
            print(json.dumps(time_knn(*shp, variant=v)), flush=True)
