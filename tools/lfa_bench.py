"""Fused LocSE + attentive pooling micro-benchmark: forward / backward / moments per encoder level at the
config-D shapes (N=40960 -> 40960/10240/2560/640, d=16/64/128/256).  CUDA-event timing.
usage: python tools/lfa_bench.py [--batch 64] [--levels 0,1,2,3] [--iters 3] [--k 16]"""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = importlib.import_module("3d_recognizer_b200.ops")


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def timeit(fn, iters):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def main():
    B, K, iters = arg("--batch", 64), arg("--k", 16), arg("--iters", 3)
    levels = [int(v) for v in arg("--levels", "0,1,2,3").split(",")]
    shapes = [(40960, 16), (10240, 64), (2560, 128), (640, 256)]
    g = torch.Generator(device="cuda").manual_seed(0)
    for l in levels:
        N, d = shapes[l]
        h = d // 2
        xyz = torch.rand(B, N, 3, device="cuda", generator=g)
        feat = torch.randn(B, N, h, device="cuda", generator=g)
        idx = ops.knn(xyz, xyz, K, idx64=False, idx32=True, dist=False)["idx32"]
        w1 = torch.randn(h, 10, device="cuda", generator=g)
        a1 = torch.rand(h, device="cuda", generator=g) + 0.5
        b1 = torch.randn(h, device="cuda", generator=g) * 0.3
        w2 = (torch.randn(h, h, device="cuda", generator=g) / h ** 0.5).contiguous()
        ws = (torch.randn(d, d, device="cuda", generator=g) / d ** 0.5).contiguous()
        w2T, wsT = w2.t().contiguous(), ws.t().contiguous()
        dp = torch.randn(B, N, d, device="cuda", generator=g)
        rows = B * N * K
        for stage in (1, 2):
            s2 = stage == 2
            f_flops = B * N * (2 * K * (10 * h + d * d + d + (h * h if s2 else 0)))
            b_flops = B * N * (2 * K * (10 * h + 3 * d * d + d + (3 * h * h if s2 else 0)))
            ms_f = timeit(lambda: ops.lfa_pool(stage, xyz, idx, feat, w1, a1, b1, w2T if s2 else None,
                                               a1 if s2 else None, b1 if s2 else None, wsT), iters)
            ms_b = timeit(lambda: ops.lfa_pool_bwd(stage, xyz, idx, feat, w1, a1, b1, w2T if s2 else None,
                                                   a1 if s2 else None, b1 if s2 else None, w2 if s2 else None, wsT,
                                                   ws, dp), iters)
            tcd = {}
            if ops.lfa_pool_tc_supported(d, K):
                ms_t = timeit(lambda: ops.lfa_pool_tc(stage, xyz, idx, feat, w1, a1, b1, w2 if s2 else None,
                                                      a1 if s2 else None, b1 if s2 else None, ws), iters)
                ref = ops.lfa_pool(stage, xyz, idx, feat, w1, a1, b1, w2T if s2 else None, a1 if s2 else None,
                                   b1 if s2 else None, wsT)
                got = ops.lfa_pool_tc(stage, xyz, idx, feat, w1, a1, b1, w2 if s2 else None, a1 if s2 else None,
                                      b1 if s2 else None, ws)
                tcd = dict(tc_fwd_ms=ms_t, tc_fwd_tflops=f_flops / ms_t * 1e-9,
                           tc_vs_cuda_core_rel=float((got - ref).abs().max() / ref.abs().max()))
            print(json.dumps(dict(level=l, N=N, d=d, K=K, B=B, stage=stage, **tcd, fwd_ms=ms_f, fwd_tflops=f_flops / ms_f * 1e-9,
                                  bwd_ms=ms_b, bwd_tflops=b_flops / ms_b * 1e-9)), flush=True)
        ms0 = timeit(lambda: ops.lfa_moments(0, xyz, idx, d), iters)
        ms1 = timeit(lambda: ops.lfa_moments(1, xyz, idx, d, w1, a1, b1), iters)
        gs = torch.randn(h, h, device="cuda", generator=g)
        gsym = (gs + gs.t()).contiguous()
        ms2 = timeit(lambda: ops.lfa_moments(2, xyz, idx, d, w1, a1, b1, gsym=gsym, gsum=a1), iters)
        print(json.dumps(dict(level=l, N=N, d=d, rows=rows, moments_rpe_ms=ms0, moments_r1_ms=ms1, moments_r1_bwd_ms=ms2)),
              flush=True)


if __name__ == "__main__":
    main()
