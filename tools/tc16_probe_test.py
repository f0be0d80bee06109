"""GPU check of the split-fp16 tcgen05 building blocks (r3d_tc16_probe): all four major-ness combinations, M = 128 and
M = 64 (prints the TMEM lane mapping of M = 64), error against an fp64 product."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cabi = importlib.import_module("3d_recognizer_b200._cabi")
L = cabi.lib()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for M in (128, 64):
    for (N, K) in ((64, 128), (128, 64), (256, 16), (64, 64)):
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                A = torch.randn(M, K, device=dev)
                B = torch.randn(N, K, device=dev)
                out = torch.zeros(128, N, device=dev)
                cabi.check(L.r3d_tc16_probe(cabi.ptr(A), cabi.ptr(B), cabi.ptr(out), M, N, K, a_mn, b_mn, 16.0, 64.0,
                                            cabi.stream_ptr(dev)), "probe")
                torch.cuda.synchronize()
                ref = (A.double() @ B.double().t())
                written = ~torch.isnan(out).all(dim=1)
                lanes = written.nonzero().flatten().tolist()
                if M == 128:
                    got = out
                else:
                    got = out[written]
                ok = got.shape == ref.shape
                err = float((got.double() - ref).abs().max() / ref.abs().max()) if ok else float("nan")
                print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: lanes written {len(lanes)} "
                      f"[{lanes[0] if lanes else None}..{lanes[-1] if lanes else None}] rel err {err:.2e}", flush=True)
                if M == 64 and a_mn == 0 and b_mn == 0 and N == 64 and K == 64:
                    print("  M=64 lanes:", lanes)
