"""tcgen05 3xTF32 GEMM (r3d_tc_gemm): accuracy vs fp64 and throughput.  usage: python tools/tc_gemm_test.py"""
import importlib, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cabi = importlib.import_module("3d_recognizer_b200._cabi")
L = cabi.lib()

def run(M, N, K, terms):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    C = torch.empty(M, N, device="cuda")
    rc = L.r3d_tc_gemm(cabi.ptr(A), cabi.ptr(W), cabi.ptr(C), M, N, K, terms, cabi.stream_ptr(A.device))
    cabi.check(rc, "r3d_tc_gemm")
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    ref32 = float(((A @ W.t()).double() - ref).abs().max() / ref.abs().max())
    return err, ref32

def bench(M, N, K, terms, iters=5):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    best = None
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.r3d_tc_gemm(cabi.ptr(A), cabi.ptr(W), cabi.ptr(C), M, N, K, terms, cabi.stream_ptr(A.device)); e1.record()
        torch.cuda.synchronize(); ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
    return best, 2.0 * M * N * K / best * 1e-9

if __name__ == "__main__":
    for (M, N, K) in [(128, 32, 8), (128, 64, 32), (1000, 64, 64), (777, 128, 128), (4096, 256, 256), (300, 256, 1024), (5000, 32, 16)]:
        for terms in (1, 3):
            err, e32 = run(M, N, K, terms)
            print(json.dumps(dict(M=M, N=N, K=K, terms=terms, rel_err=err, fp32_matmul_rel_err=e32)), flush=True)
    for (M, N, K) in [(1 << 20, 256, 256), (1 << 20, 128, 128), (1 << 20, 64, 64), (1 << 22, 32, 32)]:
        for terms in (1, 3):
            ms, tf = bench(M, N, K, terms)
            print(json.dumps(dict(M=M, N=N, K=K, terms=terms, ms=ms, tflops=tf)), flush=True)
