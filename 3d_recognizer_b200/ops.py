"""Tensor-level wrappers over the C ABI (include/r3d_b200.h).  Each wrapper allocates outputs and the
workspace with torch (plumbing), passes raw device pointers + the current CUDA stream, and raises on
any non-zero status.  No wrapper has a PyTorch or CPU fallback."""
from typing import Optional, Tuple

import ctypes
import numpy as np
import torch

from . import _cabi


def _cloud_view(x: torch.Tensor):
    """(tensor, batch stride in elements) for a (B,N,C) fp32 tensor whose clouds are dense row-major;
    prefix views x[:, :n] of a contiguous tensor pass through without a copy."""
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    if x.shape[0] == 0 or x.shape[1] == 0:
        return x.contiguous(), 0
    if x.stride(2) == 1 and x.stride(1) == x.shape[2] and (x.shape[0] == 1 or x.stride(0) >= x.shape[1] * x.shape[2]):
        return x, (x.stride(0) if x.shape[0] > 1 else x.shape[1] * x.shape[2])
    x = x.contiguous()
    return x, x.shape[1] * x.shape[2]


def knn(support: torch.Tensor, query: torch.Tensor, k: int, *, idx64: bool = True, idx32: bool = False,
        dist: bool = True, dist_sq: bool = False):
    """Exact K nearest neighbours of every ``query`` point among ``support`` (C ABI ``r3d_knn``).

    Replaces ``knn_tpk.knn`` (randlanet/utils/src/knn.cpp:43-61) and the back-ends of
    ``KNN.forward`` (randlanet/utils/modules.py:118-150).  support (B,Ns,3), query (B,Nq,3) fp32 CUDA.
    Returns a dict with the requested outputs, each (B,Nq,K): ``idx64`` int64, ``idx32`` int32,
    ``dist`` fp32 = sqrt(d2) (what KNN.forward returns), ``dist_sq`` fp32 (what knn_tpk.knn returns).
    Order: (d2, index) ascending — exact ties go to the lower index."""
    _cabi.require_cuda(support, "support")
    _cabi.require_cuda(query, "query")
    if support.dim() != 3 or query.dim() != 3 or support.shape[-1] != 3 or query.shape[-1] != 3:
        raise ValueError("support and query must have shape (B, N, 3)")
    if support.shape[0] != query.shape[0]:
        raise ValueError("support and query must have the same batch size")
    same = query is support
    support, s_stride = _cloud_view(support)
    query, q_stride = (support, s_stride) if same else _cloud_view(query)
    B, Ns, _ = support.shape
    Nq = query.shape[1]
    dev = support.device
    L = _cabi.lib()
    out = {}
    with torch.cuda.device(dev):
        if idx64:
            out["idx64"] = torch.empty((B, Nq, k), dtype=torch.int64, device=dev)
        if idx32:
            out["idx32"] = torch.empty((B, Nq, k), dtype=torch.int32, device=dev)
        if dist:
            out["dist"] = torch.empty((B, Nq, k), dtype=torch.float32, device=dev)
        if dist_sq:
            out["dist_sq"] = torch.empty((B, Nq, k), dtype=torch.float32, device=dev)
        wbytes = L.r3d_knn_workspace_bytes(B, Ns, Nq, k)
        ws = torch.empty((wbytes,), dtype=torch.uint8, device=dev)
        with _cabi.kernel_timer("knn_k1" if k == 1 else "knn", flops=8.0 * B * Ns * Nq,
                                bytes=4.0 * B * (3 * Ns + 3 * Nq + Nq * k * (len(out) + ("idx64" in out)))):
            rc = L.r3d_knn(ctypes.c_void_p(support.data_ptr()), s_stride, ctypes.c_void_p(query.data_ptr()),
                           q_stride, B, Ns, Nq, k,
                           _cabi.ptr(out.get("idx64")), _cabi.ptr(out.get("idx32")), _cabi.ptr(out.get("dist")),
                           _cabi.ptr(out.get("dist_sq")), _cabi.ptr(ws), wbytes, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_knn")
    return out


def knn_host(support: np.ndarray, query: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Host-buffer drop-in for ``knn_tpk.knn(support, querry, k)`` (bindings.cpp:5-7): numpy / CPU
    buffers in, (idx int64, d2 fp32) out; H2D, kernel and D2H happen inside ``r3d_knn_host``."""
    support = np.ascontiguousarray(support, dtype=np.float32)
    query = support if query is support else np.ascontiguousarray(query, dtype=np.float32)
    if support.ndim != 3 or query.ndim != 3 or support.shape[-1] != 3 or query.shape[-1] != 3:
        raise ValueError("support and query must have shape (B, N, 3)")
    B, Ns, _ = support.shape
    Nq = query.shape[1]
    idx = np.empty((B, Nq, k), dtype=np.int64)
    d2 = np.empty((B, Nq, k), dtype=np.float32)
    rc = _cabi.lib().r3d_knn_host(ctypes.c_void_p(support.ctypes.data), ctypes.c_void_p(query.ctypes.data), B, Ns, Nq,
                                  k, ctypes.c_void_p(idx.ctypes.data), ctypes.c_void_p(d2.ctypes.data))
    _cabi.check(rc, "r3d_knn_host")
    return idx, d2
