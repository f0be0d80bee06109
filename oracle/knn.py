"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

ctypes front-ends for
  * the C restatement of the exact KNN (oracle/knn_oracle.c -> oracle/_build/libknn_oracle.so)
  * the REFERENCE's own nanoflann KNN compiled from /root/reference
    (oracle/ref_knn_shim.cpp -> oracle/_ref/libref_knn.so), when it has been built.

Both follow the contract of the reference binding ``knn_tpk.knn(support, querry, k)``
(randlanet/utils/src/bindings.cpp:5-7, knn.cpp:43-61): (B,Ns,3)+(B,Nq,3) float32 in,
(B,Nq,K) int64 indices + (B,Nq,K) float32 SQUARED distances out, ascending.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "libknn_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_knn.so")
_REF_SRC = "/root/reference/randlanet/utils/src"

_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int64)


def build(ref: bool = True) -> None:
    """Compile the oracle (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref and os.path.isdir(_REF_SRC):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def have_ref() -> bool:
    return os.path.isfile(_REF_SO)


_libs = {}


def _lib(path):
    if path not in _libs:
        if not os.path.isfile(path):
            if path == _ORACLE_SO:
                build(ref=False)
            else:
                raise FileNotFoundError(f"{path} not built (run `make -C oracle ref` in the build container)")
        _libs[path] = ctypes.CDLL(path)
    return _libs[path]


def _prep(support, query):
    support = np.ascontiguousarray(support, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    squeeze = support.ndim == 2
    if squeeze:
        support, query = support[None], query[None]
    assert support.ndim == 3 and query.ndim == 3 and support.shape[2] == 3 and query.shape[2] == 3
    assert support.shape[0] == query.shape[0]
    return support, query, squeeze


def knn_exact(support, query, k: int, threads: int = 0):
    """Canonical exact KNN: d2 = ((dx*dx+dy*dy)+dz*dz) in fp32, order (d2, index) ascending.

    Returns (idx int64 (B,Nq,K), d2 float32 (B,Nq,K)).  ``Ns < k`` raises RuntimeError, as
    knn.cpp:15-17 does."""
    support, query, squeeze = _prep(support, query)
    B, Ns, _ = support.shape
    Nq = query.shape[1]
    idx = np.empty((B, Nq, k), dtype=np.int64)
    d2 = np.empty((B, Nq, k), dtype=np.float32)
    fn = _lib(_ORACLE_SO).oracle_knn
    fn.restype = ctypes.c_int
    fn.argtypes = [_fp, _fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _ip, _fp, ctypes.c_int]
    rc = fn(support.ctypes.data_as(_fp), query.ctypes.data_as(_fp), B, Ns, Nq, k,
            idx.ctypes.data_as(_ip), d2.ctypes.data_as(_fp), threads)
    if rc == -2:
        raise RuntimeError(f"Not enough points in support to find {k} neighboors")
    if rc != 0:
        raise ValueError(f"oracle_knn: bad argument (rc={rc})")
    return (idx[0], d2[0]) if squeeze else (idx, d2)


def ref_knn_tpk(support, query, k: int):
    """The reference's nanoflann KD-tree KNN (single threaded), via oracle/_ref."""
    support, query, squeeze = _prep(support, query)
    B, Ns, _ = support.shape
    Nq = query.shape[1]
    idx = np.empty((B, Nq, k), dtype=np.int64)
    d2 = np.empty((B, Nq, k), dtype=np.float32)
    fn = _lib(_REF_SO).ref_knn_tpk
    fn.restype = ctypes.c_int
    fn.argtypes = [_fp, _fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _ip, _fp]
    rc = fn(support.ctypes.data_as(_fp), query.ctypes.data_as(_fp), B, Ns, Nq, k,
            idx.ctypes.data_as(_ip), d2.ctypes.data_as(_fp))
    if rc == -2:
        raise RuntimeError(f"Not enough points in support to find {k} neighboors")
    return (idx[0], d2[0]) if squeeze else (idx, d2)
