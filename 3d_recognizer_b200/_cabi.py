"""ctypes binding of lib/libr3d_b200.so (include/r3d_b200.h).

This is the ONLY way the package reaches its kernels.  There is no CPU or PyTorch fallback: if the
library is missing or a kernel fails, the call raises.  Error codes are mapped onto the exception
types the reference raises at the same boundary (knn.cpp:15-17 -> RuntimeError; bad arguments ->
ValueError).
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libr3d_b200.so")

_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/r3d_b200.h declares
SIGNATURES = {
    "r3d_abi_version": (c_int, []),
    "r3d_error_string": (ctypes.c_char_p, [c_int]),
    "r3d_last_cuda_error": (ctypes.c_char_p, []),
    "r3d_launch_count": (ctypes.c_ulonglong, []),
    "r3d_knn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "r3d_knn": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_int,
                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "r3d_knn_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "r3d_knn_set_variant": (c_int, [c_int]),
    "r3d_knn_set_algorithm": (c_int, [c_int]),
    "r3d_knn_plan": (c_int, [c_int, c_int, c_int, c_int]),
    "r3d_knn_set_grid_density": (c_int, [ctypes.c_float]),
    "r3d_lfa_pool": (c_int, [c_int, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, ctypes.c_longlong,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_pool_tc": (c_int, [c_int, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, ctypes.c_longlong,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_tc_du2_floats": (ctypes.c_longlong, [c_int, c_int, c_int, c_int]),
    "r3d_absmax": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_void_p]),
    "r3d_lfa_tc_bwd": (c_int, [c_int, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, ctypes.c_longlong] +
                       [c_void_p] * 8 + [c_void_p, ctypes.c_longlong] + [c_void_p] * 10 +
                       [c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_r1_rows": (c_int, [c_void_p, ctypes.c_longlong] + [c_void_p] * 5 + [c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_du2_combine": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_tc_wide": (c_int, [c_int, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, ctypes.c_longlong] +
                        [c_void_p] * 8 + [ctypes.c_longlong] + [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_pool_bwd": (c_int, [c_int, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, ctypes.c_longlong] +
                         [c_void_p] * 10 + [c_void_p, ctypes.c_longlong] + [c_void_p] * 4 +
                         [c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_tile_points": (c_int, [c_int, c_int]),
    "r3d_lfa_tile_points_for": (c_int, [c_int, c_int, c_int, c_int]),
    "r3d_lfa_pool2_bwd_train": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_void_p, ctypes.c_longlong] +
                                [c_void_p] * 9 + [c_void_p, ctypes.c_longlong] + [c_void_p] * 3 +
                                [c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_bn2_bwd": (c_int, [c_void_p, ctypes.c_longlong] + [c_void_p] * 10 + [c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_moments": (c_int, [c_int, c_void_p, ctypes.c_longlong] + [c_void_p] * 10 +
                        [c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_rpe_rows": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "r3d_lfa_gather_concat": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "r3d_lfa_gather_concat_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int,
                                          c_int, c_int, c_void_p]),
    "r3d_lfa_attn_pool": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p]),
    "r3d_lfa_attn_pool_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int,
                                      c_void_p]),
    "r3d_add_lrelu": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_float, c_void_p]),
    "r3d_add_lrelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_float, c_void_p]),
    "r3d_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_void_p, ctypes.c_double,
                              ctypes.c_double, ctypes.c_double, ctypes.c_double, c_void_p, c_void_p]),
    "r3d_bn_from_moments": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, ctypes.c_double,
                                    c_void_p, c_void_p, c_void_p, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_bn_from_moments_bwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, ctypes.c_double] +
                                [c_void_p] * 10 + [c_void_p]),
    "r3d_pointwise_plan": (c_int, [c_int, c_int, c_int, ctypes.c_longlong, c_int]),
    "r3d_lfa_rpe1_grads": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, ctypes.c_double, c_void_p,
                                   c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_lfa_bn2_coeffs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_double, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "r3d_pointwise": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, ctypes.c_longlong, c_void_p,
                              ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_int, ctypes.c_float,
                              c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_pointwise_stats": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, ctypes.c_longlong, c_void_p,
                                    ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_int, ctypes.c_float,
                                    c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "r3d_pointwise_set_tensor_cores": (c_int, [c_int]),
    "r3d_bn_apply": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_float,
                             ctypes.c_float, c_void_p, c_void_p, c_void_p, c_int, ctypes.c_float, c_void_p, c_void_p,
                             c_void_p]),
    "r3d_bn_bwd_reduce": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int,
                                  ctypes.c_float, c_void_p, c_void_p]),
    "r3d_bn_bwd_dz": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, ctypes.c_float,
                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_tversky_loss_fwd": (c_int, [c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_int,
                                     c_int, c_int, c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "r3d_tversky_loss_bwd": (c_int, [c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_int,
                                     c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_confusion_counts": (c_int, [c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_int,
                                     c_int, c_int, c_void_p, c_void_p]),
    "r3d_upsample": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                             ctypes.c_float, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, ctypes.c_longlong,
                             c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_upsample_bwd": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                                 ctypes.c_float, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, ctypes.c_longlong,
                                 c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_feed_batch": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                               ctypes.c_float, ctypes.c_float, ctypes.c_ulonglong, ctypes.c_ulonglong, c_void_p, c_void_p,
                               c_int, c_void_p]),
    "r3d_sample_subset": (c_int, [c_void_p, c_int, ctypes.c_ulonglong, ctypes.c_ulonglong, c_void_p, c_int, c_void_p]),
    "r3d_pc_gemm_supported": (c_int, [c_int, c_int, ctypes.c_longlong]),
    "r3d_pc_gemm": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_void_p, c_int, ctypes.c_float, c_void_p,
                            ctypes.c_longlong, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p]),
    "r3d_pc_wgrad_supported": (c_int, [c_int, c_int, ctypes.c_longlong]),
    "r3d_pc_wgrad": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, ctypes.c_longlong, c_int, ctypes.c_longlong, c_void_p, c_void_p, c_void_p, c_int,
                             c_void_p]),
    "r3d_pointwise_bn": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p, c_void_p, c_int,
                                 ctypes.c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_bn_set_fused": (c_int, [c_int]),
    "r3d_bn_bwd": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, ctypes.c_float,
                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_bn_bwd_absmax": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, ctypes.c_float,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "r3d_rowreduce_gemm": (c_int, [c_void_p, c_int, c_void_p, c_int, ctypes.c_longlong, c_void_p, c_int, c_void_p]),
    "r3d_tc_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "r3d_tc16_probe": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.c_float,
                               ctypes.c_float, c_void_p]),
    "r3d_fp32_probe_floats": (c_size_t, []),
    "r3d_fp32_probe": (c_int, [c_int, c_int, c_void_p, ctypes.POINTER(ctypes.c_double), c_void_p]),
}


# ---- per-kernel device timing (bench.py's roofline): when KERNEL_TIMERS is a dict, every ops wrapper
# brackets its C-ABI call with CUDA events on the launching stream and files them under a kernel name.
KERNEL_TIMERS = None
TIMER_SHAPES = True    # per-point layer timers keyed by shape (tools/layer_table.py) instead of one aggregate


class kernel_timer:
    __slots__ = ("name", "work", "e0")

    def __init__(self, name, **work):
        self.name, self.work = name, work

    def __enter__(self):
        if KERNEL_TIMERS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if KERNEL_TIMERS is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            KERNEL_TIMERS.setdefault(self.name, []).append((self.e0, e1, self.work))
        return False


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library.  Raises (loudly) when it has not been built: run
    ``python -m 3d_recognizer_b200.build`` / ``__graft_entry__.build()``."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
                        "(python 3d_recognizer_b200/build.py). There is no CPU fallback.")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = res, args
                if handle.r3d_abi_version() != 1:
                    raise RuntimeError("libr3d_b200.so ABI version mismatch; rebuild")
                _lib = handle
    return _lib


_EXC = {-1: ValueError, -2: RuntimeError, -3: ValueError, -4: ValueError, -5: ValueError, -6: RuntimeError,
        -7: ValueError}


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    L = lib()
    msg = L.r3d_error_string(rc).decode()
    if rc == -6:
        msg += ": " + L.r3d_last_cuda_error().decode()
    raise _EXC.get(rc, RuntimeError)(f"{what}: {msg}")


def ptr(t):
    """Device (or host) address of a contiguous tensor, or NULL for None."""
    if t is None:
        return None
    assert t.is_contiguous(), "C-ABI tensors must be contiguous"
    return ctypes.c_void_p(t.data_ptr())


def raw(t):
    """Address of a tensor whose strides are passed separately (row-contiguous views)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 hot path has no CPU fallback")
