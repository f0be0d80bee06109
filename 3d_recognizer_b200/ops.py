"""Tensor-level wrappers over the C ABI (include/r3d_b200.h).  Each wrapper allocates outputs and the
workspace with torch (plumbing), passes raw device pointers + the current CUDA stream, and raises on
any non-zero status.  No wrapper has a PyTorch or CPU fallback."""
from typing import Optional, Tuple

import ctypes
import os
import numpy as np
import torch

from . import _cabi



class _ZeroPool:
    """Zero-filled scratch (accumulators of the atomics-based reductions) carved from one block per step.

    A 2 500-point training step needs ~160 small zero-filled buffers (BatchNorm statistics, weight-gradient
    accumulators, moment matrices); one fill kernel each was 5 % of the step (profiles/r01_launches_train2500.txt).
    Slices are handed out once and never reused, so a buffer stays valid for as long as anything references it; a
    block is sized from what the previous step took (``begin`` marks the step boundary)."""
    MAX_BYTES = 4 << 20          # larger requests get their own fill: nothing to win by batching them

    def __init__(self):
        self.block, self.off, self.need, self.hint, self.captured = None, 0, 0, 1 << 20, False
        self.block_stream, self.block_ready, self.waited = None, None, set()

    def begin(self):
        self.hint = max(1 << 20, self.need)
        self.block, self.off, self.need = None, 0, 0
        self.block_stream, self.block_ready, self.waited = None, None, set()

    def take(self, shape, dtype, device):
        numel = 1
        for v in (shape if isinstance(shape, (tuple, list)) else (shape,)):
            numel *= int(v)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        if nbytes > self.MAX_BYTES or nbytes == 0:
            return torch.zeros(shape, dtype=dtype, device=device)
        aligned = (nbytes + 255) & ~255
        self.need += aligned
        # a block filled outside a CUDA-graph capture must not feed a captured step: its fill would not be replayed
        on_cuda = torch.device(device).type == "cuda"
        capturing = on_cuda and torch.cuda.is_current_stream_capturing()
        if (self.block is None or self.block.device != torch.device(device) or capturing != self.captured
                or self.off + aligned > self.block.numel()):
            if on_cuda and torch.cuda.current_stream(device).cuda_stream in SIDE_STREAMS:
                # Inside a fork (engine._fork registers its streams in SIDE_STREAMS): a block filled here would be used
                # by the parent stream before the join, unordered against the fill (in a captured graph: no edge).
                # This request gets its own buffer; the next step's block, sized by ``hint`` and allocated on the
                # main stream before any fork, covers it.
                return torch.zeros(shape, dtype=dtype, device=device)
            self.block = torch.zeros(max(self.hint, aligned), dtype=torch.uint8, device=device)
            self.off, self.captured = 0, capturing
            if on_cuda:
                # the fill is ordered on the creating stream only: remember it, and an event behind the fill
                self.block_stream = torch.cuda.current_stream(device).cuda_stream
                self.block_ready = torch.cuda.Event()
                self.block_ready.record()
                self.waited = set()
        if on_cuda and self.block_stream is not None:
            # A slice taken on ANOTHER stream (a fork, or an autograd node replayed on the stream its forward ran on)
            # is only ordered after the fill if that stream has synchronised with the creating stream since; make it
            # wait for the fill once per block (cheap; a graph edge under capture).
            cur = torch.cuda.current_stream(device)
            if cur.cuda_stream != self.block_stream and cur.cuda_stream not in self.waited:
                cur.wait_event(self.block_ready)
                self.waited.add(cur.cuda_stream)
        out = self.block[self.off:self.off + nbytes].view(dtype).view(shape)
        self.off += aligned
        return out


ZEROS = _ZeroPool()
SIDE_STREAMS = set()         # raw handles of the side streams engine._fork runs launches on


def zeros(shape, dtype, device):
    """Zero-filled scratch tensor from the per-step pool (see ``_ZeroPool``)."""
    return ZEROS.take(shape, dtype, device)

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
        # the uniform-grid back-end does O(N K) work: crediting it with the 8 Nq Ns flop of the exhaustive scan would
        # put it far above any roofline, so only the brute-force kernel carries algorithmic flops
        plan = L.r3d_knn_plan(B, Ns, Nq, k)
        grid = plan == 2
        with _cabi.kernel_timer(f"knn_{ {1: 'brute', 2: 'grid', 3: 'small'}[plan]}_k{k}[Ns={Ns}]",
                                flops=0.0 if grid else 8.0 * B * Ns * Nq,
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
    # page-locked outputs (torch's caching host allocator: a buffer released by the caller is reused by the next call):
    # the device -> host copies of r3d_knn_host are then direct DMAs that overlap the search of the next chunk
    if torch.cuda.is_available():
        idx = torch.empty((B, Nq, k), dtype=torch.int64, pin_memory=True).numpy()
        d2 = torch.empty((B, Nq, k), dtype=torch.float32, pin_memory=True).numpy()
    else:
        idx = np.empty((B, Nq, k), dtype=np.int64)
        d2 = np.empty((B, Nq, k), dtype=np.float32)
    rc = _cabi.lib().r3d_knn_host(ctypes.c_void_p(support.ctypes.data), ctypes.c_void_p(query.ctypes.data), B, Ns, Nq,
                                  k, ctypes.c_void_p(idx.ctypes.data), ctypes.c_void_p(d2.ctypes.data))
    _cabi.check(rc, "r3d_knn_host")
    return idx, d2


def _rows_view(x: torch.Tensor):
    """(tensor, batch stride in floats) for a (B,n,C) fp32 tensor with contiguous rows; prefix views
    x[:, :n] of a dense tensor pass through without a copy."""
    if x.dtype != torch.float32:
        x = x.float()
    B, n, C = x.shape
    if B == 0 or n == 0:
        return x.contiguous(), 0
    ok = x.stride(2) == 1 and x.stride(1) == C and (B == 1 or x.stride(0) >= n * C)
    if not ok:
        x = x.contiguous()
    return x, (x.stride(0) if B > 1 else n * C)


# tcgen05 path of the fused LocSE + pooling kernels (csrc/lfa_cl*.cu, "channel-lane" kernels: transposed contraction,
# split-fp16 operands, warp-specialised persistent CTAs).  d in TC_WIDTHS, K in TC_NEIGHBORS; d = 256 stays on the
# FP32 CUDA-core kernels (its 256 x 256 weight image does not fit shared memory next to the row operands).
USE_TENSOR_CORES = True
TC_WIDTHS, TC_NEIGHBORS = (64, 128, 256), (16, 32)     # forward: widths where the tensor-core kernel is the faster one
TC_BWD_WIDTHS = (16, 64, 128, 256)                     # backward (d = 256: stage-1 backward and pass 1 of stage 2)
# r1 moments on the tensor cores, taken of CENTRED values and converted back in fp64 (csrc/lfa_cl_bwd.cu MODE 4): the
# moments feed BatchNorm statistics (var = E[z^2] - E[z]^2 cancels), where the tensor core's systematic accumulate
# truncation on raw second moments showed up as 1e-4-level deviations in the encoding-MLP gradients
TC_MOM_WIDTHS = (16, 64, 128)
TC_ALL_WIDTHS = (16, 32, 64, 128, 256)                 # what the kernels are built for (tests run all of them)
if os.environ.get("R3D_TC_OFF"):                      # tuning override: FP32 CUDA-core LFA kernels everywhere
    USE_TENSOR_CORES = False
if os.environ.get("R3D_TC_WIDTHS"):                   # tuning override: "fwd widths;bwd widths", e.g. "64,128;16,64,128"
    _f, _, _b = os.environ["R3D_TC_WIDTHS"].partition(";")
    TC_WIDTHS = tuple(int(v) for v in _f.split(",") if v)
    TC_BWD_WIDTHS = tuple(int(v) for v in (_b or _f).split(",") if v)
_TC_STATUS = {}


# Launches too small to amortise a tensor-core kernel's fixed cost (weight image, TMEM folds, the final d x d atomics of
# up to 148 CTAs: ~20 us forward, ~90 us backward) stay on the FP32 kernels.  Thresholds in POINTS (B*N) per launch,
# from the 8 x 2 500-point step (profiles/r02_tc_vs_fp32_train2500.txt): forward pays from ~1 k points at d = 64/128,
# the backward family from ~16 k (d = 16: ~130 k), d = 256 backward always.  TC_MIN_POINTS = 0 forces (tests).
TC_MIN_POINTS = {"fwd": {64: 1024, 128: 512, 256: 1024, 16: 1 << 30, 32: 1 << 30},
                 "bwd": {16: 131072, 32: 32768, 64: 16384, 128: 4096, 256: 0},
                 "mom": {16: 131072, 32: 32768, 64: 16384, 128: 4096}}
TC_FORCE = False


def _tc_big_enough(kind: str, d: int, points) -> bool:
    return TC_FORCE or points is None or points >= TC_MIN_POINTS[kind].get(d, 0)


def lfa_pool_tc_supported(d: int, k: int, points=None) -> bool:
    return USE_TENSOR_CORES and d in TC_WIDTHS and k in TC_NEIGHBORS and _tc_big_enough("fwd", d, points)


def lfa_bwd_tc_supported(d: int, k: int, points=None) -> bool:
    return USE_TENSOR_CORES and d in TC_BWD_WIDTHS and k in TC_NEIGHBORS and _tc_big_enough("bwd", d, points)


def lfa_mom_tc_supported(d: int, k: int, points=None) -> bool:
    return (USE_TENSOR_CORES and d in TC_MOM_WIDTHS and d <= 128 and k in TC_NEIGHBORS
            and _tc_big_enough("mom", d, points))


def _lfa_tc_wide(mode: int, name: str, flops: float, nbytes: float, xyz, xs, idx32, feat, fs, w_rpe1, a_rpe1, b_rpe1,
                 w_score, rmat=None, pooled=None, dpooled=None, dfeat=None, dws=None, g1=None, du2_part=None, sums=None,
                 scal=None):
    """One launch of the d = 256 tensor-core kernel (C ABI ``r3d_lfa_tc_wide``, csrc/lfa_cl_wide.cu)."""
    B, N, K = idx32.shape
    dev = xyz.device
    with torch.cuda.device(dev), _cabi.kernel_timer(f"{name}[N={N},d=256]", flops=flops, bytes=nbytes):
        rc = _cabi.lib().r3d_lfa_tc_wide(
            mode, _cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.raw(feat), fs, _cabi.ptr(w_rpe1), _cabi.ptr(a_rpe1),
            _cabi.ptr(b_rpe1), _cabi.ptr(rmat), _cabi.ptr(w_score), _cabi.ptr(pooled), _cabi.ptr(dpooled),
            _cabi.raw(dfeat), 0, _cabi.raw(dws), _cabi.raw(g1), _cabi.ptr(du2_part), _cabi.raw(sums), _cabi.raw(scal),
            _cabi.ptr(tc_status(dev)), B, N, K, 256, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_tc_wide")


def lfa_r2_rows(xyz, xs, idx32, w_rpe1, a_rpe1, b_rpe1, w_rpe2, a_rpe2, b_rpe2) -> torch.Tensor:
    """r2 = relu(a2 (W2 r1) + c2) for every (point, neighbour) row, (B*N*K, h): the materialised encoding the d = 256
    kernels read (``r3d_lfa_r1_rows`` + one per-point layer).  w_rpe2 (h,h) [out][in]."""
    B, N, K = idx32.shape
    h = w_rpe1.shape[0]
    dev = xyz.device
    rows = B * N * K
    r1 = torch.empty((rows, h), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_r1_rows[N={N},h={h}]", flops=22.0 * rows * h, bytes=4.0 * rows * h):
        rc = _cabi.lib().r3d_lfa_r1_rows(_cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.ptr(w_rpe1), _cabi.ptr(a_rpe1),
                                         _cabi.ptr(b_rpe1), _cabi.ptr(r1), B, N, K, h, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_r1_rows")
    if pc_gemm_ok(rows, h, h):
        return pc_gemm(r1, w_rpe2.contiguous(), scale=a_rpe2, shift=b_rpe2, act="relu")
    return pointwise(r1.unsqueeze(0), w_rpe2.contiguous(), a_rpe2, b_rpe2, act="relu", w_out_in=True).squeeze(0)


def tc_status(device) -> torch.Tensor:
    """Per-device int32 status word of the tensor-core kernels (bit 0: an activation left the range of the fixed
    split-fp16 operand scale, |x| >= 4094).  Allocated once, never reset by the kernels."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    t = _TC_STATUS.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", key))
        _TC_STATUS[key] = t
    return t


def check_tc_status(device) -> None:
    """Raise if a tensor-core LFA kernel has reported an out-of-range activation since the last check (one device ->
    host read: call it where the host synchronises anyway, e.g. after reading the loss or the predictions)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    t = _TC_STATUS.get(key)
    if t is not None and int(t.item()) != 0:
        t.zero_()
        raise FloatingPointError("a fused LFA tensor-core kernel saw an activation with |x| >= 4094: outside the range "
                                 "of its split-fp16 operands (set ops.USE_TENSOR_CORES = False for this model)")


def lfa_pool_tc(stage: int, xyz: torch.Tensor, idx32: torch.Tensor, feat: torch.Tensor, w_rpe1, a_rpe1, b_rpe1,
                w_rpe2, a_rpe2, b_rpe2, w_score, cache: Optional[dict] = None) -> torch.Tensor:
    """``lfa_pool`` on the tensor cores (C ABI ``r3d_lfa_pool_tc``; ``r3d_lfa_tc_wide`` for d = 256): weights in their
    stored [out][in] layout.  ``cache``: a dict the d = 256 stage-2 forward leaves its r2 rows in for the backward."""
    _cabi.require_cuda(xyz, "xyz")
    xyz, xs = _cloud_view(xyz)
    feat, fs = _rows_view(feat.detach())
    B, N, K = idx32.shape
    h = feat.shape[2]
    d = 2 * h
    dev = xyz.device
    pooled = torch.empty((B, N, d), dtype=torch.float32, device=dev)
    flops = float(B) * N * (2 * K * (10 * h + d * d + d + (h * h if stage == 2 else 0)))
    nbytes = float(B) * N * (12 + 4 * K + 4 * h + 4 * d)
    if d == 256:
        rmat = None
        if stage == 2:
            rmat = lfa_r2_rows(xyz, xs, idx32, w_rpe1, a_rpe1, b_rpe1, w_rpe2, a_rpe2, b_rpe2)
            if cache is not None:
                cache["r2"] = rmat
        _lfa_tc_wide(0, f"lfa_cl_fwd{stage}", flops, nbytes, xyz, xs, idx32, feat, fs, w_rpe1, a_rpe1, b_rpe1,
                     w_score.contiguous(), rmat=rmat, pooled=pooled)
        return pooled
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_cl_fwd{stage}[N={N},d={d}]", flops=flops, bytes=nbytes):
        rc = _cabi.lib().r3d_lfa_pool_tc(stage, _cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.raw(feat), fs,
                                         _cabi.ptr(w_rpe1), _cabi.ptr(a_rpe1), _cabi.ptr(b_rpe1), _cabi.ptr(w_rpe2),
                                         _cabi.ptr(a_rpe2), _cabi.ptr(b_rpe2), _cabi.ptr(w_score), _cabi.ptr(pooled),
                                         _cabi.ptr(tc_status(dev)), B, N, K, d, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_pool_tc")
    return pooled


def lfa_pool(stage: int, xyz: torch.Tensor, idx32: torch.Tensor, feat: torch.Tensor, w_rpe1, a_rpe1, b_rpe1,
             w_rpe2T, a_rpe2, b_rpe2, w_scoreT) -> torch.Tensor:
    """Fused LocSE + attentive pooling of one LFA half (C ABI ``r3d_lfa_pool``; modules.py:316-323).
    xyz (B,N,3), idx32 (B,N,K) int32, feat (B,N,h) -> pooled (B,N,d), d = 2h."""
    _cabi.require_cuda(xyz, "xyz")
    if not lfa_fused_supported(2 * feat.shape[2], idx32.shape[2]):
        return lfa_pool_rows(stage, xyz, idx32, feat, w_rpe1, a_rpe1, b_rpe1, w_rpe2T, a_rpe2, b_rpe2, w_scoreT)
    xyz, xs = _cloud_view(xyz)
    feat, fs = _rows_view(feat.detach())
    B, N, K = idx32.shape
    h = feat.shape[2]
    d = 2 * h
    dev = xyz.device
    pooled = torch.empty((B, N, d), dtype=torch.float32, device=dev)
    # algorithmic work per point: flops 2K(10h + d^2 + d) [+ 2K h^2 in stage 2]
    flops = float(B) * N * (2 * K * (10 * h + d * d + d + (h * h if stage == 2 else 0)))
    nbytes = float(B) * N * (12 + 4 * K + 4 * h + 4 * d)
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_pool{stage}[N={N},d={d}]", flops=flops, bytes=nbytes):
        rc = _cabi.lib().r3d_lfa_pool(stage, _cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.raw(feat), fs,
                                      _cabi.ptr(w_rpe1), _cabi.ptr(a_rpe1), _cabi.ptr(b_rpe1), _cabi.ptr(w_rpe2T),
                                      _cabi.ptr(a_rpe2), _cabi.ptr(b_rpe2), _cabi.ptr(w_scoreT), _cabi.ptr(pooled),
                                      B, N, K, d, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_pool")
    return pooled


_ACT = {None: 0, "none": 0, "relu": 1, "lrelu": 2}


def pointwise(xa: torch.Tensor, wT: torch.Tensor, scale=None, shift=None, act=None, slope: float = 0.0,
              gidx: Optional[torch.Tensor] = None, xb: Optional[torch.Tensor] = None, n_rows: Optional[int] = None,
              transpose_out: bool = False, out: Optional[torch.Tensor] = None,
              stats: Optional[torch.Tensor] = None, w_out_in: bool = False) -> torch.Tensor:
    """Per-point layer y = act(scale * (W [xa[g] ; xb]) + shift) (C ABI ``r3d_pointwise``).
    xa (B,na,ca) [gathered through gidx int32 (n,) shared or (B,n) per cloud], xb (B,n,cb) optional,
    wT (ca+cb, cout).  Returns (B,n,cout), or (B,cout,n) with ``transpose_out``."""
    _cabi.require_cuda(xa, "xa")
    xa, xas = _rows_view(xa.detach())
    B, na, ca = xa.shape
    n = n_rows if n_rows is not None else (gidx.shape[-1] if gidx is not None else na)
    cb, xbs = 0, 0
    if xb is not None:
        xb, xbs = _rows_view(xb.detach())
        cb = xb.shape[2]
        assert xb.shape[1] >= n
    cout = wT.shape[0] if w_out_in else wT.shape[1]
    assert wT.shape[1 if w_out_in else 0] == ca + cb and wT.is_contiguous()
    dev = xa.device
    if out is None:
        out = torch.empty((B, cout, n) if transpose_out else (B, n, cout), dtype=torch.float32, device=dev)
    gs = 0
    if gidx is not None:
        assert gidx.dtype == torch.int32 and gidx.is_contiguous()
        gs = 0 if gidx.dim() == 1 else gidx.stride(0)
    flops = 2.0 * B * n * (ca + cb) * cout
    nbytes = 4.0 * B * n * (ca + cb + cout)
    tname = "pointwise"
    if _cabi.KERNEL_TIMERS is not None:
        tname = ("pw_small", "pw_gemm", "pw_gemm_fast", "pw_tc", "pw_rows", "pw_expand")[
            _cabi.lib().r3d_pointwise_plan(ca, cb, cout, B * n, 1 if transpose_out else 0)]
        if _cabi.TIMER_SHAPES:
            tname += f"[M={B * n},{ca + cb}->{cout}]"
    with torch.cuda.device(dev), _cabi.kernel_timer(tname, flops=flops, bytes=nbytes):
        rc = _cabi.lib().r3d_pointwise_stats(_cabi.raw(xa), xas, ca, _cabi.ptr(gidx), gs, _cabi.raw(xb), xbs, cb,
                                             _cabi.ptr(wT), _cabi.ptr(scale), _cabi.ptr(shift), _ACT[act],
                                             float(slope), _cabi.ptr(out), 0, 0, cout, B, n,
                                             1 if transpose_out else 0, _cabi.ptr(stats), 1 if w_out_in else 0,
                                             _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_pointwise")
    return out


# Per-point layers on the tensor cores (csrc/pw_cl.cu).  Measured against the FP32 kernels on the train40960 shapes
# (tools/pw_cl_bench.py): ahead wherever both widths reach 32 (1.5x at 2.6 M rows x 32<->64, 2-3.5x on the mid-width layers
# of levels 1-3 and the decoder); narrower layers are bound by memory traffic and stay on pw_rows / rowreduce_gemm_narrow.
# C_in > 128 runs as one pass per 128-channel block with read-modify-write of y: correct (tested) but slower than pw_tc
# on the decoder's 256...1024-channel inputs, so the dispatch stops at 128.
PC_MIN_ROWS = 8192


def pc_gemm_ok(M: int, cin: int, cout: int) -> bool:
    return (USE_TENSOR_CORES and M >= PC_MIN_ROWS and cin % 16 == 0 and 32 <= cin <= 128 and cout >= 32
            and cin + cout >= 96)


def pc_wgrad_ok(M: int, ca: int, cb: int) -> bool:
    return (USE_TENSOR_CORES and M >= PC_MIN_ROWS and ca % 8 == 0 and cb % 8 == 0 and max(ca, cb) >= 64
            and min(ca, cb) >= 32)


def pc_gemm(x: torch.Tensor, w: torch.Tensor, transposed: bool = False, scale=None, shift=None, act=None,
            slope: float = 0.0, stats: Optional[torch.Tensor] = None, absmax_out: Optional[torch.Tensor] = None
            ) -> torch.Tensor:
    """Per-point layer on the tensor cores (C ABI ``r3d_pc_gemm``): x (M,cin) dense rows, w the layer's (cout,cin)
    weight -> act(scale * (x w^T) + shift) (M,cout); ``transposed``: the input gradient x (M,cout) -> x w (M,cin).
    ``stats`` (2*channels fp64, zero-filled) += batch sums of x w^T and its square; ``absmax_out`` (1,) fp32 zero-filled
    receives max |x| (the weight-gradient kernel's operand scale)."""
    _cabi.require_cuda(x, "x")
    assert x.dim() == 2 and x.stride(1) == 1 and w.is_contiguous()
    M, cin = x.shape
    cout = w.shape[1] if transposed else w.shape[0]
    assert (w.shape[0] if transposed else w.shape[1]) == cin
    so, si = (1, w.shape[1]) if transposed else (w.shape[1], 1)
    y = torch.empty((M, cout), dtype=torch.float32, device=x.device)
    name = f"pc_gemm[M={M},{cin}->{cout}]" if _cabi.TIMER_SHAPES else "pc_gemm"
    with torch.cuda.device(x.device), _cabi.kernel_timer(name, flops=2.0 * M * cin * cout, bytes=4.0 * M * (cin + cout)):
        rc = _cabi.lib().r3d_pc_gemm(_cabi.raw(x.detach()), x.stride(0), _cabi.ptr(w.detach()), so, si, _cabi.ptr(scale),
                                     _cabi.ptr(shift), _ACT[act], float(slope), _cabi.ptr(y), cout, _cabi.ptr(stats),
                                     _cabi.ptr(absmax_out), M, cin, cout, _cabi.stream_ptr(x.device))
    _cabi.check(rc, "r3d_pc_gemm")
    return y


def pc_wgrad(a: torch.Tensor, b: torch.Tensor, absmax_a: Optional[torch.Tensor] = None,
             absmax_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a (M,Ca), b (M,Cb) dense rows -> a^T b (Ca,Cb) on the tensor cores (C ABI ``r3d_pc_wgrad``).  absmax_*: (1,) fp32
    device bounds of |a|, |b|; computed here (``r3d_absmax``) when not supplied by the tensors' producers."""
    _cabi.require_cuda(a, "a")
    a, b = a.contiguous(), b.contiguous()
    M, Ca = a.shape
    Cb = b.shape[1]
    if absmax_a is None or absmax_b is None:
        slots = zeros((2,), torch.float32, a.device)
        if absmax_a is None:
            absmax_a = slots[0:1]
            _absmax_into(a, absmax_a)
        if absmax_b is None:
            absmax_b = slots[1:2]
            _absmax_into(b, absmax_b)
    out = zeros((Ca, Cb), torch.float32, a.device)
    name = f"pc_wgrad[M={M},{Ca}x{Cb}]" if _cabi.TIMER_SHAPES else "pc_wgrad"
    with torch.cuda.device(a.device), _cabi.kernel_timer(name, flops=2.0 * M * Ca * Cb, bytes=4.0 * M * (Ca + Cb)):
        rc = _cabi.lib().r3d_pc_wgrad(_cabi.ptr(a), Ca, Ca, _cabi.ptr(b), Cb, Cb, M, _cabi.raw(absmax_a), _cabi.raw(absmax_b),
                                      _cabi.ptr(out), Cb, _cabi.stream_ptr(a.device))
    _cabi.check(rc, "r3d_pc_wgrad")
    return out


# ------------------------------------------------------------------------- row-form LFA block (any K, any width)
# Shapes the fused LocSE + pooling kernels are instantiated for; everything else runs in row form (csrc/lfa_rows.cu).
LFA_FUSED_WIDTHS, LFA_FUSED_NEIGHBORS = (16, 32, 64, 128, 256), (16, 32)


def lfa_fused_supported(d: int, k: int) -> bool:
    return d in LFA_FUSED_WIDTHS and k in LFA_FUSED_NEIGHBORS


def lfa_rpe_rows(xyz: torch.Tensor, idx32: torch.Tensor) -> torch.Tensor:
    """(B*N*K, 10) rows [p_i, p_j, p_i - p_j, |p_i - p_j|] (C ABI ``r3d_lfa_rpe_rows``; modules.py:170-186)."""
    _cabi.require_cuda(xyz, "xyz")
    xyz, xs = _cloud_view(xyz)
    idx32 = idx32.contiguous()
    B, N, K = idx32.shape
    out = torch.empty((B * N * K, 10), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device), _cabi.kernel_timer("lfa_rpe_rows", flops=8.0 * B * N * K, bytes=68.0 * B * N * K):
        rc = _cabi.lib().r3d_lfa_rpe_rows(_cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.ptr(out), B, N, K,
                                          _cabi.stream_ptr(xyz.device))
    _cabi.check(rc, "r3d_lfa_rpe_rows")
    return out


def lfa_gather_concat(r: torch.Tensor, feat: torch.Tensor, idx32: torch.Tensor) -> torch.Tensor:
    """r (B*N*K,h) rows, feat (B,N,h), idx32 (B,N,K) -> X (B*N*K,2h) = [r ; feat at the neighbours]
    (C ABI ``r3d_lfa_gather_concat``; PointFeatureAugmentation, modules.py:200-208)."""
    _cabi.require_cuda(r, "r")
    feat, fs = _rows_view(feat.detach())
    r = r.detach().contiguous()
    idx32 = idx32.contiguous()
    B, N, K = idx32.shape
    h = feat.shape[2]
    assert r.shape == (B * N * K, h) and r.dtype == torch.float32
    out = torch.empty((B * N * K, 2 * h), dtype=torch.float32, device=r.device)
    with torch.cuda.device(r.device), _cabi.kernel_timer("lfa_gather_concat", flops=0.0, bytes=16.0 * B * N * K * h):
        rc = _cabi.lib().r3d_lfa_gather_concat(_cabi.ptr(r), _cabi.raw(feat), fs, _cabi.ptr(idx32), _cabi.ptr(out), B, N, K,
                                               h, _cabi.stream_ptr(r.device))
    _cabi.check(rc, "r3d_lfa_gather_concat")
    return out


def lfa_gather_concat_bwd(dout: torch.Tensor, idx32: torch.Tensor, need_r: bool, need_feat: bool):
    """Backward of ``lfa_gather_concat``: (dr (B*N*K,h) | None, dfeat (B,N,h) | None)."""
    dout = dout.contiguous()
    idx32 = idx32.contiguous()
    B, N, K = idx32.shape
    h = dout.shape[1] // 2
    dev = dout.device
    dr = torch.empty((B * N * K, h), dtype=torch.float32, device=dev) if need_r else None
    dfeat = zeros((B, N, h), torch.float32, dev) if need_feat else None
    with torch.cuda.device(dev), _cabi.kernel_timer("lfa_gather_concat_bwd", flops=float(B * N * K * h),
                                                    bytes=16.0 * B * N * K * h):
        rc = _cabi.lib().r3d_lfa_gather_concat_bwd(_cabi.ptr(dout), _cabi.ptr(idx32), _cabi.ptr(dr), _cabi.ptr(dfeat), 0,
                                                   B, N, K, h, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_gather_concat_bwd")
    return dr, dfeat


def lfa_attn_pool(S: torch.Tensor, X: torch.Tensor, K: int) -> torch.Tensor:
    """S, X (points*K, d) -> pooled (points, d) = sum_k softmax_k(S) X (C ABI ``r3d_lfa_attn_pool``; modules.py:246-252)."""
    _cabi.require_cuda(S, "S")
    S, X = S.detach().contiguous(), X.detach().contiguous()
    rows, d = S.shape
    assert X.shape == S.shape and rows % K == 0
    pooled = torch.empty((rows // K, d), dtype=torch.float32, device=S.device)
    with torch.cuda.device(S.device), _cabi.kernel_timer("lfa_attn_pool", flops=6.0 * rows * d, bytes=8.0 * rows * d):
        rc = _cabi.lib().r3d_lfa_attn_pool(_cabi.ptr(S), _cabi.ptr(X), _cabi.ptr(pooled), rows // K, K, d,
                                           _cabi.stream_ptr(S.device))
    _cabi.check(rc, "r3d_lfa_attn_pool")
    return pooled


def lfa_attn_pool_bwd(S: torch.Tensor, X: torch.Tensor, dpooled: torch.Tensor, K: int):
    """Backward of ``lfa_attn_pool``: (dS, dX) with dX the direct term g A only (C ABI ``r3d_lfa_attn_pool_bwd``)."""
    S, X, dpooled = S.contiguous(), X.contiguous(), dpooled.contiguous()
    rows, d = S.shape
    dS, dX = torch.empty_like(S), torch.empty_like(X)
    with torch.cuda.device(S.device), _cabi.kernel_timer("lfa_attn_pool_bwd", flops=12.0 * rows * d, bytes=24.0 * rows * d):
        rc = _cabi.lib().r3d_lfa_attn_pool_bwd(_cabi.ptr(S), _cabi.ptr(X), _cabi.ptr(dpooled), _cabi.ptr(dS), _cabi.ptr(dX),
                                               rows // K, K, d, _cabi.stream_ptr(S.device))
    _cabi.check(rc, "r3d_lfa_attn_pool_bwd")
    return dS, dX


def lfa_pool_rows(stage: int, xyz, idx32, feat, w_rpe1, a_rpe1, b_rpe1, w_rpe2T, a_rpe2, b_rpe2, w_scoreT) -> torch.Tensor:
    """``lfa_pool`` in row form for shapes outside the fused kernels' template lists (same arguments and result):
    rpe rows -> mlp_rpe1 [-> mlp_rpe2] as per-point layers over the B*N*K rows -> gather + concat -> score Linear as a
    per-point layer -> softmax over K and weighted sum.  Inference only (the training path composes the same pieces
    under autograd, engine.lfa_block_rows)."""
    B, N, K = idx32.shape
    rows = lfa_rpe_rows(xyz, idx32)
    r = pointwise(rows.unsqueeze(0), w_rpe1.contiguous(), a_rpe1, b_rpe1, "relu", w_out_in=True)
    if stage == 2:
        r = pointwise(r, w_rpe2T.contiguous(), a_rpe2, b_rpe2, "relu")
    X = lfa_gather_concat(r.squeeze(0), feat, idx32)
    S = pointwise(X.unsqueeze(0), w_scoreT.contiguous()).squeeze(0)
    return lfa_attn_pool(S, X, K).view(B, N, -1)


# ------------------------------------------------------------------------- element-wise pieces of the training step
def add_lrelu(a: torch.Tensor, b: torch.Tensor, slope: float) -> torch.Tensor:
    """LeakyReLU_slope(a + b) in one launch (C ABI ``r3d_add_lrelu``): the residual sum closing an LFA block."""
    _cabi.require_cuda(a, "a")
    a, b = a.detach().contiguous(), b.detach().contiguous()
    assert a.shape == b.shape and a.dtype == b.dtype == torch.float32
    y = torch.empty_like(a)
    n = a.numel()
    with torch.cuda.device(a.device), _cabi.kernel_timer("add_lrelu", flops=2.0 * n, bytes=12.0 * n):
        rc = _cabi.lib().r3d_add_lrelu(_cabi.ptr(a), _cabi.ptr(b), _cabi.ptr(y), n, float(slope), _cabi.stream_ptr(a.device))
    _cabi.check(rc, "r3d_add_lrelu")
    return y


def add_lrelu_bwd(dy: torch.Tensor, y: torch.Tensor, slope: float) -> torch.Tensor:
    """dy * LeakyReLU'(a + b) from the sign of y (C ABI ``r3d_add_lrelu_bwd``): the gradient of both summands."""
    dy = dy.contiguous()
    d = torch.empty_like(y)
    n = y.numel()
    with torch.cuda.device(y.device), _cabi.kernel_timer("add_lrelu_bwd", flops=1.0 * n, bytes=12.0 * n):
        rc = _cabi.lib().r3d_add_lrelu_bwd(_cabi.ptr(dy), _cabi.ptr(y), _cabi.ptr(d), n, float(slope), _cabi.stream_ptr(y.device))
    _cabi.check(rc, "r3d_add_lrelu_bwd")
    return d


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: torch.Tensor, lr, beta1: float,
              beta2: float, eps: float) -> None:
    """torch.optim.Adam's update over flat fp32 buffers, in place (C ABI ``r3d_adam_step``).  ``step``: device scalar
    (fp32) advanced by one; ``lr``: a float or a device scalar tensor (a scheduler writes into it)."""
    _cabi.require_cuda(p, "p")
    n = p.numel()
    assert g.numel() == m.numel() == v.numel() == n and all(t.is_contiguous() and t.dtype == torch.float32 for t in (p, g, m, v))
    assert step.dtype == torch.float32 and step.is_cuda
    lr_dev = lr if isinstance(lr, torch.Tensor) else None
    if lr_dev is not None:
        assert lr_dev.is_cuda and lr_dev.dtype == torch.float32
    with torch.cuda.device(p.device), _cabi.kernel_timer("adam_step", flops=12.0 * n, bytes=28.0 * n):
        rc = _cabi.lib().r3d_adam_step(_cabi.raw(p), _cabi.raw(g), _cabi.raw(m), _cabi.raw(v), n, _cabi.raw(lr_dev),
                                       0.0 if lr_dev is not None else float(lr), float(beta1), float(beta2), float(eps),
                                       _cabi.raw(step), _cabi.stream_ptr(p.device))
    _cabi.check(rc, "r3d_adam_step")


UP_WEIGHTING = {"nni": (0, 1.0), "nna": (1, 1.0), "idw": (1, 1.0), "isdw": (1, 2.0), "mean": (2, 1.0)}


def _up_args(approach: str, idx: torch.Tensor, dist: Optional[torch.Tensor]):
    if approach not in UP_WEIGHTING:
        raise ValueError(f"Upsampling approach {approach} not understood!")
    weighting, power = UP_WEIGHTING[approach]
    if idx.dtype not in (torch.int32, torch.int64) or idx.dim() != 3:
        raise ValueError("idx must be an int32 / int64 tensor of shape (B, N2, K)")
    idx = idx.contiguous()
    if weighting == 1:
        if dist is None or dist.shape != idx.shape:
            raise ValueError("the inverse-distance approaches need dist (B, N2, K)")
        dist = dist.float().contiguous()
    return weighting, power, idx, dist


def upsample(approach: str, feat: torch.Tensor, idx: torch.Tensor, dist: Optional[torch.Tensor] = None,
             skip: Optional[torch.Tensor] = None, channel_major: bool = False) -> torch.Tensor:
    """UpSampler gather (modules.py:343-414) [+ decoder skip concat, :600-602] in one launch (C ABI ``r3d_upsample``).
    feat (B,N1,F), idx (B,N2,K) neighbours of the fine points among the coarse ones, dist (B,N2,K) for nna/idw/isdw,
    skip (B,N2,Fs) -> (B,N2,F+Fs), or (B,F,N2) with ``channel_major``."""
    _cabi.require_cuda(feat, "feat")
    weighting, power, idx, dist = _up_args(approach, idx, dist)
    feat, fbs = _rows_view(feat.detach())
    B, N1, F = feat.shape
    N2, K = idx.shape[1], idx.shape[2]
    Fs, sbs = 0, 0
    if skip is not None:
        skip, sbs = _rows_view(skip.detach())
        Fs = skip.shape[2]
        assert skip.shape[1] >= N2
    dev = feat.device
    out = torch.empty((B, F, N2) if channel_major else (B, N2, F + Fs), dtype=torch.float32, device=dev)
    kk = 1 if weighting == 0 else K
    with torch.cuda.device(dev), _cabi.kernel_timer("upsample", flops=2.0 * B * N2 * F * kk,
                                                    bytes=4.0 * B * N2 * (F * (kk + 1) + 2 * Fs + 2 * kk)):
        rc = _cabi.lib().r3d_upsample(_cabi.raw(feat), fbs, F, F, _cabi.ptr(idx), int(idx.dtype == torch.int64),
                                      _cabi.ptr(dist), K, weighting, power, _cabi.raw(skip), sbs, Fs, Fs, _cabi.ptr(out),
                                      0, 0, 1 if channel_major else 0, B, N1, N2, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_upsample")
    return out


def upsample_bwd(approach: str, dout: torch.Tensor, idx: torch.Tensor, dist: Optional[torch.Tensor], n_coarse: int,
                 n_feat: int, need_skip: bool = True):
    """Backward of ``upsample``: dout (B,N2,F+Fs) -> (dfeat (B,N1,F) scatter-added, dskip (B,N2,Fs) or None)
    (C ABI ``r3d_upsample_bwd``)."""
    _cabi.require_cuda(dout, "dout")
    weighting, power, idx, dist = _up_args(approach, idx, dist)
    dout, dbs = _rows_view(dout.detach())
    B, N2, C = dout.shape
    F, Fs, K = n_feat, C - n_feat, idx.shape[2]
    dev = dout.device
    dfeat = zeros((B, n_coarse, F), torch.float32, dev)
    dskip = torch.empty((B, N2, Fs), dtype=torch.float32, device=dev) if (need_skip and Fs) else None
    kk = 1 if weighting == 0 else K
    with torch.cuda.device(dev), _cabi.kernel_timer("upsample_bwd", flops=2.0 * B * N2 * F * kk,
                                                    bytes=4.0 * B * N2 * (F * (kk + 1) + 2 * Fs + 2 * kk)):
        rc = _cabi.lib().r3d_upsample_bwd(_cabi.raw(dout), dbs, C, _cabi.ptr(idx), int(idx.dtype == torch.int64),
                                          _cabi.ptr(dist), K, weighting, power, _cabi.ptr(dfeat), 0, 0, F,
                                          _cabi.ptr(dskip), 0, 0, Fs, B, n_coarse, N2, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_upsample_bwd")
    return dfeat, dskip


def bn_apply(z: torch.Tensor, stats: torch.Tensor, bn: torch.nn.BatchNorm2d, bias: Optional[torch.Tensor], act,
             slope: float = 0.0, track: bool = True):
    """Train-mode BatchNorm + activation of a per-point layer (C ABI ``r3d_bn_apply``): z (M,C) conv output without
    bias, stats (2C) fp64 from ``pointwise(..., stats=)``.  Updates bn's running statistics in place.
    Returns (y (M,C), save (3,C) = a, mean, rstd)."""
    M, C = z.shape
    y = torch.empty_like(z)
    save = torch.empty((3, C), dtype=torch.float32, device=z.device)
    track = track and bn.track_running_stats and bn.running_mean is not None
    with torch.cuda.device(z.device), _cabi.kernel_timer(f"bn_apply[M={M},C={C}]", flops=4.0 * M * C, bytes=8.0 * M * C):
        rc = _cabi.lib().r3d_bn_apply(_cabi.ptr(z), _cabi.ptr(stats), M, C, _cabi.ptr(bn.weight.detach()),
                                      _cabi.ptr(bn.bias.detach()), _cabi.ptr(bias.detach()) if bias is not None else None,
                                      float(bn.eps), float(bn.momentum),
                                      _cabi.ptr(bn.running_mean) if track else None,
                                      _cabi.ptr(bn.running_var) if track else None,
                                      _cabi.ptr(bn.num_batches_tracked) if track else None, _ACT[act], float(slope),
                                      _cabi.ptr(y), _cabi.ptr(save), _cabi.stream_ptr(z.device))
    _cabi.check(rc, "r3d_bn_apply")
    return y, save


def pointwise_bn(x: torch.Tensor, w: torch.Tensor, stats: torch.Tensor, bn: torch.nn.BatchNorm2d,
                 bias: Optional[torch.Tensor], act, slope: float = 0.0):
    """Train-mode SharedMLP forward on dense rows (C ABI ``r3d_pointwise_bn``): x (M,cin), w (cout,cin), stats (2*cout)
    zero-filled fp64.  One cooperative launch when the layer's grid is co-resident, GEMM + ``bn_apply`` otherwise.
    Updates bn's running statistics in place.  Returns (z (M,cout), y (M,cout), save (3,cout))."""
    _cabi.require_cuda(x, "x")
    M, cin = x.shape
    cout = w.shape[0]
    dev = x.device
    if not (_cabi.lib().r3d_bn_set_fused(-1) & 1):
        # default: two launches, each under its own name in bench.py's kernel table
        z = pointwise(x.unsqueeze(0), w, stats=stats, w_out_in=True).squeeze(0)
        y, save = bn_apply(z, stats, bn, bias, act, slope)
        return z, y, save
    z = torch.empty((M, cout), dtype=torch.float32, device=dev)
    y = torch.empty((M, cout), dtype=torch.float32, device=dev)
    save = torch.empty((3, cout), dtype=torch.float32, device=dev)
    track = bn.track_running_stats and bn.running_mean is not None
    tname = "pointwise_bn"
    if _cabi.KERNEL_TIMERS is not None:
        tname = ("pw_small", "pw_gemm", "pw_gemm_fast", "pw_tc", "pw_rows", "pw_expand")[_cabi.lib().r3d_pointwise_plan(cin, 0, cout, M, 0)]
        if _cabi.TIMER_SHAPES:
            tname += f"[M={M},{cin}->{cout},bn]"
    with torch.cuda.device(dev), _cabi.kernel_timer(tname, flops=2.0 * M * cin * cout + 4.0 * M * cout,
                                                    bytes=4.0 * M * (cin + 2 * cout)):
        rc = _cabi.lib().r3d_pointwise_bn(
            _cabi.ptr(x), M, cin, _cabi.ptr(w), cout, _cabi.ptr(stats), _cabi.ptr(bn.weight.detach()),
            _cabi.ptr(bn.bias.detach()), _cabi.ptr(bias.detach()) if bias is not None else None, float(bn.eps),
            float(bn.momentum), _cabi.ptr(bn.running_mean) if track else None,
            _cabi.ptr(bn.running_var) if track else None, _cabi.ptr(bn.num_batches_tracked) if track else None,
            _ACT[act], float(slope), _cabi.ptr(z), _cabi.ptr(y), _cabi.ptr(save), _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_pointwise_bn")
    return z, y, save


def bn_backward(dy: torch.Tensor, z: torch.Tensor, save: torch.Tensor, beta: torch.Tensor, act, slope: float = 0.0,
                stats2: Optional[torch.Tensor] = None, absmax_out: Optional[torch.Tensor] = None):
    """BatchNorm(+activation) backward with batch statistics (C ABI ``r3d_bn_bwd_absmax``: the reduce and dz passes, one
    cooperative launch for small tensors).
    ``stats2``: optional zero-filled (2C) fp64 scratch (the forward allocates it together with its own statistics
    buffer: one fill instead of two); ``absmax_out``: optional zero-filled (1,) fp32 that receives max |dz|.
    Returns (dz (M,C), dgamma (C), dbeta (C))."""
    M, C = z.shape
    dy = dy.contiguous()
    if stats2 is None:
        stats2 = zeros(2 * C, torch.float64, z.device)
    dz = torch.empty_like(z)
    s2 = torch.empty(2 * C, dtype=torch.float32, device=z.device)
    L = _cabi.lib()
    with torch.cuda.device(z.device), _cabi.kernel_timer(f"bn_backward[M={M},C={C}]", flops=12.0 * M * C, bytes=20.0 * M * C):
        rc = L.r3d_bn_bwd_absmax(_cabi.ptr(dy), _cabi.ptr(z), M, C, _cabi.ptr(save), _cabi.ptr(beta), _ACT[act],
                                 float(slope), _cabi.ptr(stats2), _cabi.ptr(dz), _cabi.ptr(s2), _cabi.raw(absmax_out),
                                 _cabi.stream_ptr(z.device))
    _cabi.check(rc, "r3d_bn_bwd")
    return dz, s2[C:], s2[:C]


def bn_backward_fixed(dy: torch.Tensor, z: torch.Tensor, save: torch.Tensor, beta: torch.Tensor, act, slope: float = 0.0):
    """Backward of an EVAL-mode BatchNorm(+activation) (fixed statistics): the reduce pass gives dbeta = sum du and
    dgamma = sum du zhat, and the dz pass run on zero batch sums is dz = a du (C ABI ``r3d_bn_bwd_reduce`` +
    ``r3d_bn_bwd_dz``).  Returns (dz (M,C), dgamma (C), dbeta (C))."""
    M, C = z.shape
    dy = dy.contiguous()
    sums = zeros(4 * C, torch.float64, z.device)        # [0:2C] the reduce pass' sums, [2C:4C] stays zero
    dz = torch.empty_like(z)
    L = _cabi.lib()
    with torch.cuda.device(z.device), _cabi.kernel_timer(f"bn_backward[M={M},C={C}]", flops=12.0 * M * C, bytes=20.0 * M * C):
        rc = L.r3d_bn_bwd_reduce(_cabi.ptr(dy), _cabi.ptr(z), M, C, _cabi.ptr(save), _cabi.ptr(beta), _ACT[act], float(slope),
                                 _cabi.ptr(sums), _cabi.stream_ptr(z.device))
        _cabi.check(rc, "r3d_bn_bwd_reduce")
        rc = L.r3d_bn_bwd_dz(_cabi.ptr(dy), _cabi.ptr(z), M, C, _cabi.ptr(save), _cabi.ptr(beta), _ACT[act], float(slope),
                             _cabi.ptr(sums[2 * C:]), _cabi.ptr(dz), None, _cabi.stream_ptr(z.device))
    _cabi.check(rc, "r3d_bn_bwd_dz")
    return dz, sums[C:2 * C].float(), sums[:C].float()


def rowreduce_gemm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a (M,Ca), b (M,Cb) dense -> a^T b (Ca,Cb): the weight gradient of a per-point layer (``r3d_rowreduce_gemm``)."""
    a, b = a.contiguous(), b.contiguous()
    M, Ca = a.shape
    Cb = b.shape[1]
    out = zeros((Ca, Cb), torch.float32, a.device)
    with torch.cuda.device(a.device), _cabi.kernel_timer(f"rowreduce_gemm[M={M},{Ca}x{Cb}]" if _cabi.TIMER_SHAPES else "rowreduce_gemm", flops=2.0 * M * Ca * Cb,
                                                         bytes=4.0 * M * (Ca + Cb)):
        rc = _cabi.lib().r3d_rowreduce_gemm(_cabi.ptr(a), Ca, _cabi.ptr(b), Cb, M, _cabi.ptr(out), Cb,
                                            _cabi.stream_ptr(a.device))
    _cabi.check(rc, "r3d_rowreduce_gemm")
    return out


def _lfa_tc_bwd(mode: int, name: str, flops: float, nbytes: float, xyz, xs, idx32, feat, fs, w_rpe1, a_rpe1, b_rpe1,
                w_rpe2=None, a_rpe2=None, b_rpe2=None, w_score=None, dpooled=None, dfeat=None, dws=None, g1=None,
                du2=None, sums=None, bn2=None, dw2=None, m_r1=None, s_r1=None, scal=None):
    """One launch of the tensor-core backward / moments kernel family (C ABI ``r3d_lfa_tc_bwd``, csrc/lfa_cl_bwd.cu)."""
    B, N, K = idx32.shape
    d = 2 * w_rpe1.shape[0]
    dev = xyz.device
    with torch.cuda.device(dev), _cabi.kernel_timer(f"{name}[N={N},d={d}]", flops=flops, bytes=nbytes):
        rc = _cabi.lib().r3d_lfa_tc_bwd(
            mode, _cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.raw(feat), fs, _cabi.ptr(w_rpe1), _cabi.ptr(a_rpe1),
            _cabi.ptr(b_rpe1), _cabi.ptr(w_rpe2), _cabi.ptr(a_rpe2), _cabi.ptr(b_rpe2), _cabi.ptr(w_score),
            _cabi.ptr(dpooled), _cabi.raw(dfeat), 0, _cabi.raw(dws), _cabi.raw(g1), _cabi.raw(du2), _cabi.raw(sums),
            _cabi.ptr(bn2), _cabi.raw(dw2), _cabi.raw(m_r1), _cabi.raw(s_r1), _cabi.raw(scal),
            _cabi.ptr(tc_status(dev)), B, N, K, d, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_tc_bwd")


def _absmax_into(x: torch.Tensor, out: torch.Tensor) -> None:
    """out[0] = max |x| (C ABI ``r3d_absmax``; ``out`` zero-filled): the gradient-side operand scale of the tensor-core
    backward is a power of two derived from it on the device."""
    with torch.cuda.device(x.device), _cabi.kernel_timer("absmax", flops=float(x.numel()), bytes=4.0 * x.numel()):
        rc = _cabi.lib().r3d_absmax(_cabi.ptr(x), x.numel(), _cabi.raw(out), _cabi.stream_ptr(x.device))
    _cabi.check(rc, "r3d_absmax")


def lfa_pool_bwd(stage: int, xyz, idx32, feat, w_rpe1, a_rpe1, b_rpe1, w_rpe2T, a_rpe2, b_rpe2, w_rpe2s, w_scoreT,
                 w_score, dpooled, g1_acc=None):
    """Backward of ``lfa_pool`` (C ABI ``r3d_lfa_pool_bwd``).  Returns (dfeat (B,N,h), dw_score (d,d)
    [out][in], g1 (h,16) fp64, g2m (h,h) fp64 | None, g2c (h,16) fp64 | None) — see include/r3d_b200.h.
    ``g1_acc``: an (h,16) fp64 buffer to accumulate g1 INTO (the block's two halves share one)."""
    xyz, xs = _cloud_view(xyz)
    feat, fs = _rows_view(feat.detach())
    B, N, K = idx32.shape
    h = feat.shape[2]
    d = 2 * h
    dev = xyz.device
    dpooled = dpooled.contiguous()
    # one zero-filled fp32 block for the feature gradient and the score-weight gradient, one fp64 block for the
    # encoding-MLP accumulators (they cancel against the BatchNorm moment terms and need the extra digits)
    n_df = B * N * h
    acc = zeros(n_df + d * d, torch.float32, dev)
    dfeat = acc[:n_df].view(B, N, h)
    dws = acc[n_df:].view(d, d)
    acc64 = zeros(h * 16 + (h * h + h * 16 if stage == 2 else 0), torch.float64, dev)
    g1 = acc64[:h * 16].view(h, 16) if g1_acc is None else g1_acc
    g2m = g2c = None
    if stage == 2:
        g2m = acc64[h * 16:h * 16 + h * h].view(h, h)
        g2c = acc64[h * 16 + h * h:].view(h, 16)
    flops = float(B) * N * (2 * K * (10 * h + 3 * d * d + d + (3 * h * h if stage == 2 else 0)))
    nbytes = float(B) * N * (12 + 4 * K + 4 * h + 4 * d + 4 * K * h)
    if stage == 1 and d == 256 and lfa_bwd_tc_supported(d, K, B * N):
        scal = zeros(2, torch.float32, dev)
        _absmax_into(dpooled, scal)
        _lfa_tc_wide(1, "lfa_cl_bwd1", flops, nbytes, xyz, xs, idx32, feat, fs, w_rpe1, a_rpe1, b_rpe1, w_score,
                     dpooled=dpooled, dfeat=dfeat, dws=dws, g1=g1, scal=scal)
        return dfeat, dws, g1, g2m, g2c
    if stage == 1 and lfa_bwd_tc_supported(d, K, B * N):
        scal = zeros(2, torch.float32, dev)
        _absmax_into(dpooled, scal)
        _lfa_tc_bwd(1, "lfa_cl_bwd1", flops, nbytes, xyz, xs, idx32, feat, fs, w_rpe1, a_rpe1, b_rpe1, w_score=w_score,
                    dpooled=dpooled, dfeat=dfeat, dws=dws, g1=g1, scal=scal)
        return dfeat, dws, g1, g2m, g2c
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_pool{stage}_bwd[N={N},d={d}]", flops=flops, bytes=nbytes):
        rc = _cabi.lib().r3d_lfa_pool_bwd(
            stage, _cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.raw(feat), fs, _cabi.ptr(w_rpe1), _cabi.ptr(a_rpe1),
            _cabi.ptr(b_rpe1), _cabi.ptr(w_rpe2T), _cabi.ptr(a_rpe2), _cabi.ptr(b_rpe2), _cabi.ptr(w_rpe2s),
            _cabi.ptr(w_scoreT), _cabi.ptr(w_score), _cabi.ptr(dpooled), _cabi.raw(dfeat), 0, _cabi.raw(dws),
            _cabi.raw(g1), _cabi.raw(g2m), _cabi.raw(g2c), B, N, K, d, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_pool_bwd")
    return dfeat, dws, g1, g2m, g2c


def lfa_moments(mode: int, xyz, idx32, d: int, w_rpe1=None, a_rpe1=None, b_rpe1=None, gsym=None, gsum=None):
    """BatchNorm moments of the fused LocSE (C ABI ``r3d_lfa_moments``).
    mode 0 -> m_rpe (16,16) float64; mode 1 -> (m_r1 (h,h), s_r1 (h,16)) float64; mode 2 -> g1 (h,16) float64."""
    xyz, xs = _cloud_view(xyz)
    B, N, K = idx32.shape
    h = d // 2
    dev = xyz.device
    m_rpe = m_r1 = s_r1 = g1 = None
    if mode == 0:
        m_rpe = zeros((16, 16), torch.float64, dev)
    elif mode == 1:
        buf = zeros(h * h + h * 16, torch.float64, dev)
        m_r1, s_r1 = buf[:h * h].view(h, h), buf[h * h:].view(h, 16)
    else:
        g1 = zeros((h, 16), torch.float64, dev)
    flops = float(B) * N * K * 2 * (256 if mode == 0 else 10 * h + h * h + (16 * h if mode == 2 else 0))
    nbytes = float(B) * N * (12 + 4 * K)
    if mode == 1 and lfa_mom_tc_supported(d, K, B * N):
        _lfa_tc_bwd(4, "lfa_cl_mom", flops, nbytes, xyz, xs, idx32, None, 0, w_rpe1, a_rpe1, b_rpe1, m_r1=m_r1, s_r1=s_r1)
        return m_r1, s_r1
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_moments{mode}[N={N},d={d}]", flops=flops, bytes=nbytes):
        rc = _cabi.lib().r3d_lfa_moments(mode, _cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.ptr(w_rpe1),
                                         _cabi.ptr(a_rpe1), _cabi.ptr(b_rpe1), _cabi.raw(m_rpe), _cabi.raw(m_r1),
                                         _cabi.raw(s_r1), _cabi.ptr(gsym), _cabi.ptr(gsum), _cabi.raw(g1), B, N, K, d,
                                         _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_moments")
    if mode == 0:
        return m_rpe
    if mode == 1:
        return m_r1, s_r1
    return g1


def bn_from_moments(w: torch.Tensor, s: torch.Tensor, m: torch.Tensor, count: float, bn, bias):
    """BatchNorm affine (a, c) of y = W x from the input moments (C ABI ``r3d_bn_from_moments``): w (cout,cin),
    s (>=cin) fp64 sums, m (ld,ld) fp64 second-moment sums (ld >= cin), ``count`` rows.  Updates bn's running
    statistics.  Returns (a, c, save)."""
    cout, cin = w.shape
    dev = w.device
    a = torch.empty(cout, dtype=torch.float32, device=dev)
    c = torch.empty(cout, dtype=torch.float32, device=dev)
    save = torch.empty((5, cout), dtype=torch.float64, device=dev)
    track = bn.track_running_stats and bn.running_mean is not None
    with torch.cuda.device(dev), _cabi.kernel_timer(f"bn_from_moments[{cout}x{cin}]", flops=2.0 * cout * cin * cin, bytes=8.0 * cin * cin):
        rc = _cabi.lib().r3d_bn_from_moments(
            _cabi.ptr(w), cout, cin, _cabi.raw(s), s.stride(0), _cabi.raw(m), m.stride(0), float(count),
            _cabi.ptr(bn.weight.detach()), _cabi.ptr(bn.bias.detach()),
            _cabi.ptr(bias.detach()) if bias is not None else None, float(bn.eps), float(bn.momentum),
            _cabi.ptr(bn.running_mean) if track else None, _cabi.ptr(bn.running_var) if track else None,
            _cabi.ptr(bn.num_batches_tracked) if track else None, _cabi.ptr(a), _cabi.ptr(c), _cabi.ptr(save),
            _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_bn_from_moments")
    return a, c, save


def bn_from_moments_bwd(w, s, m, count: float, gamma, save, ga, gc, need_moments: bool):
    """Backward of ``bn_from_moments``: ga, gc fp64 -> (dW fp64, dgamma, dbeta, dM | None, dS | None)."""
    cout, cin = w.shape
    dev = w.device
    dw = torch.empty((cout, cin), dtype=torch.float64, device=dev)
    dgb = torch.empty((2, cout), dtype=torch.float32, device=dev)
    scal = torch.empty((2, cout), dtype=torch.float64, device=dev)
    dm = torch.empty((cin, cin), dtype=torch.float64, device=dev) if need_moments else None
    ds = torch.empty(cin, dtype=torch.float64, device=dev) if need_moments else None
    with torch.cuda.device(dev), _cabi.kernel_timer(f"bn_from_moments_bwd[{cout}x{cin}]", flops=4.0 * cout * cin * cin,
                                                    bytes=8.0 * cin * cin):
        rc = _cabi.lib().r3d_bn_from_moments_bwd(
            _cabi.ptr(w), cout, cin, _cabi.raw(s), s.stride(0), _cabi.raw(m), m.stride(0), float(count),
            _cabi.ptr(gamma), _cabi.ptr(save), _cabi.ptr(ga.double().contiguous()), _cabi.ptr(gc.double().contiguous()),
            _cabi.ptr(dw),
            _cabi.raw(dgb[0]), _cabi.raw(dgb[1]), _cabi.ptr(scal), _cabi.ptr(dm), _cabi.ptr(ds), _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_bn_from_moments_bwd")
    return dw, dgb[0], dgb[1], dm, ds


def lfa_pool2_bwd_train(xyz, idx32, feat, w_rpe1, a_rpe1, b_rpe1, w_rpe2T, a_rpe2, b_rpe2, w_scoreT, w_score, dpooled,
                        w_rpe2=None, cache: Optional[dict] = None):
    """Pass 1 of the train-mode stage-2 backward (C ABI ``r3d_lfa_pool2_bwd_train``, or mode 2 of ``r3d_lfa_tc_bwd``
    on the tensor cores, which wants ``w_rpe2`` in its stored [out][in] layout).
    Returns (dfeat (B,N,h), dw_score (d,d), du2_tiles, sum_du2 (2,h) fp64); du2_tiles is opaque (its layout belongs to
    the kernel family that wrote it) and goes to ``lfa_bn2_bwd`` unchanged."""
    xyz, xs = _cloud_view(xyz)
    feat, fs = _rows_view(feat.detach())
    B, N, K = idx32.shape
    h = feat.shape[2]
    d = 2 * h
    dev = xyz.device
    L = _cabi.lib()
    if d == 256 and lfa_bwd_tc_supported(d, K, B * N):
        if w_rpe2 is None:
            w_rpe2 = w_rpe2T.t().contiguous()
        rmat = cache.pop("r2", None) if cache is not None else None
        if rmat is None:
            rmat = lfa_r2_rows(xyz, xs, idx32, w_rpe1, a_rpe1, b_rpe1, w_rpe2, a_rpe2, b_rpe2)
        dpooled = dpooled.contiguous()
        n_df = B * N * h
        acc = zeros(n_df + d * d + 2, torch.float32, dev)
        dfeat, dws, scal = acc[:n_df].view(B, N, h), acc[n_df:n_df + d * d].view(d, d), acc[n_df + d * d:]
        rows = B * N * K
        part = torch.empty((2, rows, h), dtype=torch.float32, device=dev)
        sums = zeros((2, h), torch.float64, dev)
        _absmax_into(dpooled, scal)
        _lfa_tc_wide(2, "lfa_cl_bwd2", float(B) * N * (2 * K * (10 * h + 3 * d * d + d + h * h)),
                     float(B) * N * (12 + 4 * K + 4 * h + 4 * d + 8 * K * h), xyz, xs, idx32, feat, fs, w_rpe1, a_rpe1,
                     b_rpe1, w_score, rmat=rmat, dpooled=dpooled, dfeat=dfeat, dws=dws, du2_part=part, sums=sums, scal=scal)
        pts = L.r3d_lfa_tile_points_for(K, d, B, N)
        du2 = torch.empty(B * (-(-N // pts)) * h * pts * K, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_du2_combine[N={N},d={d}]", flops=float(rows) * h,
                                                        bytes=12.0 * rows * h):
            rc = L.r3d_lfa_du2_combine(_cabi.ptr(part), _cabi.ptr(du2), B, N, K, h, pts, _cabi.stream_ptr(dev))
        _cabi.check(rc, "r3d_lfa_du2_combine")
        return dfeat, dws, du2, sums
    if lfa_bwd_tc_supported(d, K, B * N):
        if w_rpe2 is None:
            w_rpe2 = w_rpe2T.t().contiguous()
        dpooled = dpooled.contiguous()
        n_df = B * N * h
        acc = zeros(n_df + d * d + 2, torch.float32, dev)
        dfeat, dws, scal = acc[:n_df].view(B, N, h), acc[n_df:n_df + d * d].view(d, d), acc[n_df + d * d:]
        du2 = torch.empty(L.r3d_lfa_tc_du2_floats(B, N, K, d), dtype=torch.float32, device=dev)
        sums = zeros((2, h), torch.float64, dev)
        _absmax_into(dpooled, scal)
        _lfa_tc_bwd(2, "lfa_cl_bwd2", float(B) * N * (2 * K * (10 * h + 3 * d * d + d + h * h)),
                    float(B) * N * (12 + 4 * K + 4 * h + 4 * d + 8 * K * h), xyz, xs, idx32, feat, fs, w_rpe1, a_rpe1,
                    b_rpe1, w_rpe2=w_rpe2, a_rpe2=a_rpe2, b_rpe2=b_rpe2, w_score=w_score, dpooled=dpooled, dfeat=dfeat,
                    dws=dws, du2=du2, sums=sums, scal=scal)
        return dfeat, dws, (du2, scal), sums
    pts = L.r3d_lfa_tile_points_for(K, d, B, N)
    if pts <= 0:
        raise ValueError(f"r3d_lfa_tile_points_for: unsupported shape d={d}, K={K}")
    tiles = -(-N // pts)
    dpooled = dpooled.contiguous()
    n_df = B * N * h
    acc = zeros(n_df + d * d, torch.float32, dev)
    dfeat, dws = acc[:n_df].view(B, N, h), acc[n_df:].view(d, d)
    du2 = torch.empty(B * tiles * h * pts * K, dtype=torch.float32, device=dev)
    sums = zeros((2, h), torch.float64, dev)
    flops = float(B) * N * (2 * K * (10 * h + 3 * d * d + d + h * h))
    nbytes = float(B) * N * (12 + 4 * K + 4 * h + 4 * d + 8 * K * h)
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_pool2_bwd[N={N},d={d}]", flops=flops, bytes=nbytes):
        rc = L.r3d_lfa_pool2_bwd_train(_cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.raw(feat), fs, _cabi.ptr(w_rpe1),
                                       _cabi.ptr(a_rpe1), _cabi.ptr(b_rpe1), _cabi.ptr(w_rpe2T), _cabi.ptr(a_rpe2),
                                       _cabi.ptr(b_rpe2), _cabi.ptr(w_scoreT), _cabi.ptr(w_score), _cabi.ptr(dpooled),
                                       _cabi.raw(dfeat), 0, _cabi.raw(dws), _cabi.ptr(du2), _cabi.ptr(sums), B, N, K, d,
                                       _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_pool2_bwd_train")
    return dfeat, dws, du2, sums


def lfa_bn2_bwd(xyz, idx32, w_rpe1, a_rpe1, b_rpe1, du2_tiles, w_rpe2T, w_rpe2, bn2, h: int, g1_acc=None):
    """Pass 2 (C ABI ``r3d_lfa_bn2_bwd``): -> (g1 (h,16) fp64, dw2 (h,h) fp64).  ``g1_acc``: accumulate g1 into this
    (h,16) fp64 buffer instead of a fresh one."""
    xyz, xs = _cloud_view(xyz)
    B, N, K = idx32.shape
    d = 2 * h
    dev = xyz.device
    buf = zeros(h * 16 + h * h, torch.float64, dev)
    g1, dw2 = (buf[:h * 16].view(h, 16) if g1_acc is None else g1_acc), buf[h * 16:].view(h, h)
    flops = float(B) * N * K * 2 * (10 * h + 3 * h * h + 16 * h)
    nbytes = float(B) * N * (12 + 4 * K + 4 * K * h)
    if isinstance(du2_tiles, tuple):                 # written by the tensor-core pass 1: (tiles, operand-scale scalars)
        du2, scal = du2_tiles
        _lfa_tc_bwd(3, "lfa_cl_bn2", flops, nbytes, xyz, xs, idx32, None, 0, w_rpe1, a_rpe1, b_rpe1, w_rpe2=w_rpe2,
                    g1=g1, du2=du2, bn2=bn2, dw2=dw2, scal=scal)
        return g1, dw2
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_bn2_bwd[N={N},d={d}]", flops=flops, bytes=nbytes):
        rc = _cabi.lib().r3d_lfa_bn2_bwd(_cabi.raw(xyz), xs, _cabi.ptr(idx32), _cabi.ptr(w_rpe1), _cabi.ptr(a_rpe1),
                                         _cabi.ptr(b_rpe1), _cabi.ptr(du2_tiles), _cabi.ptr(w_rpe2T), _cabi.ptr(w_rpe2),
                                         _cabi.ptr(bn2), _cabi.raw(g1), _cabi.raw(dw2), B, N, K, d,
                                         _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_bn2_bwd")
    return g1, dw2


def lfa_rpe1_grads(w, s, m, count: float, gamma, save, g1):
    """mlp_rpe1's parameter gradients from the block's g1 accumulator (C ABI ``r3d_lfa_rpe1_grads``):
    w (h,cin) fp32, s / m the moments given to ``bn_from_moments``, save its scratch, g1 (h,16) fp64.
    Returns (dW (h,cin), dgamma (h), dbeta (h)) fp32."""
    cout, cin = w.shape
    dev = w.device
    dw = torch.empty((cout, cin), dtype=torch.float32, device=dev)
    dgb = torch.empty((2, cout), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_rpe1_grads[{cout}x{cin}]", flops=4.0 * cout * cin * cin,
                                                    bytes=8.0 * cin * cin):
        rc = _cabi.lib().r3d_lfa_rpe1_grads(_cabi.ptr(w), cout, cin, _cabi.raw(s), s.stride(0), _cabi.raw(m), m.stride(0),
                                            float(count), _cabi.ptr(gamma), _cabi.ptr(save), _cabi.ptr(g1), g1.stride(0),
                                            _cabi.ptr(dw), _cabi.raw(dgb[0]), _cabi.raw(dgb[1]), _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_rpe1_grads")
    return dw, dgb[0], dgb[1]


def lfa_bn2_coeffs(sums, a2, c2, save, rows: float):
    """Coefficients of ``lfa_bn2_bwd`` and mlp_rpe2's (dgamma, dbeta) from pass 1's sums (C ABI ``r3d_lfa_bn2_coeffs``).
    Returns (bn2 (5,h) fp32, dgamma (h), dbeta (h))."""
    h = a2.shape[0]
    dev = a2.device
    bn2 = torch.empty((5, h), dtype=torch.float32, device=dev)
    dgb = torch.empty((2, h), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _cabi.kernel_timer(f"lfa_bn2_coeffs[{h}]", flops=16.0 * h, bytes=64.0 * h):
        rc = _cabi.lib().r3d_lfa_bn2_coeffs(_cabi.ptr(sums), _cabi.ptr(a2), _cabi.ptr(c2), _cabi.ptr(save), float(rows), h,
                                            _cabi.ptr(bn2), _cabi.raw(dgb[0]), _cabi.raw(dgb[1]), _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_lfa_bn2_coeffs")
    return bn2, dgb[0], dgb[1]
