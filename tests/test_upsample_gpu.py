"""GPU parity of the up-sampling kernels (C ABI r3d_upsample / r3d_upsample_bwd) against the oracle's UpSampler
restatement (oracle.network.upsample, modules.py:328-456) in the forward direction and against fp64 autograd of the
same formula in the backward direction, with and without the decoder's skip concat (modules.py:600-602)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import network as onet
from test_forward_gpu import rel_err

pytestmark = pytest.mark.gpu


def _case(B, N1, N2, F, Fs, seed):
    rng = np.random.RandomState(seed)
    feat = torch.from_numpy(rng.randn(B, N1, F).astype(np.float32))
    skip = torch.from_numpy(rng.randn(B, N2, Fs).astype(np.float32)) if Fs else None
    xyz = torch.from_numpy(rng.rand(B, N1, 3).astype(np.float32))
    xyz_up = torch.from_numpy(rng.rand(B, N2, 3).astype(np.float32))
    return feat, skip, xyz, xyz_up


@pytest.mark.parametrize("approach", ["nni", "nna", "idw", "isdw"])
@pytest.mark.parametrize("B,N1,N2,F,Fs", [(2, 700, 3000, 8, 0), (3, 160, 640, 512, 256), (1, 97, 1000, 5, 3),
                                          (2, 64, 257, 2, 0)])
def test_upsample_kernels_vs_oracle_and_fp64_autograd(approach, B, N1, N2, F, Fs):
    engine = importlib.import_module("3d_recognizer_b200.engine")
    ops = importlib.import_module("3d_recognizer_b200.ops")
    feat, skip, xyz, xyz_up = _case(B, N1, N2, F, Fs, F + N1)
    fg = feat.cuda().requires_grad_(True)
    sg = skip.cuda().requires_grad_(True) if Fs else None
    got = engine.upsample(approach, fg, xyz.cuda(), xyz_up.cuda(), skip=sg)
    assert got.shape == (B, N2, F + Fs)

    # forward: the oracle's UpSampler on (B,F,N1,1) features, then the plain concat
    ref = onet.upsample(approach, feat.transpose(1, 2).unsqueeze(-1), xyz, xyz_up).squeeze(-1).transpose(1, 2)
    ref = ref if skip is None else torch.cat((ref, skip), dim=-1)
    if approach == "nni":
        assert torch.equal(got.detach().cpu(), ref)                  # a gather: bit-exact
    else:
        assert rel_err(got.detach().cpu(), ref) < 2e-6

    # backward: fp64 autograd of the reference formula on the kernel's own neighbours (exact search, same as the oracle's)
    gout = torch.randn(B, N2, F + Fs, generator=torch.Generator().manual_seed(1))
    (got * gout.cuda()).sum().backward()
    k = 1 if approach == "nni" else 8
    nn_ = ops.knn(xyz.cuda(), xyz_up.cuda(), k, idx64=True, dist=True)
    idx, dist = nn_["idx64"].cpu(), nn_["dist"].cpu().double()
    fd = feat.double().requires_grad_(True)
    sd = skip.double().requires_grad_(True) if Fs else None
    nf = torch.stack([fd[b][idx[b]] for b in range(B)])              # (B,N2,K,F)
    if approach == "nni":
        up = nf[:, :, 0]
    else:
        w = (1.0 + 1e-7) / (dist ** ops.UP_WEIGHTING[approach][1] + 1e-7)
        w = w / w.sum(dim=-1, keepdim=True)
        up = (w.unsqueeze(-1) * nf).sum(dim=2)
    out = up if sd is None else torch.cat((up, sd), dim=-1)
    (out * gout.double()).sum().backward()
    assert rel_err(fg.grad.cpu(), fd.grad) < 2e-6
    if Fs:
        assert torch.equal(sg.grad.cpu(), gout[:, :, F:])


@pytest.mark.parametrize("idx_dtype", [torch.int32, torch.int64])
def test_upsample_channel_major_and_index_types(idx_dtype):
    """(B,F,N2) output (what Model.upsample returns) equals the transposed point-major result, for both index types,
    and mean weighting (the reference's inverse_distance_weighting=False branch, modules.py:408-412)."""
    ops = importlib.import_module("3d_recognizer_b200.ops")
    feat, _, xyz, xyz_up = _case(2, 300, 2000, 3, 0, 7)
    nn_ = ops.knn(xyz.cuda(), xyz_up.cuda(), 8, idx64=True, idx32=True, dist=True)
    idx = nn_["idx64"] if idx_dtype == torch.int64 else nn_["idx32"]
    for approach in ("nni", "idw", "isdw", "mean"):
        rows = ops.upsample(approach, feat.cuda(), idx, nn_["dist"])
        cm = ops.upsample(approach, feat.cuda(), idx, nn_["dist"], channel_major=True)
        assert cm.shape == (2, 3, 2000) and torch.equal(cm.transpose(1, 2), rows)
    mean = ops.upsample("mean", feat.cuda(), idx, None).cpu()
    ref = torch.stack([feat[b][nn_["idx64"][b].cpu()] for b in range(2)]).mean(dim=2)
    assert rel_err(mean, ref) < 1e-6


def test_upsample_argument_errors():
    ops = importlib.import_module("3d_recognizer_b200.ops")
    feat = torch.zeros(1, 10, 4, device="cuda")
    idx = torch.zeros(1, 20, 8, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        ops.upsample("cubic", feat, idx)
    with pytest.raises(ValueError):
        ops.upsample("idw", feat, idx, None)                          # inverse-distance weights need distances
    with pytest.raises(ValueError):
        ops.upsample("nni", feat, idx.float())
    with pytest.raises(ValueError):                                   # K above the compiled maximum (R3D_EKMAX)
        ops.upsample("mean", feat, torch.zeros(1, 20, 17, dtype=torch.int32, device="cuda"))
    assert ops.upsample("nni", feat, idx[:, :0]).shape == (1, 0, 4)   # empty query set
