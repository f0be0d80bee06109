"""GPU parity of the tensor-core per-point layer kernels (C ABI r3d_pc_gemm / r3d_pc_wgrad, csrc/pw_cl.cu) against an
fp64 evaluation of the same products, next to the FP32 CUDA-core kernels they replace (r3d_pointwise, r3d_rowreduce_gemm):
the split-fp16 tensor-core result must be as close to the fp64 value as the FP32 kernel's (tolerance 2e-6 of the output's
largest entry, or 3x the FP32 kernel's own error)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    return importlib.import_module("3d_recognizer_b200.ops")


def _err(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M,cin,cout", [(655360, 64, 128), (163840, 128, 256), (40960, 128, 512), (70001, 32, 64),
                                        (130, 48, 40), (64, 16, 8), (20000, 128, 32), (9999, 96, 200),
                                        (40960, 1024, 256), (40960, 256, 512), (163840, 512, 128), (5000, 272, 72)])
@pytest.mark.parametrize("transposed", [False, True])
def test_pc_gemm_vs_fp64(ops, M, cin, cout, transposed):
    g = torch.Generator(device="cuda").manual_seed(M + cin)
    # rows of very different magnitude (a row scale each) and a few zero rows
    x = torch.randn(M, cin, device="cuda", generator=g) * torch.exp(3 * torch.randn(M, 1, device="cuda", generator=g))
    x[::97] = 0.0
    w = torch.randn((cin, cout) if transposed else (cout, cin), device="cuda", generator=g) * 0.1
    ref = x.double() @ (w.double() if transposed else w.double().t())
    amax = torch.zeros(1, device="cuda")
    stats = torch.zeros(2 * cout, dtype=torch.float64, device="cuda")
    got = ops.pc_gemm(x, w, transposed=transposed, stats=stats, absmax_out=amax)
    plain = ops.pointwise(x.unsqueeze(0), w if transposed else w.t().contiguous()).squeeze(0)
    e_tc, e_fp32 = _err(got, ref), _err(plain, ref)
    assert e_tc < max(2e-6, 3 * e_fp32), (e_tc, e_fp32)
    # row-wise too: every row is held to ITS OWN magnitude (the per-row operand scale)
    row_scale = ref.abs().amax(dim=1).clamp_min(1e-30)
    assert float(((got.double() - ref).abs().amax(dim=1) / row_scale).max()) < 1e-5
    assert float(amax) == float(x.abs().max())
    assert _err(stats[:cout], ref.sum(dim=0)) < 1e-6 or float(ref.sum(dim=0).abs().max()) < 1e-3 * float(ref.abs().sum(dim=0).max())
    assert _err(stats[cout:], (ref * ref).sum(dim=0)) < 3e-6        # fp32 partial sums of 32 rows, then fp64


def test_pc_gemm_epilogue_and_views(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    big = torch.randn(5000, 96, device="cuda", generator=g)
    x = big[:, 16:80]                                         # row-strided view: ldx = 96, 16-byte aligned
    w = torch.randn(72, 64, device="cuda", generator=g) * 0.2
    scale, shift = torch.rand(72, device="cuda", generator=g) + 0.5, torch.randn(72, device="cuda", generator=g)
    for act, fn in (("relu", torch.relu), ("lrelu", lambda t: torch.nn.functional.leaky_relu(t, 0.2)), (None, lambda t: t)):
        got = ops.pc_gemm(x, w, scale=scale, shift=shift, act=act, slope=0.2)
        ref = fn((x.double() @ w.double().t()) * scale.double() + shift.double())
        assert _err(got, ref) < 2e-6


@pytest.mark.parametrize("M,ca,cb", [(40960, 256, 1024), (163840, 256, 128), (655360, 32, 256), (655360, 128, 64),
                                     (163840, 64, 128), (2621440, 32, 64), (1000, 8, 8), (4097, 40, 136), (63, 16, 24),
                                     (300000, 512, 8)])
def test_pc_wgrad_vs_fp64(ops, M, ca, cb):
    g = torch.Generator(device="cuda").manual_seed(M + ca)
    a = torch.randn(M, ca, device="cuda", generator=g) * 1e-4 * torch.exp(torch.randn(1, ca, device="cuda", generator=g))
    b = torch.randn(M, cb, device="cuda", generator=g) + 0.3            # activations: not centred
    ref = torch.zeros(ca, cb, dtype=torch.float64, device="cuda")
    for s in range(0, M, 262144):                                       # fp64 reference in slices (memory)
        ref += a[s:s + 262144].double().t() @ b[s:s + 262144].double()
    got = ops.pc_wgrad(a, b)
    plain = ops.rowreduce_gemm(a, b)
    e_tc, e_fp32 = _err(got, ref), _err(plain, ref)
    assert e_tc < max(2e-6, 3 * e_fp32), (e_tc, e_fp32)
    # loose bounds (what a producer-side running maximum may hand over) cost nothing
    loose = ops.pc_wgrad(a, b, absmax_a=a.abs().max().reshape(1) * 300.0, absmax_b=b.abs().max().reshape(1) * 1000.0)
    assert _err(loose, ref) < max(2e-6, 3 * e_fp32)


def test_pc_wgrad_same_sign_sums_are_not_biased(ops):
    """All-positive operands, the worst case for the accumulator's truncation on every MMA (each add loses up to one ulp
    of the running sum, always downwards: 2.5e-5 over the ~800 MMAs of a CTA's row slice if left alone).  With the
    two-level fold a first-level sum sees 24 MMAs: the deficit stays below 1e-6."""
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.rand(1 << 20, 64, device="cuda", generator=g) + 0.5
    b = torch.rand(1 << 20, 64, device="cuda", generator=g) + 0.5
    ref = a.double().t() @ b.double()
    got = ops.pc_wgrad(a, b)
    rel = (got.double() - ref) / ref
    assert float(rel.abs().max()) < 2e-6 and abs(float(rel.mean())) < 1e-6
