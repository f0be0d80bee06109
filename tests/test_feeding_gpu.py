"""GPU parity of the data-feeding kernels (C ABI r3d_feed_batch / r3d_sample_subset, 3d_recognizer_b200/dataset.py)
against the fixtures written from the reference's PointCloudPreprocessor (numpy stream: equal to round-off) and, for the
device random streams, through the properties the augmentation defines (subset uniformity, jitter distribution and
clipping, exact similarity transform)."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import feeding

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ds():
    return importlib.import_module("3d_recognizer_b200.dataset")


@pytest.mark.parametrize("consistent", [True, False])
@pytest.mark.parametrize("name", list(feeding.FEED_CASES))
def test_loader_vs_reference_golden(ds, name, consistent):
    """numpy stream, batches of 2 (then 1): every item of the epoch equals what the reference's DataLoader yields, and
    what exact (fp64) arithmetic on the same inputs and random numbers gives.  Tolerances relative to the largest
    coordinate: 2e-6 against the fp64 evaluation, and against the reference where it computes in float64 (clouds 0, 1);
    5e-5 against the reference on the float32 cloud, whose centre it accumulates sequentially in float32
    (oracle.feeding.FEED_TOL)."""
    g = np.load(os.path.join(GOLDEN, "feed_golden.npz"))
    norm, aug = feeding.FEED_CASES[name]
    data = feeding.feed_dataset()
    settings = ds.AugmentationSettings() if aug else None
    loader = ds.get_data_loader(data, feeding.FEED_N, 2, shuffle=False, consistent_sampling=consistent,
                                augmentation_settings=settings, normalization=norm, device="cuda", rng="numpy")
    assert len(loader) == 2
    np.random.seed(feeding.FEED_SEED)
    exact = [feeding.preprocess(d[0].astype(np.float64), d[1], d[2], feeding.FEED_N, consistent, settings, norm)[0]
             for d in data]
    np.random.seed(feeding.FEED_SEED)
    seen = 0
    for inp, labels, idx in loader:
        assert inp.is_cuda and inp.dtype == torch.float32 and labels.dtype == torch.int64 and idx.dtype == torch.int64
        for j, i in enumerate(idx.tolist()):
            ref = g[f"{name}/{int(consistent)}/{i}/input"]
            assert np.array_equal(labels[j].cpu().numpy(), g[f"{name}/{int(consistent)}/{i}/labels"])
            got = inp[j].cpu().numpy()
            assert np.array_equal(got[:, 3:], ref[:, 3:])                      # features: a gather
            scale = np.abs(ref[:, :3]).max()
            assert np.abs(got[:, :3] - exact[i]).max() < 2e-6 * scale, (name, i)
            assert np.abs(got - ref).max() < feeding.FEED_TOL[i] * scale, (name, i)
            if norm is None and not aug:
                assert np.array_equal(got, ref)
            seen += 1
    assert seen == 3


def test_perturbate_point_cloud_vs_oracle(ds):
    xyz = (np.random.RandomState(3).rand(4000, 3) * [1.0, 2.0, 0.3]).astype(np.float32)
    s = ds.AugmentationSettings()
    np.random.seed(21)
    ref = feeding.perturbate(xyz, s)
    np.random.seed(21)
    got = ds.perturbate_point_cloud(xyz, s)
    assert np.abs(got - ref).max() < 2e-6 * np.abs(ref).max()


def test_sample_subset_properties(ds):
    sizes = torch.tensor([5000, 300, 1000, 1001], dtype=torch.int32, device="cuda")
    n = 1000
    a = ds.sample_subset(sizes, n, seed=7, counter=0)
    assert torch.equal(a, ds.sample_subset(sizes, n, seed=7, counter=0))          # a function of (seed, counter)
    assert not torch.equal(a[0], ds.sample_subset(sizes, n, seed=7, counter=1)[0])
    assert not torch.equal(a[0], ds.sample_subset(sizes, n, seed=8, counter=0)[0])
    a = a.cpu().numpy()
    for row, N in ((0, 5000), (3, 1001)):                                        # N > n: n distinct points, ascending
        assert np.all(np.diff(a[row]) > 0) and a[row][0] >= 0 and a[row][-1] < N
    assert np.array_equal(a[2], np.arange(1000))                                 # N == n: every point once
    assert np.array_equal(a[1][:300], np.arange(300))                            # N < n: all points, then duplicates
    assert a[1][300:].min() >= 0 and a[1][300:].max() < 300 and len(np.unique(a[1][300:])) > 200
    # uniformity: inclusion frequency of every point over 400 draws of 250 out of 1000 (p = 1/4, sd 0.0217)
    sizes = torch.full((400,), 1000, dtype=torch.int32, device="cuda")
    draws = ds.sample_subset(sizes, 250, seed=3, counter=5).cpu().numpy()
    freq = np.bincount(draws.reshape(-1), minlength=1000) / 400.0
    assert abs(freq.mean() - 0.25) < 1e-9 and np.abs(freq - 0.25).max() < 0.1 and 0.018 < freq.std() < 0.026
    # adjacent points are chosen independently
    inc = np.zeros((400, 1000), dtype=bool)
    inc[np.arange(400)[:, None], draws] = True
    assert abs(np.mean(inc[:, :-1] & inc[:, 1:]) - 0.0625) < 0.004


def test_device_jitter_distribution_and_similarity_transform(ds):
    from scipy import stats
    rng = np.random.RandomState(1)
    N = 40960
    xyz = (rng.rand(N, 3) * [2.0, 1.0, 0.5]).astype(np.float32)
    cache = ds.CloudCache([(xyz, np.zeros((N, 0), np.float32), np.zeros(N, np.int64))], "cuda")
    idx = torch.arange(N, dtype=torch.int32, device="cuda").unsqueeze(0)
    radius = feeding.mean_radius(xyz.astype(np.float64))
    # (a) jitter only (scale 1, no rotation, no shift), device normals
    s = ds.AugmentationSettings(jitter_variance=0.03, jitter_limit=0.05)
    ident = np.array([[1.0, 0, 0, 0, 0, 0, 0]])
    out, _ = ds.feed_batch(cache, [0], idx, None, ident, None, s, seed=11, counter=3)
    delta = (out[0].cpu().numpy().astype(np.float64) - xyz).reshape(-1)
    sigma = radius * 0.03
    assert np.abs(delta).max() <= 0.05 + 1e-6 and (np.abs(delta) > 0.0499).mean() > 1e-3      # clipped tails exist
    inner = delta[np.abs(delta) < 0.049] / sigma
    assert abs(delta.mean()) < 4 * sigma / np.sqrt(delta.size)
    # Kolmogorov-Smirnov against the truncated normal the clip leaves
    lim = 0.049 / sigma
    assert stats.kstest(inner, stats.truncnorm(-lim, lim).cdf).pvalue > 1e-3
    # the three coordinates of a point and neighbouring points are uncorrelated
    d3 = delta.reshape(N, 3)
    assert np.abs(np.corrcoef(d3.T) - np.eye(3)).max() < 0.02
    assert abs(np.corrcoef(d3[:-1, 0], d3[1:, 0])[0, 1]) < 0.02
    out2, _ = ds.feed_batch(cache, [0], idx, None, ident, None, s, seed=11, counter=4)
    assert not torch.equal(out, out2)                                                          # a fresh draw per batch
    assert torch.equal(out, ds.feed_batch(cache, [0], idx, None, ident, None, s, seed=11, counter=3)[0])
    # (b) no jitter: an exact similarity transform about the centre, shifted by s r u
    s0 = ds.AugmentationSettings(jitter_variance=0.0)
    params = np.array([[1.15, 0.1, -0.17, 0.05, 0.08, -0.1, 0.02]])
    out, _ = ds.feed_batch(cache, [0], idx, None, params, None, s0)
    got = out[0].cpu().numpy().astype(np.float64)
    c = xyz.astype(np.float64).mean(axis=0)
    want = (xyz - c) * 1.15 @ feeding.rotation(params[0, 1:4]).T + c + 1.15 * radius * params[0, 4:]
    assert np.abs(got - want).max() < 1e-6
    pick = rng.randint(0, N, (200, 2))
    d_in = np.linalg.norm(xyz[pick[:, 0]].astype(np.float64) - xyz[pick[:, 1]], axis=1)
    d_out = np.linalg.norm(got[pick[:, 0]] - got[pick[:, 1]], axis=1)
    assert np.abs(d_out - 1.15 * d_in).max() < 1e-5


def test_device_stream_loader_at_training_size(ds):
    """rng='device': 8 cached clouds of 100 000 points -> batches of 4 x 40960 with features and augmentation; every
    output row is a row of its cloud (features and labels carried along), coordinates moved by at most what the settings
    allow."""
    rng = np.random.RandomState(2)
    data = []
    for c in range(8):
        xyz = rng.rand(100000, 3).astype(np.float32)
        feat = np.stack((np.arange(100000, dtype=np.float32), np.full(100000, c, np.float32)), axis=1)
        data.append((xyz, feat, (np.arange(100000) % 3).astype(np.int64)))
    torch.manual_seed(0)
    np.random.seed(0)
    loader = ds.get_data_loader(data, 40960, 4, shuffle=True, consistent_sampling=False,
                                augmentation_settings=ds.AugmentationSettings(), device="cuda", rng="device", seed=5)
    items = []
    for inp, labels, idx in loader:
        assert inp.shape == (4, 40960, 5) and labels.shape == (4, 40960)
        rows = inp[:, :, 3].long()
        assert torch.equal(inp[:, :, 4], idx.cuda().float().view(4, 1).expand(4, 40960))      # rows of the right cloud
        assert torch.equal(labels, rows % 3)
        assert bool((rows[:, 1:] > rows[:, :-1]).all())                                       # distinct, ascending
        for j, i in enumerate(idx.tolist()):
            src = torch.from_numpy(data[i][0]).cuda()[rows[j]]
            moved = (inp[j, :, :3] - src).norm(dim=1)
            # scale 1.2 about the centre (extent 0.87) + rotation 0.18 rad x 3 + shift 0.1 r sqrt(3) + jitter
            assert float(moved.max()) < 0.2 * 0.87 + 3 * 0.18 * 0.87 + 0.1 * 0.6 * 1.8 + 0.1
            assert float(moved.mean()) > 1e-3
        items += idx.tolist()
    assert sorted(items) == list(range(8))


def test_broaden_annotation_vs_oracle(ds):
    rng = np.random.RandomState(6)
    cloud = rng.rand(5000, 3).astype(np.float32) * 0.3
    ann = (rng.rand(5000) < 0.01).astype(np.uint8)
    ref = feeding.broaden_annotation(cloud.astype(np.float64), ann, 0.01)
    got = ds.broaden_annotation(torch.from_numpy(cloud).cuda(), torch.from_numpy(ann).cuda(), 0.01).cpu().numpy()
    assert got.dtype == np.uint8 and ref.sum() > ann.sum()
    # points whose nearest annotated point is within fp32 round-off of the radius may fall on either side
    marked = cloud[ann.astype(bool)].astype(np.float64)
    nearest = np.min(np.linalg.norm(cloud[:, None, :].astype(np.float64) - marked[None], axis=2), axis=1)
    clear = np.abs(nearest - 0.01) > 1e-6
    assert np.array_equal(got[clear], ref[clear]) and clear.mean() > 0.999
    none = ds.broaden_annotation(torch.from_numpy(cloud).cuda(), torch.zeros(5000, dtype=torch.uint8).cuda())
    assert int(none.sum()) == 0
