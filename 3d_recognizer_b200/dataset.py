"""Data feeding for the training step: clouds cached in device memory, batches assembled by one kernel.

Reference: randlanet/utils/dataset.py:11-97 (PointCloudPreprocessor, get_data_loader), randlanet/utils/augmentation.py
(AugmentationSettings, perturbate_point_cloud) and the label helper dataset.py:8-18 (broaden_annotation).  There a
single-process ``DataLoader`` re-reads every item, samples, normalises and augments it with numpy, and stacks the batch
on the host — a few ms per cloud, which an 85 ms step over 64 clouds per GPU (times 8 GPUs) outruns.  Here

  * ``CloudCache`` holds the whole dataset in HBM once (all clouds back to back, fp32 rows [xyz, features] + int64
    labels: 28 bytes per point without features, so hundreds of millions of points fit beside the network);
  * ``DeviceDataLoader`` draws the random numbers and launches ``r3d_feed_batch`` (csrc/feed.cu): gather at the sample
    indices, normalisation, jitter / scale / rotation / shift, batch layout — one launch per batch, no host copy of
    point data.

Random numbers.  ``rng="numpy"`` consumes numpy's global stream in exactly the reference's order (sample indices, the
(N,3) jitter normals, scale, three angles, three shifts — per item), so a seeded epoch reproduces the reference's batches
to fp32 round-off; the host then draws 3 N normals per cloud.  ``rng="device"`` keeps only the 7 per-cloud numbers on
the host stream and draws sample subsets and jitter normals on the device (Philox4x32-10 keyed by seed, batch counter,
cloud, point): statistically the same augmentation, a different stream — results are reproducible for a given seed but
not equal to the reference's.  Batch order: ``torch.utils.data.RandomSampler`` / ``SequentialSampler``, the samplers the
reference's ``DataLoader(shuffle=...)`` instantiates, so the item order follows torch's seed as it does there."""
from dataclasses import dataclass
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi, ops
from .preprocessing import sample_points


@dataclass
class AugmentationSettings:
    """augmentation.py:7-21 (same fields, same defaults)."""
    jitter_variance: float = 0.01
    jitter_limit: float = 0.05
    scale_limit: float = 0.2
    shift_limit: float = 0.1
    rotation_angle_variances: Tuple[float, float, float] = (0.06, 0.06, 0.06)
    rotation_angle_limits: Tuple[float, float, float] = (0.18, 0.18, 0.18)


NORMALIZATIONS = {None: 0, "mean": 1, "max": 2, "stdev": 3}


def _normalization_code(normalization: Optional[str]) -> int:
    # dataset.py:84-92: any other string falls through to radius 1.0 (centring only) — code 4 is handled as "centre"
    return NORMALIZATIONS.get(normalization, 4)


def draw_cloud_parameters(settings: AugmentationSettings) -> np.ndarray:
    """[scale, angle_x, angle_y, angle_z, shift_x, shift_y, shift_z] from numpy's global stream, in the order
    random_scale_point_cloud / random_rotate_point_cloud / random_shift_point_cloud draw them
    (augmentation.py:73, :99-102, :154)."""
    assert len(settings.rotation_angle_variances) == 3, "angle_sigmas should have length 3"
    assert len(settings.rotation_angle_limits) == 3, "angle_clips should have length 3"
    scale = np.random.uniform(1 - settings.scale_limit, 1 + settings.scale_limit)
    angles = [np.clip(sigma * np.random.randn(), -limit, limit)
              for sigma, limit in zip(settings.rotation_angle_variances, settings.rotation_angle_limits)]
    shifts = np.random.uniform(-settings.shift_limit, settings.shift_limit, 3)
    return np.array([scale, *angles, *shifts], dtype=np.float64)


class CloudCache:
    """The dataset resident in device memory: ``points`` (rows, 3+F) fp32, ``labels`` (rows) int64, ``row_start``
    (n_clouds + 1) int64 on the host and the device."""

    def __init__(self, dataset: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]], device: torch.device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("CloudCache keeps the clouds in GPU memory: a CUDA device is required")
        rows, labels, sizes = [], [], []
        for i in range(len(dataset)):
            xyz, features, lab = dataset[i]
            N = xyz.shape[0]
            assert xyz.shape[1] == 3, "Point coordinates should have shape (N, 3)!"
            assert features.shape[0] == N, "Features should have shape (N, F)!"
            assert lab.shape == (N,), "Labels should have shape (N,)!"
            rows.append(np.concatenate((np.asarray(xyz, dtype=np.float32),
                                        np.asarray(features, dtype=np.float32).reshape(N, -1)), axis=1))
            labels.append(np.asarray(lab).astype(np.int64))
            sizes.append(N)
        widths = {r.shape[1] for r in rows}
        assert len(widths) <= 1, "all clouds must carry the same number of features"
        self.device = device
        self.sizes = np.asarray(sizes, dtype=np.int64)
        self.row_start = np.concatenate(([0], np.cumsum(self.sizes))).astype(np.int64)
        self.width = widths.pop() if widths else 3
        self.points = torch.from_numpy(np.concatenate(rows, axis=0) if rows else np.zeros((0, 3), np.float32)).to(device)
        self.labels = torch.from_numpy(np.concatenate(labels) if labels else np.zeros((0,), np.int64)).to(device)

    def __len__(self) -> int:
        return len(self.sizes)


def feed_batch(cache: CloudCache, items: Sequence[int], sample_idx: torch.Tensor, normalization: Optional[str] = None,
               aug: Optional[np.ndarray] = None, noise: Optional[torch.Tensor] = None,
               settings: Optional[AugmentationSettings] = None, seed: int = 0, counter: int = 0
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """One batch from the cache (C ABI ``r3d_feed_batch``): items = cloud numbers, sample_idx (B,n) int32 device rows
    within each cloud, aug (B,7) per-cloud parameters (``draw_cloud_parameters``) or None, noise (B,n,3) standard
    normals or None (device Philox).  Returns (input (B,n,3+F) fp32, labels (B,n) int64) on the device."""
    dev = cache.device
    B, n = sample_idx.shape
    assert sample_idx.dtype == torch.int32 and sample_idx.is_cuda and sample_idx.is_contiguous()
    code = _normalization_code(normalization)
    starts = torch.from_numpy(cache.row_start[np.asarray(items, dtype=np.int64)]).to(dev, non_blocking=True)
    out = torch.empty((B, n, cache.width), dtype=torch.float32, device=dev)
    labels = torch.empty((B, n), dtype=torch.int64, device=dev)
    aug_t = None
    sigma = limit = 0.0
    if aug is not None:
        settings = settings or AugmentationSettings()
        aug_t = torch.from_numpy(np.ascontiguousarray(aug, dtype=np.float32)).to(dev, non_blocking=True)
        assert aug_t.shape == (B, 7)
        sigma, limit = float(settings.jitter_variance), float(settings.jitter_limit)
        if noise is not None:
            noise = noise.to(dev, torch.float32, non_blocking=True).contiguous()
            assert noise.shape == (B, n, 3)
    passes = 1 + (2 if code else 0) + (3 if aug is not None else 0)
    with torch.cuda.device(dev), _cabi.kernel_timer("feed_batch", flops=0.0,
                                                    bytes=4.0 * B * n * (cache.width * (passes + 1) + 5)):
        rc = _cabi.lib().r3d_feed_batch(_cabi.ptr(cache.points), cache.width, _cabi.ptr(cache.labels), _cabi.ptr(starts),
                                        _cabi.ptr(sample_idx), n, code, _cabi.ptr(aug_t), _cabi.ptr(noise), sigma, limit,
                                        seed & (2 ** 64 - 1), counter & (2 ** 64 - 1), _cabi.ptr(out), _cabi.ptr(labels),
                                        B, _cabi.stream_ptr(dev))
    _cabi.check(rc, "r3d_feed_batch")
    return out, labels


def sample_subset(sizes: torch.Tensor, n: int, seed: int, counter: int) -> torch.Tensor:
    """Device-side stand-in for preprocessing.sample_points (C ABI ``r3d_sample_subset``): sizes (B) int32 device ->
    (B,n) int32, a uniform random subset in ascending order, or every point once plus draws with replacement."""
    _cabi.require_cuda(sizes, "sizes")
    assert sizes.dtype == torch.int32 and sizes.is_contiguous()
    B = sizes.shape[0]
    out = torch.empty((B, n), dtype=torch.int32, device=sizes.device)
    with torch.cuda.device(sizes.device):
        rc = _cabi.lib().r3d_sample_subset(_cabi.ptr(sizes), n, seed & (2 ** 64 - 1), counter & (2 ** 64 - 1),
                                           _cabi.ptr(out), B, _cabi.stream_ptr(sizes.device))
    _cabi.check(rc, "r3d_sample_subset")
    return out


class DeviceDataLoader:
    """Iterable of ``(input (B,n,3+F) fp32, labels (B,n) int64, idx (B,) int64)`` like the reference's
    ``DataLoader(PointCloudPreprocessor(...), batch_size, shuffle)`` (dataset.py:100-138), with the tensors on the
    device; the last batch may be smaller (DataLoader's drop_last=False)."""

    def __init__(self, dataset, n_sample_points: int, batch_size: int, shuffle: bool = False,
                 consistent_sampling: bool = True, augmentation_settings: Optional[AugmentationSettings] = None,
                 normalization: Optional[str] = None, device: Optional[torch.device] = None, rng: str = "numpy",
                 seed: int = 0):
        if rng not in ("numpy", "device"):
            raise ValueError(f"rng must be 'numpy' or 'device', got {rng!r}")
        device = torch.device(device if device is not None else "cuda")
        self.cache = dataset if isinstance(dataset, CloudCache) else CloudCache(dataset, device)
        self.n_sample_points, self.batch_size = n_sample_points, batch_size
        self.consistent_sampling, self.augmentation_settings = consistent_sampling, augmentation_settings
        self.normalization, self.rng, self.seed = normalization, rng, seed
        n = len(self.cache)
        self.sampler = (torch.utils.data.RandomSampler(range(n)) if shuffle
                        else torch.utils.data.SequentialSampler(range(n)))
        self._consistent_idx = {}             # consistent sampling: the same indices for a given cloud size, always
        self._sizes32 = torch.from_numpy(self.cache.sizes.astype(np.int32)).to(self.cache.device)
        self._batches = 0

    def __len__(self) -> int:
        return (len(self.cache) + self.batch_size - 1) // self.batch_size

    def _host_indices(self, N: int) -> np.ndarray:
        if not self.consistent_sampling:
            return sample_points(N, self.n_sample_points, consistent=False)
        if N not in self._consistent_idx:
            self._consistent_idx[N] = sample_points(N, self.n_sample_points, consistent=True).astype(np.int32)
        return self._consistent_idx[N]

    def _batch(self, items: List[int]):
        cache, n, s = self.cache, self.n_sample_points, self.augmentation_settings
        dev = cache.device
        B = len(items)
        aug = noise = None
        if self.rng == "numpy":
            # per item, in the reference's order: indices, jitter normals, scale, angles, shifts
            idx = np.empty((B, n), dtype=np.int32)
            if s is not None:
                aug, noise_h = np.empty((B, 7)), np.empty((B, n, 3), dtype=np.float32)
            for j, it in enumerate(items):
                idx[j] = self._host_indices(int(cache.sizes[it]))
                if s is not None:
                    noise_h[j] = np.random.randn(n, 3)
                    aug[j] = draw_cloud_parameters(s)
            sample_idx = torch.from_numpy(idx).to(dev, non_blocking=True)
            if s is not None:
                noise = torch.from_numpy(noise_h)
        else:
            if self.consistent_sampling:
                idx = np.stack([self._host_indices(int(cache.sizes[it])) for it in items])
                sample_idx = torch.from_numpy(idx).to(dev, non_blocking=True)
            else:
                sizes = self._sizes32.index_select(0, torch.as_tensor(items, device=dev))
                sample_idx = sample_subset(sizes, n, self.seed, self._batches)
            if s is not None:
                aug = np.stack([draw_cloud_parameters(s) for _ in items])
        inp, labels = feed_batch(cache, items, sample_idx, self.normalization, aug, noise, s, self.seed, self._batches)
        self._batches += 1
        return inp, labels, torch.as_tensor(items, dtype=torch.int64)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        items: List[int] = []
        for it in self.sampler:
            items.append(int(it))
            if len(items) == self.batch_size:
                yield self._batch(items)
                items = []
        if items:
            yield self._batch(items)


def get_data_loader(dataset, n_sample_points: int, batch_size: int, shuffle: bool = False,
                    consistent_sampling: bool = True, augmentation_settings: Optional[AugmentationSettings] = None,
                    normalization: Optional[str] = None, device: Optional[torch.device] = None, rng: str = "numpy",
                    seed: int = 0) -> DeviceDataLoader:
    """dataset.py:100-138 with the same leading arguments; ``device`` / ``rng`` / ``seed`` are this package's."""
    return DeviceDataLoader(dataset, n_sample_points, batch_size, shuffle, consistent_sampling, augmentation_settings,
                            normalization, device, rng, seed)


def perturbate_point_cloud(xyz: np.ndarray, settings: AugmentationSettings, device: Optional[torch.device] = None
                           ) -> np.ndarray:
    """augmentation.py:143-167 for one cloud (N,3) on the device, numpy stream: jitter -> scale -> rotate -> shift."""
    N = xyz.shape[0]
    cache = CloudCache([(xyz, np.zeros((N, 0), np.float32), np.zeros((N,), np.int64))], device or "cuda")
    noise = torch.from_numpy(np.random.randn(N, 3).astype(np.float32)).unsqueeze(0)
    aug = draw_cloud_parameters(settings)[None]
    idx = torch.arange(N, dtype=torch.int32, device=cache.device).unsqueeze(0)
    out, _ = feed_batch(cache, [0], idx, None, aug, noise, settings)
    return out[0].cpu().numpy()


def broaden_annotation(point_cloud: torch.Tensor, annotation: torch.Tensor, radius: float = 0.01) -> torch.Tensor:
    """dataset.py:8-18: a point is annotated if it lies within ``radius`` of ANY annotated point.  The reference loops
    over the annotated points in python (O(A N) numpy); here one 1-NN search of every point among the annotated ones.
    point_cloud (N,3) fp32 CUDA, annotation (N,) -> uint8 (N,) on the device."""
    _cabi.require_cuda(point_cloud, "point_cloud")
    marked = point_cloud[annotation.to(point_cloud.device).bool()]
    if marked.shape[0] == 0:
        return torch.zeros(point_cloud.shape[0], dtype=torch.uint8, device=point_cloud.device)
    nn_ = ops.knn(marked.unsqueeze(0), point_cloud.unsqueeze(0), 1, idx64=False, idx32=True, dist=True)
    return (nn_["dist"][0, :, 0] < radius).to(torch.uint8)
