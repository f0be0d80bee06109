"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's data feeding, used as the checker of
csrc/feed.cu; pinned against the imported reference by oracle/make_golden.py (gen_feed), fixtures in
tests/golden/feed_golden.npz.  Never imported by the product.

Follows randlanet/utils/preprocessing.py:6-62 (sample_points), randlanet/utils/dataset.py:61-97 (preprocess),
randlanet/utils/augmentation.py:24-167 (perturbate_point_cloud) and dataset.py:8-18 (broaden_annotation).  All random
numbers come from numpy's global stream in the reference's order."""
import numpy as np


def sample_points(n_points, n_sample, consistent=False):                     # preprocessing.py:35-62
    def choice(size, replace):
        if consistent:                                                       # :23-31
            state = np.random.get_state()
            np.random.seed(0)
        v = np.random.choice(n_points, size, replace, None)
        if consistent:
            np.random.set_state(state)
        return v
    ids = choice(min(n_sample, n_points), False)
    if n_sample > n_points:
        ids = np.r_[ids, choice(n_sample - n_points, True)]
    return ids


def mean_radius(xyz):                                                        # augmentation.py:24-33
    return float(np.mean(np.linalg.norm(xyz - np.mean(xyz, axis=0, keepdims=True), axis=1)))


def rotation(angles):                                                        # augmentation.py:103-124
    cx, sx, cy, sy, cz, sz = (np.cos(angles[0]), np.sin(angles[0]), np.cos(angles[1]), np.sin(angles[1]),
                              np.cos(angles[2]), np.sin(angles[2]))
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def perturbate(xyz, s, record=None):
    """augmentation.py:143-167; s: any object with the AugmentationSettings fields.  ``record`` (a dict) receives the
    random numbers drawn: noise (N,3) and params [scale, 3 angles, 3 shifts]."""
    noise = np.random.randn(*xyz.shape)                                      # :48-52
    x = np.clip(mean_radius(xyz) * s.jitter_variance * noise, -s.jitter_limit, s.jitter_limit) + xyz
    scale = np.random.uniform(1 - s.scale_limit, 1 + s.scale_limit)          # :73-77
    c = np.mean(x, axis=0, keepdims=True)
    x = (x - c) * scale + c
    angles = [np.clip(sig * np.random.randn(), -lim, lim)                    # :99-102
              for sig, lim in zip(s.rotation_angle_variances, s.rotation_angle_limits)]
    c = np.mean(x, axis=0, keepdims=True)
    x = (x - c) @ rotation(angles).T + c                                     # :125-128
    shifts = np.random.uniform(-s.shift_limit, s.shift_limit, 3)             # :153-155
    x = x + mean_radius(x) * shifts
    if record is not None:
        record["noise"], record["params"] = noise, np.array([scale, *angles, *shifts])
    return x


def preprocess(xyz, features, labels, n_sample, consistent=True, augmentation=None, normalization=None, record=None):
    """dataset.py:61-97 -> (xyz (n,3), features (n,F), labels (n,))."""
    ids = sample_points(xyz.shape[0], n_sample, consistent)
    x, f, lab = xyz[ids], features[ids], labels[ids]
    if normalization is not None:                                            # :81-92
        x = x - np.mean(x, axis=0, keepdims=True)
        d = np.linalg.norm(x, axis=1)
        radius = {"mean": np.mean(d), "max": np.max(d), "stdev": np.std(d)}.get(normalization, 1.0)
        x = x / radius
    if augmentation:
        x = perturbate(x, augmentation, record)
    if record is not None:
        record["ids"] = ids
    return x, f, lab


def broaden_annotation(point_cloud, annotation, radius=0.01):                # dataset.py:8-18
    marked = point_cloud[annotation.astype(bool)]
    out = np.zeros(point_cloud.shape[0], dtype=bool)
    for p in marked:
        out |= np.abs(np.linalg.norm(p - point_cloud, axis=1)) < radius
    return out.astype(np.uint8)


# ---- the seeded cases of tests/golden/feed_golden.npz (written by oracle/make_golden.py from the reference)
FEED_N, FEED_SEED = 512, 1234


def feed_dataset():
    """Three small synthetic clouds (ragged sizes, 2 features): one larger than the sample size, one smaller (up-sampling
    with duplicates), one equal.  Clouds 0 and 1 are float64 arrays of float32-representable values, so the reference
    computes on them in float64 and a kernel working from an fp32 cache sees the same inputs; cloud 2 is float32, where
    the reference normalises in float32 (its centre is a sequential float32 sum: ~1e-5 relative, FEED_TOL)."""
    rng = np.random.RandomState(5)
    data = []
    for N, dtype in ((900, np.float64), (350, np.float64), (512, np.float32)):
        xyz = (rng.rand(N, 3) * [2.0, 1.0, 0.5] + [10.0, -3.0, 1.0]).astype(np.float32).astype(dtype)
        data.append((xyz, rng.rand(N, 2).astype(np.float32), (rng.rand(N) < 0.2).astype(np.int64)))
    return data


FEED_TOL = (2e-6, 2e-6, 5e-5)        # per cloud, relative to the largest coordinate
FEED_CASES = {"plain": (None, False), "mean": ("mean", False), "max_aug": ("max", True), "stdev_aug": ("stdev", True),
              "aug": (None, True), "centre_aug": ("other", True)}
