"""CPU: the oracle's data-feeding restatement (oracle/feeding.py) against the fixtures written from the reference's
PointCloudPreprocessor + perturbate_point_cloud (tests/golden/feed_golden.npz, oracle/make_golden.py gen_feed), and the
host half of the product's loader: the order in which it consumes numpy's global stream."""
import importlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import feeding


@pytest.mark.parametrize("consistent", [True, False])
@pytest.mark.parametrize("name", list(feeding.FEED_CASES))
def test_oracle_feeding_vs_reference_golden(name, consistent):
    g = np.load(os.path.join(GOLDEN, "feed_golden.npz"))
    ds = importlib.import_module("3d_recognizer_b200.dataset")
    norm, aug = feeding.FEED_CASES[name]
    data = feeding.feed_dataset()
    np.random.seed(feeding.FEED_SEED)
    for i, item in enumerate(data):
        x, f, lab = feeding.preprocess(*item, feeding.FEED_N, consistent, ds.AugmentationSettings() if aug else None, norm)
        ref = g[f"{name}/{int(consistent)}/{i}/input"]
        assert np.array_equal(lab, g[f"{name}/{int(consistent)}/{i}/labels"])
        assert np.abs(np.concatenate((x, f), axis=1) - ref).max() < 1e-6 * np.abs(ref).max()
        assert ref.dtype == np.float32 and item[0].dtype == (np.float32 if i == 2 else np.float64)


def test_cloud_parameters_follow_the_reference_draw_order():
    """draw_cloud_parameters consumes the stream exactly like scale -> rotate -> shift of augmentation.py, so that after
    the (N,3) jitter normals the loader's numbers are the reference's."""
    ds = importlib.import_module("3d_recognizer_b200.dataset")
    s = ds.AugmentationSettings(scale_limit=0.3, shift_limit=0.2, rotation_angle_variances=(0.5, 0.06, 0.01))
    xyz = np.random.RandomState(0).rand(100, 3)
    np.random.seed(9)
    rec = {}
    feeding.perturbate(xyz, s, rec)
    after_oracle = np.random.rand()
    np.random.seed(9)
    noise = np.random.randn(100, 3)
    params = ds.draw_cloud_parameters(s)
    assert np.array_equal(noise, rec["noise"]) and np.array_equal(params, rec["params"])
    assert np.random.rand() == after_oracle
    assert abs(params[1]) <= 0.18                       # clipped angle (sigma 0.5 against the 0.18 limit)


def test_settings_defaults_match_reference():
    ds = importlib.import_module("3d_recognizer_b200.dataset")
    s = ds.AugmentationSettings()
    assert (s.jitter_variance, s.jitter_limit, s.scale_limit, s.shift_limit) == (0.01, 0.05, 0.2, 0.1)
    assert s.rotation_angle_variances == (0.06, 0.06, 0.06) and s.rotation_angle_limits == (0.18, 0.18, 0.18)
    with pytest.raises(RuntimeError):
        ds.CloudCache(feeding.feed_dataset(), "cpu")      # the cache lives in GPU memory: no CPU path
