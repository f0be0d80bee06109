"""Checkpoint + UI integration: a zip written by the REFERENCE's Model.save is loaded as is, the
Predictor's 30-point warm-up call works (predict.py:22-24), and the library is usable from a spawned child process while
the parent holds its own CUDA context (train.py:108-115 trains in a spawned child, main.py:71-89 predicts in the parent)."""
import importlib
import multiprocessing as mp
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import network as onet

pytestmark = pytest.mark.gpu
CKPT = os.path.join(GOLDEN, "ref_checkpoint.zip")


def test_load_checkpoint_written_by_reference():
    """tests/golden/ref_checkpoint.zip was written by randlanet.model.Model.save (model.py:107-121) in
    oracle/make_golden.py; predictions from it must equal the reference façade's own (predict_golden.npz)."""
    from pathlib import Path
    model_mod = importlib.import_module("3d_recognizer_b200.model")
    g = np.load(os.path.join(GOLDEN, "predict_golden.npz"))
    for ap in ("nni", "idw"):
        m = model_mod.Model.load(Path(CKPT), upsampling=ap)
        assert m.settings.n_points == 2500 and m.settings.n_neighbors == 32 and m.settings.upsampling == ap
        np.random.seed(3)
        conf = m.predict(g["cloud"])
        assert conf.shape == g[f"conf_{ap}"].shape
        assert np.abs(conf - g[f"conf_{ap}"]).max() < 1e-4
    sd = m.module.state_dict()
    ref = onet.synth_state_dict(dict(n_classes=2, n_points=2500, n_features=0, n_neighbors=32, knn="naive"), 21)
    assert set(sd) == set(ref) and all(torch.equal(sd[k].cpu(), ref[k]) for k in ref)


def test_predictor_warmup_cloud_of_30_points():
    """predict.py:22-24: the Predictor warms the model up with a random 30-point cloud.  30 < n_points, so
    sample_points up-samples WITH duplicates (preprocessing.py:55-61): every KNN row is full of exact ties."""
    from pathlib import Path
    model_mod = importlib.import_module("3d_recognizer_b200.model")
    m = model_mod.Model.load(Path(CKPT))
    rng = np.random.RandomState(0)
    cloud = rng.random_sample((30, 3))
    np.random.seed(4)                                    # the forward draws its down-sampling permutation from the global RNG
    conf = m.predict(cloud)
    assert conf.shape == (2, 30) and np.isfinite(conf).all()
    assert np.allclose(conf.sum(axis=0), 1.0, atol=1e-5)
    st = dict(n_classes=2, n_points=2500, n_features=0, n_neighbors=32, knn="naive", upsampling="nni")
    np.random.seed(4)
    ref = onet.predict(onet.synth_state_dict(st, 21), st, cloud.astype(np.float32))
    assert np.abs(conf - ref).max() < 1e-4


def _child(ckpt, q):
    try:
        import importlib as il
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from pathlib import Path
        mm = il.import_module("3d_recognizer_b200.model")
        syn = il.import_module("3d_recognizer_b200.synthetic")
        m = mm.Model.load(Path(ckpt))
        x, y = syn.fingertip_batch(3, 2, 2500, n_raw=20000)
        opt = m.make_optimizer(1e-2)
        losses = [float(m.train_step(torch.from_numpy(x), torch.from_numpy(y), opt, "dice")) for _ in range(3)]
        conf = m.predict(x[0])
        q.put(("ok", losses, conf.shape, bool(np.isfinite(conf).all())))
    except Exception as e:          # noqa: BLE001 - reported to the parent
        q.put(("error", repr(e)))


@pytest.mark.timeout(300)
def test_spawned_child_trains_while_parent_predicts():
    from pathlib import Path
    model_mod = importlib.import_module("3d_recognizer_b200.model")
    parent = model_mod.Model.load(Path(CKPT))
    cloud = np.random.RandomState(1).random_sample((5000, 3)).astype(np.float32)
    def predict():
        np.random.seed(5)                                # same down-sampling permutation every time
        return parent.predict(cloud)

    before = predict()                                   # CUDA context + library live in the parent
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_child, args=(CKPT, q))
    p.start()
    during = [predict() for _ in range(5)]               # the Tk thread keeps predicting meanwhile
    res = q.get(timeout=240)
    p.join(timeout=60)
    assert res[0] == "ok", res
    assert p.exitcode == 0
    assert all(np.isfinite(v) for v in res[1]) and res[2] == (2, 2500) and res[3]
    assert all(np.array_equal(before, d) for d in during)
