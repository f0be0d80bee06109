"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden-vector generator; runs ONLY in the build container.

Imports the UNMODIFIED reference package from /root/reference (read-only) with
  * a stand-in ``faiss`` module (faiss-cpu 1.7.2 is pinned in requirements.txt:12 but absent here and
    approximate by construction — SURVEY.md F3/F8), and
  * ``randlanet.utils.modules.knn_naive`` / ``knn_approximate`` rebound to the canonical exact KNN
    (the reference's intended kdtree wiring, modules.py:135-138),
runs it on seeded inputs, checks the oracle restatement (oracle/network.py, oracle/knn.py) against it,
and writes small fixtures to tests/golden/.  Nothing here travels to the GPU box except the fixtures.

    python -m oracle.make_golden
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

from . import network as onet          # noqa: E402
from .knn import knn_exact, ref_knn_tpk, build  # noqa: E402


def import_reference():
    sys.modules.setdefault("faiss", types.ModuleType("faiss"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import randlanet.utils.modules as rmod  # noqa

    def exact(xyz, xyz_query, n_neighbors, *a, **kw):
        idx, d2 = knn_exact(xyz.detach().cpu().numpy(), xyz_query.detach().cpu().numpy(), n_neighbors)
        return torch.from_numpy(idx), torch.from_numpy(d2)

    rmod.knn_naive = exact
    rmod.knn_approximate = exact
    return rmod


def knn_cases():
    """name -> (support (B,Ns,3), query (B,Nq,3), K)"""
    rng = np.random.RandomState(1234)
    cases = {}
    s = rng.rand(2, 2048, 3).astype(np.float32)
    cases["uniform_self_k16"] = (s, s, 16)
    cases["uniform_self_k32"] = (s[:1, :1500], s[:1, :1500], 32)
    cases["uniform_cross_k1"] = (s[:, :512], s, 1)
    cases["uniform_cross_k8"] = (s[:, :700], rng.rand(2, 333, 3).astype(np.float32), 8)
    # lattice: massive exact ties
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(9), indexing="ij"), -1).reshape(-1, 3)
    lat = (g.astype(np.float32) * 0.125)[rng.permutation(len(g))][None]
    cases["lattice_ties_k16"] = (lat, lat, 16)
    # duplicated points (sample_points duplicates when n_points > N, preprocessing.py:55-61)
    base = rng.rand(1, 300, 3).astype(np.float32)
    dup = base[:, rng.randint(0, 300, 1100)]
    cases["duplicates_k16"] = (dup, dup, 16)
    cases["ns_equals_k"] = (s[:1, :32], s[:1, :77], 32)
    # real LiDAR geometry (exact d2 ties)
    mock = os.path.join(REF, "data", "mock", "2022_07_29__09_32_07_088822000_data.npy")
    cloud = np.load(mock).astype(np.float32)
    ids = onet.sample_points(cloud.shape[0], 4096, consistent=True)
    m = cloud[ids][None]
    cases["mock_lidar_k16"] = (m, m, 16)
    cases["mock_lidar_k32"] = (m[:, :2500], m[:, :2500], 32)
    return cases


def gen_knn():
    out = {}
    for name, (s, q, k) in knn_cases().items():
        idx, d2 = knn_exact(s, q, k)
        ridx, rd2 = ref_knn_tpk(s, q, k)
        assert np.array_equal(d2, rd2), f"{name}: oracle d2 differs from the reference's nanoflann"
        # tie-free rows (incl. the K-th boundary) must agree on indices too
        if s.shape[1] > k:
            _, d2p = knn_exact(s, q, k + 1)
            tie = (d2p[..., 1:] == d2p[..., :-1]).any(-1)
        else:
            tie = (d2[..., 1:] == d2[..., :-1]).any(-1)
        assert np.array_equal(idx[~tie], ridx[~tie]), f"{name}: indices differ on tie-free rows"
        print(f"knn {name}: d2 bit-identical to reference; tied rows {tie.mean():.3f}; "
              f"idx equal on all tie-free rows")
        out[name + "/support"] = s
        out[name + "/query"] = q
        out[name + "/k"] = np.int64(k)
        out[name + "/idx"] = idx.astype(np.int32)
        out[name + "/d2"] = d2
        out[name + "/ref_idx"] = ridx.astype(np.int32)
        out[name + "/tied"] = tie
    np.savez_compressed(os.path.join(GOLD, "knn_golden.npz"), **out)


E2E = [
    # name, settings, B, N, seed
    ("k16_n1024", dict(n_classes=2, n_points=1024, n_features=0, n_neighbors=16, knn="naive"), 2, 1024, 11),
    ("k32_n2500", dict(n_classes=2, n_points=2500, n_features=0, n_neighbors=32, knn="naive"), 1, 2500, 12),
    ("k16_f2_c3_n1100", dict(n_classes=3, n_points=1100, n_features=2, n_neighbors=16, knn="approximate"), 2, 1100, 13),
]


def make_input(B, N, F, seed):
    rng = np.random.RandomState(seed)
    x = rng.rand(B, N, 3 + F).astype(np.float32)
    x[..., :3] = x[..., :3] * np.array([0.78, 0.61, 0.55], np.float32) + np.array([-0.44, -0.31, 0.05], np.float32)
    return x


# settings outside the template lists of the fused kernels (any n_neighbors / layer size: modules.py:298-325, 484-500):
# served by the row-form kernels (csrc/lfa_rows.cu); the oracle port is pinned to the reference on them as well
E2E_ROWS = [
    ("k8_sizes_8_24_40_n1024", dict(n_classes=2, n_points=1024, n_features=0, n_neighbors=8, layer_sizes=[8, 24, 40],
                                    knn="naive"), 2, 1024, 131),
    ("k20_n1600", dict(n_classes=2, n_points=1600, n_features=0, n_neighbors=20, knn="naive"), 2, 1600, 32),
    ("k16_sizes_16_48_96_256_n2048", dict(n_classes=2, n_points=2048, n_features=0, n_neighbors=16,
                                          layer_sizes=[16, 48, 96, 256], knn="naive"), 1, 2048, 34),
]


def gen_e2e(rmod, cases=None, filename="e2e_golden.npz"):
    out = {}
    upsampler_too = cases is None
    for name, st, B, N, seed in (E2E if cases is None else cases):
        settings = rmod.RandLANetSettings(**st)
        net = rmod.RandLANet(settings, torch.device("cpu"))
        schema = onet.state_dict_schema(st)
        ref_sd = net.state_dict()
        assert list(ref_sd.keys()) == list(schema.keys()), "state_dict key order/schema differs"
        for k_, v in ref_sd.items():
            assert tuple(v.shape) == schema[k_][0] and v.dtype == schema[k_][1], k_
        sd = onet.synth_state_dict(st, seed)
        net.load_state_dict(sd)
        x = torch.from_numpy(make_input(B, N, st["n_features"], seed))
        labels = torch.from_numpy(np.random.RandomState(seed).randint(0, st["n_classes"], (B, N)))

        # ---- eval forward
        net.eval()
        np.random.seed(seed)
        with torch.no_grad():
            ref_logits = net(x)
        np.random.seed(seed)
        perm = np.random.permutation(N)
        with torch.no_grad():
            o_logits = onet.forward({k_: v.clone() for k_, v in sd.items()}, st, x, perm)
        err = (ref_logits - o_logits).abs().max().item() / ref_logits.abs().max().item()
        print(f"e2e {name}: eval   max|oracle-ref|/max|ref| = {err:.3e}")
        assert err < 1e-5
        out[f"{name}/eval_logits"] = ref_logits.numpy()

        # ---- train forward + backward (Dropout disabled so the mask does not enter parity)
        net.train()
        net.fc_end[2].p = 0.0
        np.random.seed(seed)
        logits = net(x)
        loss = onet.dice_loss(logits, labels)
        net.zero_grad()
        loss.backward()
        ref_grads = {k_: p.grad.clone() for k_, p in net.named_parameters()}
        ref_after = {k_: v.clone() for k_, v in net.state_dict().items()}

        osd = {k_: v.clone().requires_grad_(v.dtype == torch.float32 and "running" not in k_)
               for k_, v in sd.items()}
        o_logits = onet.forward(osd, st, x, perm, training=True, dropout_p=0.0)
        o_loss = onet.dice_loss(o_logits, labels)
        o_loss.backward()
        gerr, gname = onet.grad_parity({k_: osd[k_].grad for k_ in ref_grads}, ref_grads)
        lerr = (logits - o_logits).abs().max().item() / logits.abs().max().item()
        rerr = max((osd[k_].detach() - ref_after[k_]).abs().max().item() for k_ in sd if "running" in k_)
        print(f"e2e {name}: train  logits rel {lerr:.3e}  worst grad rel {gerr:.3e} ({gname})  running-stat abs {rerr:.3e}")
        assert lerr < 1e-5 and gerr < 1e-4 and rerr < 1e-4
        out[f"{name}/train_logits"] = logits.detach().numpy()
        out[f"{name}/train_loss"] = np.float32(loss.item())
        for k_, g in ref_grads.items():
            out[f"{name}/grad/{k_}"] = onet.grad_fixture_view(g).numpy().astype(np.float32)
            out[f"{name}/gradnorm/{k_}"] = np.float32(g.double().norm().item())
        for k_, v in ref_after.items():
            if "running" in k_:
                out[f"{name}/after/{k_}"] = v.numpy()
        out[f"{name}/perm"] = perm.astype(np.int32)

    if not upsampler_too:
        np.savez_compressed(os.path.join(GOLD, filename), **out)
        return
    # ---- UpSampler variants (modules.py:416-456) on reference module
    rng = np.random.RandomState(5)
    feat = torch.from_numpy(rng.rand(2, 5, 200, 1).astype(np.float32))
    xyz = torch.from_numpy(rng.rand(2, 200, 3).astype(np.float32))
    xyz_up = torch.from_numpy(rng.rand(2, 901, 3).astype(np.float32))
    for ap in ["nni", "nna", "idw", "isdw", "none"]:
        up = rmod.UpSampler(ap, torch.device("cpu"))
        r = up(feat, xyz, xyz_up)
        o = onet.upsample(ap, feat, xyz, xyz_up)
        assert torch.allclose(r, o, rtol=1e-6, atol=1e-7), ap
        out[f"upsample/{ap}"] = r.numpy()
    out["upsample/feat"], out["upsample/xyz"], out["upsample/xyz_up"] = feat.numpy(), xyz.numpy(), xyz_up.numpy()
    print("upsampler variants: oracle == reference")
    np.savez_compressed(os.path.join(GOLD, filename), **out)


def gen_predict(rmod):
    """Model.predict (model.py:146-235) through the reference façade on a mock LiDAR cloud with a
    synthesised checkpoint (the shipped one is a stripped blob)."""
    sys.modules.setdefault("faiss", types.ModuleType("faiss"))
    import randlanet.model as rmodel
    import tempfile
    from pathlib import Path
    st = dict(n_classes=2, n_points=2500, n_features=0, n_neighbors=32, knn="naive")   # train.py:50-51
    settings = rmod.RandLANetSettings(**st)
    sd = onet.synth_state_dict(st, 21)
    model = rmodel.Model(settings, weights=sd, use_gpu=False)
    cloud = np.load(os.path.join(REF, "data", "mock", "2022_07_29__14_15_36_850990000_data.npy")).astype(np.float32)
    cloud = cloud[:: 8]          # keep the fixture small: ~1/8 of the frame
    out = {}
    for ap in ["nni", "idw"]:
        model._model.settings.upsampling = ap
        model._upsampler = rmod.UpSampler(ap, torch.device("cpu"))
        np.random.seed(3)
        conf = model.predict(cloud)
        np.random.seed(3)
        oconf = onet.predict({k_: v.clone() for k_, v in sd.items()}, dict(st, upsampling=ap), cloud)
        err = np.abs(conf - oconf).max()
        print(f"predict {ap}: max|oracle-ref| = {err:.3e} on {cloud.shape}")
        assert err < 1e-5
        out[f"conf_{ap}"] = conf.astype(np.float32)
    out["cloud"] = cloud
    # checkpoint written by the REFERENCE's Model.save (zip: config + model, model.py:107-121)
    with tempfile.TemporaryDirectory() as tmp:
        model._model.settings.upsampling = "nni"
        p = Path(tmp) / "ckpt"
        model.save(p)
        data = p.read_bytes()
    # the archive itself (written by the reference's Model.save, 5 MB): the product's Model.load must read it as is
    with open(os.path.join(GOLD, "ref_checkpoint.zip"), "wb") as f:
        f.write(data)
    import zipfile, io, json
    z = zipfile.ZipFile(io.BytesIO(data))
    out["ckpt_config_json"] = np.frombuffer(z.read("config"), dtype=np.uint8)
    out["ckpt_names"] = np.array(sorted(z.namelist()))
    np.savez_compressed(os.path.join(GOLD, "predict_golden.npz"), **out)


LOSS_PARAMS = {"dice": (0.5, 1.0), "tversky": (0.7, 1.0), "focal_tversky": (0.7, 4.0 / 3.0)}   # trainer.py:245-269


def gen_loss():
    """The reference's FocalTverskyLoss (losses.py:59-87) with the three parameter sets of Trainer._get_loss
    (trainer.py:245-269) on seeded logits: loss value and d loss / d logits; also pins oracle.network.dice_loss."""
    import_reference()
    from randlanet.utils.losses import FocalTverskyLoss
    from randlanet.utils.metrics import accuracy, iou
    out = {}
    rng = np.random.RandomState(77)
    for cname, (B, C, N) in {"b2c2n300": (2, 2, 300), "b3c3n257": (3, 3, 257), "b1c5n64": (1, 5, 64)}.items():
        logits = (rng.randn(B, C, N) * 2.0).astype(np.float32)
        labels = rng.randint(0, C, (B, N)).astype(np.int64)
        out[f"{cname}/logits"], out[f"{cname}/labels"] = logits, labels
        # metrics.py:8-59 on the same batch (also with a class that never occurs: labels capped at C - 2)
        for tag, lab in (("", labels), ("absent/", np.minimum(labels, max(C - 2, 0)))):
            oa, pca = accuracy(torch.from_numpy(logits), torch.from_numpy(lab))
            miou, pci = iou(torch.from_numpy(logits), torch.from_numpy(lab))
            out[f"{cname}/{tag}metrics"] = np.array([oa, miou] + list(pca) + list(pci), dtype=np.float64)
        for name, (alpha, gamma) in LOSS_PARAMS.items():
            x = torch.from_numpy(logits).requires_grad_(True)
            loss = FocalTverskyLoss(alpha=alpha, gamma=gamma, neglect_background=True)(x, torch.from_numpy(labels))
            loss.backward()
            xo = torch.from_numpy(logits).requires_grad_(True)
            lo = onet.dice_loss(xo, torch.from_numpy(labels), alpha, gamma)
            lo.backward()
            assert abs(float(lo) - float(loss)) < 1e-7 and float((xo.grad - x.grad).abs().max()) < 1e-7 * float(
                x.grad.abs().max()) + 1e-12, f"oracle loss differs from the reference ({cname}, {name})"
            out[f"{cname}/{name}/loss"] = np.float32(loss.item())
            out[f"{cname}/{name}/dlogits"] = x.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "loss_golden.npz"), **out)
    print("loss golden:", len(out), "arrays; oracle.network.dice_loss pinned against the reference's FocalTverskyLoss")


def gen_feed():
    """The reference's PointCloudPreprocessor (utils/dataset.py:11-97) + perturbate_point_cloud on a seeded numpy stream,
    consistent and non-consistent sampling: every item's (input, labels); also pins oracle.feeding."""
    import_reference()
    from randlanet.utils.augmentation import AugmentationSettings
    from randlanet.utils.dataset import PointCloudPreprocessor
    from . import feeding
    data, n = feeding.feed_dataset(), feeding.FEED_N
    out = {}
    for name, (norm, aug) in feeding.FEED_CASES.items():
        for consistent in (True, False):
            s = AugmentationSettings() if aug else None
            pre = PointCloudPreprocessor(data, n, consistent_sampling=consistent, augmentation_settings=s,
                                         normalization=norm)
            np.random.seed(feeding.FEED_SEED)
            ref = [pre[i] for i in range(len(data))]
            np.random.seed(feeding.FEED_SEED)
            for i, (inp, lab, idx) in enumerate(ref):
                x, f, l2 = feeding.preprocess(*data[i], n, consistent, s, norm)
                mine = np.concatenate((x, f), axis=1).astype(np.float32)
                assert idx == i and np.array_equal(l2, lab.numpy())
                assert np.abs(mine - inp.numpy()).max() < 1e-6 * np.abs(inp.numpy()).max(), (name, consistent, i)
                out[f"{name}/{int(consistent)}/{i}/input"] = inp.numpy()
                out[f"{name}/{int(consistent)}/{i}/labels"] = lab.numpy()
    # (broaden_annotation lives in the top-level dataset.py, which does not import under numpy >= 1.24 — np.bool,
    # SURVEY F11; oracle.feeding restates its ten lines and has no fixture)
    np.savez_compressed(os.path.join(GOLD, "feed_golden.npz"), **out)
    print("feed golden:", len(out), "arrays; oracle.feeding pinned against the reference's PointCloudPreprocessor")


def main():
    os.makedirs(GOLD, exist_ok=True)
    build(ref=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if "--only-rows" in sys.argv:          # add the row-form fixtures without rewriting the others
        gen_e2e(import_reference(), E2E_ROWS, "e2e_rows_golden.npz")
        return
    gen_knn()
    rmod = import_reference()
    gen_e2e(rmod)
    gen_e2e(rmod, E2E_ROWS, "e2e_rows_golden.npz")
    gen_predict(rmod)
    gen_loss()
    gen_feed()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
