"""Synthetic point clouds for benchmarks and tests: there is no network for datasets.

``fingertip_cloud`` imitates what the reference trains on — an L515 depth frame of a hand above a
table (camera/realsense_camera.py:117 crops z to (0.05, 0.6); data/mock/*.npy span
x[-0.44,0.34] y[-0.31,0.30]) with fingertip annotations broadened to r = 0.01 (dataset.py:8-18):
a depth image back-projected through a pinhole model, so neighbouring points sit on a pixel grid and
exact d2 ties occur as on the real sensor."""
import numpy as np

from .preprocessing import sample_points


def fingertip_cloud(rng: np.random.RandomState, n_raw: int = 150_000):
    """-> (xyz float32 (n_raw,3), labels int64 (n_raw,)) with 1-5 fingertip blobs labelled 1."""
    w = int(np.sqrt(n_raw * 4 / 3)) + 1
    h = n_raw // w + 1
    u, v = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    u, v = u.ravel()[:n_raw], v.ravel()[:n_raw]
    cx, cy, f = w / 2, h / 2, 0.9 * w
    # table plane + a "hand": five elongated bumps (fingers) on a palm blob, random pose
    z = np.full(n_raw, 0.55, np.float32) + 0.02 * np.sin(u / w * 3.1).astype(np.float32)
    px, py = rng.uniform(0.35, 0.65) * w, rng.uniform(0.45, 0.7) * h
    palm = np.exp(-(((u - px) / (0.11 * w)) ** 2 + ((v - py) / (0.13 * h)) ** 2))
    z -= (0.18 * palm).astype(np.float32)
    tips = []
    n_fingers = rng.randint(1, 6)
    for i in range(n_fingers):
        ang = -1.2 + 0.55 * i + rng.normal(0, 0.05)
        length = rng.uniform(0.16, 0.24) * h
        tx, ty = px + np.sin(ang) * length, py - np.cos(ang) * length
        # distance of every pixel to the finger segment (palm centre -> tip)
        dx, dy = tx - px, ty - py
        t = np.clip(((u - px) * dx + (v - py) * dy) / (dx * dx + dy * dy), 0, 1)
        dseg = np.hypot(u - (px + t * dx), v - (py + t * dy))
        z -= (0.12 * np.exp(-(dseg / (0.018 * w)) ** 2)).astype(np.float32)
        tips.append((tx, ty))
    z += rng.normal(0, 4e-4, n_raw).astype(np.float32)          # sensor noise
    z = np.clip(z, 0.06, 0.59)
    x = (u - cx) / f * z
    y = (v - cy) / f * z
    xyz = np.stack([x, y, z], axis=1).astype(np.float32)
    labels = np.zeros(n_raw, np.int64)
    for tx, ty in tips:
        j = int(np.clip(round(ty), 0, h - 1)) * w + int(np.clip(round(tx), 0, w - 1))
        j = min(j, n_raw - 1)
        labels[np.linalg.norm(xyz - xyz[j], axis=1) < 0.01] = 1
    return xyz, labels


def fingertip_batch(seed: int, batch: int, n_points: int, n_raw: int = 150_000):
    """A training batch as the reference's data loader yields it (utils/dataset.py:75-80): every raw
    cloud randomly sub-sampled to ``n_points``.  -> (input (B,N,3) float32, labels (B,N) int64)."""
    rng = np.random.RandomState(seed)
    state = np.random.get_state()
    np.random.seed(seed)
    try:
        xs, ls = [], []
        for _ in range(batch):
            xyz, lab = fingertip_cloud(rng, n_raw)
            ids = sample_points(xyz.shape[0], n_points, consistent=False)
            xs.append(xyz[ids])
            ls.append(lab[ids])
    finally:
        np.random.set_state(state)
    return np.stack(xs), np.stack(ls)


def uniform_clouds(seed: int, batch: int, n_points: int) -> np.ndarray:
    """xyz ~ U[0,1)^3 fp32 (tie-free with high probability) — the KNN micro-benchmark input."""
    return np.random.RandomState(seed).rand(batch, n_points, 3).astype(np.float32)
