"""Host-side random sub/up-sampling index draw used by ``Model.predict`` and the data loader
(reference: randlanet/utils/preprocessing.py:6-62).  The indices come from numpy's legacy global
MT19937 stream so that they are bit-identical to the reference's for the same seed; that forces this
to stay on the host."""
from contextlib import contextmanager

import numpy as np


@contextmanager
def _fixed_seed(active: bool):
    """``consistent`` draws reseed the GLOBAL generator to 0 and put its state back afterwards
    (preprocessing.py:23-31)."""
    if not active:
        yield
        return
    saved = np.random.get_state()
    np.random.seed(0)
    try:
        yield
    finally:
        np.random.set_state(saved)


def sample_points(n_points: int, n_sample_points: int, consistent: bool = False) -> np.ndarray:
    """Indices of ``n_sample_points`` points out of ``n_points``: a draw without replacement of
    min(n_sample_points, n_points), followed — when more points are wanted than exist — by a draw
    WITH replacement of the remainder (duplicated points; preprocessing.py:49-61)."""
    with _fixed_seed(consistent):
        ids = np.random.choice(n_points, min(n_sample_points, n_points), False, None)
    extra = n_sample_points - n_points
    if extra > 0:
        with _fixed_seed(consistent):
            dup = np.random.choice(n_points, extra, True, None)
        ids = np.r_[ids, dup]
    return ids
