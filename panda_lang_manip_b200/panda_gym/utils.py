"""Drop-in for reference panda_gym/utils.py:4-30 (host-side numpy helpers; the batched GPU versions are
panda_lang_manip_b200.compute_reward / is_success)."""
import numpy as np


def distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """utils.py:4-15 -- Euclidean distance over the last axis."""
    assert a.shape == b.shape
    return np.linalg.norm(a - b, axis=-1)


def angle_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """utils.py:18-30 -- 1 - <a, b>^2, row-wise (the reference's np.inner form returns an [N, N] matrix for batches,
    SURVEY App. E.4; identical for the single goals the env itself passes)."""
    assert a.shape == b.shape
    return 1 - np.einsum("...i,...i->...", a, b) ** 2
