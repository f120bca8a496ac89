"""Synthetic meshes for the STL front-end tests (shared by the CPU and the GPU suite)."""
import numpy as np


def grid_mesh(w: int, h: int, seed: int = 0, jitter: float = 0.25):
    """a jittered height field: every inner vertex is shared by six triangles"""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:h, 0:w]
    v = np.stack([xs + jitter * rng.standard_normal((h, w)), ys + jitter * rng.standard_normal((h, w)),
                  np.sin(xs * 0.3) * np.cos(ys * 0.2) * 4], axis=-1).astype(np.float32).reshape(-1, 3)
    i = (ys[:-1, :-1] * w + xs[:-1, :-1]).reshape(-1)
    t = np.concatenate([np.stack([i, i + 1, i + w], 1), np.stack([i + 1, i + w + 1, i + w], 1)]).astype(np.uint32)
    t = t[rng.permutation(t.shape[0])]
    return v, t


def soup(nt: int, seed: int, pool: int):
    """nt triangles over a pool of `pool` candidate positions on a coarse lattice: many exact duplicates,
    equal x with different y, equal (x, y) with different z, negative values"""
    rng = np.random.default_rng(seed)
    v = (rng.integers(-8, 8, (pool, 3)) * 0.5).astype(np.float32)
    t = rng.integers(0, pool, (nt, 3)).astype(np.uint32)
    return v, t


def cases():
    out = {}
    for nt in (1, 2, 85, 86, 255, 256, 257, 1365, 1366, 4097):
        out[f"soup{nt}"] = soup(nt, nt, max(3, nt // 2))
    out["soup_dense"] = soup(20000, 7, 50)               # 60,000 corners on 50 positions
    out["distinct"] = (np.random.default_rng(3).standard_normal((3 * 5000, 3)).astype(np.float32),
                       np.arange(3 * 5000, dtype=np.uint32).reshape(-1, 3))
    out["grid"] = grid_mesh(120, 90, 5)
    one = np.zeros((3, 3), np.float32)
    out["one_point"] = (one, np.zeros((700, 3), np.uint32))   # every corner the same vertex: no sort pass runs
    big = np.array([[1e30, -1e30, 3e-39], [-3e-39, 1e-45, -1e-45], [np.inf, -np.inf, 2.5], [2.5, np.inf, -np.inf]], np.float32)
    out["extremes"] = (big, np.array([[0, 1, 2], [3, 2, 1], [1, 1, 0]], np.uint32))
    z = np.array([[0.0, 1.0, -0.0], [-0.0, 1.0, 0.0], [0.0, -0.0, 5.0], [-0.0, 0.0, 5.0], [-0.0, 3.0, 1.0], [0.0, 2.0, 1.0]], np.float32)
    out["signed_zero"] = (z, np.array([[0, 2, 4], [1, 3, 5], [5, 4, 0], [3, 1, 2]], np.uint32))   # not compared with the reference
    return out
