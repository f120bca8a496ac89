"""Deterministic synthetic meshes for the benchmarks and the large parity tests.

Grid W x H, vertex i = y*W + x:
    px = 0.01*x + j*0.01*(u0 - 0.5)
    py = 0.01*y + j*0.01*(u1 - 0.5)
    pz = 5*sin(0.37*px)*cos(0.21*py) + j*0.01*(u2 - 0.5)
with u_k a counter-based hash of (seed, 3*i + k) (splitmix64), so the same arrays can be produced
on the host (numpy) and on the device (torch) without a sequential generator.  Two triangles per
cell; vertex ids are then shuffled inside consecutive blocks of 64 so the index byte planes look
like a real scanned mesh rather than a regular grid.  Data only - no codec logic lives here.
"""
from __future__ import annotations

import numpy as np

_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_G = 0x9E3779B97F4A7C15


def _splitmix_np(i: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = i + np.uint64(_G)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        return z ^ (z >> np.uint64(31))


def _u_np(i):
    return (_splitmix_np(i) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def grid_vertices(W: int, H: int, jitter: float = 1.0, seed: int = 0, dtype=np.float32) -> np.ndarray:
    n = W * H
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64)
        k = i * np.uint64(3) + np.uint64(seed) * np.uint64(0x100000001B3)
        x = (i % np.uint64(W)).astype(np.float64)
        y = (i // np.uint64(W)).astype(np.float64)
        px = 0.01 * x + jitter * 0.01 * (_u_np(k) - 0.5)
        py = 0.01 * y + jitter * 0.01 * (_u_np(k + np.uint64(1)) - 0.5)
        pz = 5.0 * np.sin(0.37 * px) * np.cos(0.21 * py) + jitter * 0.01 * (_u_np(k + np.uint64(2)) - 0.5)
    return np.stack([px, py, pz], axis=1).astype(dtype)


def block_permutation(n: int, block: int = 64, seed: int = 0) -> np.ndarray:
    """perm[old] = new, a random permutation inside every consecutive block of `block` ids"""
    nb = (n + block - 1) // block
    with np.errstate(over="ignore"):
        keys = _splitmix_np(np.arange(nb * block, dtype=np.uint64) + np.uint64(seed + 12345) * np.uint64(0x100000001B3))
    keys = keys.reshape(nb, block)
    # ids beyond n sort last so they never displace a real id
    valid = (np.arange(nb * block).reshape(nb, block) < n)
    keys = np.where(valid, keys >> np.uint64(2), np.uint64(2 ** 62) + np.arange(block, dtype=np.uint64)[None, :])
    order = np.argsort(keys, axis=1, kind="stable")                  # order[b, slot] = old local id
    perm = np.empty(nb * block, dtype=np.int64)
    base = (np.arange(nb, dtype=np.int64) * block)[:, None]
    perm[(base + order).reshape(-1)] = (base + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return perm[:n]


def grid_triangles(W: int, H: int, dtype=np.uint32) -> np.ndarray:
    x = np.arange(W - 1, dtype=np.int64)
    y = np.arange(H - 1, dtype=np.int64)
    i = (y[:, None] * W + x[None, :]).reshape(-1)
    a = np.stack([i, i + 1, i + W], axis=1)
    b = np.stack([i + 1, i + W + 1, i + W], axis=1)
    return np.stack([a, b], axis=1).reshape(-1, 3).astype(dtype)


def grid_mesh(W: int, H: int, jitter: float = 1.0, seed: int = 0, vdtype=np.float32, idtype=np.uint32, shuffle: int = 64):
    """-> (vertices [n,3], triangles [2(W-1)(H-1),3]) with ids shuffled in blocks of `shuffle`"""
    v = grid_vertices(W, H, jitter, seed, vdtype)
    t = grid_triangles(W, H, np.int64)
    if shuffle:
        perm = block_permutation(W * H, shuffle, seed)
        vs = np.empty_like(v)
        vs[perm] = v
        v = vs
        t = perm[t]
    return np.ascontiguousarray(v), np.ascontiguousarray(t.astype(idtype))


def colors_rgba(v: np.ndarray, seed: int = 0) -> np.ndarray:
    """u32 RGBA per point, memory order r,g,b,a"""
    n = v.shape[0]
    with np.errstate(over="ignore"):
        k = np.arange(n, dtype=np.uint64) * np.uint64(3) + np.uint64(seed + 99) * np.uint64(0x100000001B3)
    d = v.astype(np.float64)
    r = 128 + 100 * np.sin(0.5 * d[:, 0]) + 8 * (_u_np(k) - 0.5)
    g = 128 + 100 * np.sin(0.5 * d[:, 1]) + 8 * (_u_np(k + np.uint64(1)) - 0.5)
    b = 128 + 20 * d[:, 2] + 8 * (_u_np(k + np.uint64(2)) - 0.5)
    r, g, b = (np.clip(c, 0, 255).astype(np.uint32) for c in (r, g, b))
    return r | (g << np.uint32(8)) | (b << np.uint32(16)) | np.uint32(255 << 24)


# ------------------------------------------------------------------------------------ torch (GPU)
def grid_mesh_torch(W: int, H: int, device, jitter: float = 1.0, seed: int = 0, double: bool = False, long_index: bool = False, shuffle: int = 64, triangles: bool = True):
    """Same mesh family generated on `device` with torch (bench sizes: 1e8 vertices in < 1 s).
    Not bit-identical to the numpy path (device sin/cos); both bench arms consume THESE arrays."""
    import torch

    def splitmix(i):
        # int64 arithmetic wraps; logical right shifts are emulated by masking the sign extension
        def lsr(z, s):
            return (z >> s) & ((1 << (64 - s)) - 1)
        z = i + (_G - (1 << 64))
        z = (z ^ lsr(z, 30)) * (_M1 - (1 << 64))
        z = (z ^ lsr(z, 27)) * (_M2 - (1 << 64))
        return z ^ lsr(z, 31)

    def u(i):
        return ((splitmix(i) >> 11) & ((1 << 53) - 1)).to(torch.float64) * 2.0 ** -53

    n = W * H
    i = torch.arange(n, dtype=torch.int64, device=device)
    k = i * 3 + seed * 0x100000001B3
    x = (i % W).to(torch.float64)
    y = (i // W).to(torch.float64)
    px = 0.01 * x + jitter * 0.01 * (u(k) - 0.5)
    py = 0.01 * y + jitter * 0.01 * (u(k + 1) - 0.5)
    pz = 5.0 * torch.sin(0.37 * px) * torch.cos(0.21 * py) + jitter * 0.01 * (u(k + 2) - 0.5)
    v = torch.stack([px, py, pz], dim=1).to(torch.float64 if double else torch.float32)
    del px, py, pz, x, y, k
    if not triangles:                       # point cloud: no connectivity (and nothing to shuffle against)
        return v.contiguous(), None
    xs = torch.arange(W - 1, dtype=torch.int64, device=device)
    ys = torch.arange(H - 1, dtype=torch.int64, device=device)
    c = (ys[:, None] * W + xs[None, :]).reshape(-1)
    t = torch.stack([torch.stack([c, c + 1, c + W], 1), torch.stack([c + 1, c + W + 1, c + W], 1)], 1).reshape(-1, 3)
    del c
    if shuffle:
        nb = (n + shuffle - 1) // shuffle
        ids = torch.arange(nb * shuffle, dtype=torch.int64, device=device)
        keys = (splitmix(ids + (seed + 12345) * 0x100000001B3) >> 2) & ((1 << 62) - 1)
        keys = torch.where(ids < n, keys, (1 << 62) + (ids % shuffle)).reshape(nb, shuffle)
        order = torch.argsort(keys, dim=1, stable=True)
        base = (torch.arange(nb, dtype=torch.int64, device=device) * shuffle)[:, None]
        perm = torch.empty(nb * shuffle, dtype=torch.int64, device=device)
        perm[(base + order).reshape(-1)] = (base + torch.arange(shuffle, dtype=torch.int64, device=device)[None, :]).reshape(-1)
        perm = perm[:n]
        vs = torch.empty_like(v)
        vs[perm] = v
        v = vs
        t = perm[t]
        del perm, order, keys, ids
    t = t.contiguous() if long_index else t.to(torch.int32).contiguous()   # int32 bit pattern == uint32 (ids < 2^31)
    return v.contiguous(), t
