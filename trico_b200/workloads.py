"""Synthetic workloads of the five BASELINE.json configurations, generated on the device with torch
(SURVEY.md 8d).  Data only - no codec logic lives here.  A workload is a list of streams
``(name, stream_type, tensor, count)``: `tensor` holds the stream's scalars (flat or [n, k]), `count`
is the value the archive stores for it (trico.c:221: vertices, triangles, list entries).
"""
from __future__ import annotations

import math
import os

import numpy as np

from .synth import grid_mesh_torch, _G, _M1, _M2

C2_GRID = (10000, 10001)          # 100,010,000 vertices / 199,980,000 triangles
C3_GRID = (7071, 7072)            # 50,006,112 vertices / 99,984,140 triangles
C4_SHARD_GRID = (11181, 11180)    # 125,003,580 points per GPU (8 shards = 1.0 B points)
C5_MESHES = 1024


def _splitmix_t(i):
    def lsr(z, s):
        return (z >> s) & ((1 << (64 - s)) - 1)
    z = i + (_G - (1 << 64))
    z = (z ^ lsr(z, 30)) * (_M1 - (1 << 64))
    z = (z ^ lsr(z, 27)) * (_M2 - (1 << 64))
    return z ^ lsr(z, 31)


def _u_t(i):
    import torch
    return ((_splitmix_t(i) >> 11) & ((1 << 53) - 1)).to(torch.float64) * 2.0 ** -53


def c2(device, seed=1, grid=C2_GRID):
    v, t = grid_mesh_torch(grid[0], grid[1], device, jitter=1.0, seed=seed)
    return [("vertices", 1, v, v.shape[0]), ("triangles", 3, t, t.shape[0])]


def c3(device, seed=2, grid=C3_GRID):
    """double positions, double unit normals of the height field, per-vertex uv as double, uint64 indices"""
    import torch
    v, t = grid_mesh_torch(grid[0], grid[1], device, jitter=1.0, seed=seed, double=True, long_index=True)
    px, py = v[:, 0], v[:, 1]
    dzdx = 5.0 * 0.37 * torch.cos(0.37 * px) * torch.cos(0.21 * py)
    dzdy = -5.0 * 0.21 * torch.sin(0.37 * px) * torch.sin(0.21 * py)
    inv = 1.0 / torch.sqrt(dzdx * dzdx + dzdy * dzdy + 1.0)
    nrm = torch.stack([-dzdx * inv, -dzdy * inv, inv], dim=1).contiguous()
    del dzdx, dzdy, inv
    uv = torch.stack([px / (0.01 * (grid[0] - 1)), py / (0.01 * (grid[1] - 1))], dim=1).contiguous()
    nv = v.shape[0]
    return [("vertices", 2, v, nv), ("normals", 10, nrm, nv), ("uv", 6, uv, nv), ("triangles", 4, t, t.shape[0])]


def colours(v, seed=0):
    """u32 RGBA per point, memory order r,g,b,a (trico_io/ioply.c:187-190); int32 bit pattern"""
    import torch
    n = v.shape[0]
    k = torch.arange(n, dtype=torch.int64, device=v.device) * 3 + (seed + 99) * 0x100000001B3
    d = v.to(torch.float64)
    r = 128 + 100 * torch.sin(0.5 * d[:, 0]) + 8 * (_u_t(k) - 0.5)
    g = 128 + 100 * torch.sin(0.5 * d[:, 1]) + 8 * (_u_t(k + 1) - 0.5)
    b = 128 + 20 * d[:, 2] + 8 * (_u_t(k + 2) - 0.5)
    r, g, b = (c.clamp(0, 255).to(torch.int64) for c in (r, g, b))
    col = r | (g << 8) | (b << 16) | (255 << 24)
    col = torch.where(col >= (1 << 31), col - (1 << 32), col).to(torch.int32)
    return col.contiguous()


def c4_shard(device, rank=0, grid=C4_SHARD_GRID):
    """one GPU's share of the 1 B-point cloud: float xyz + uint32 RGBA (no indices)"""
    v, _ = grid_mesh_torch(grid[0], grid[1], device, jitter=1.0, seed=40 + rank, shuffle=0, triangles=False)
    col = colours(v, seed=40 + rank)
    return [("points", 1, v, v.shape[0]), ("colours", 13, col, col.shape[0])]


def c5_mesh_sizes():
    return [32 + 8 * (m % 64) for m in range(C5_MESHES)]


def c5_assign(world):
    """whole meshes to GPUs by a size-balanced greedy: largest first to the least loaded"""
    sides = c5_mesh_sizes()
    order = sorted(range(C5_MESHES), key=lambda m: -sides[m])
    load = [0] * world
    owner = [0] * C5_MESHES
    for m in order:
        r = min(range(world), key=lambda q: load[q])
        owner[m] = r
        load[r] += sides[m] * sides[m]
    return owner


def c5_meshes(device, mesh_ids):
    """the meshes `mesh_ids` of the 1024-mesh batch: float vertices, uint32 triangles, and per-vertex
    float / uint8 / uint16 / uint64 attribute lists"""
    import torch
    out = []
    for m in mesh_ids:
        side = 32 + 8 * (m % 64)
        v, t = grid_mesh_torch(side, side, device, jitter=1.0, seed=1000 + m)
        nv = v.shape[0]
        i = torch.arange(nv, dtype=torch.int64, device=device)
        x, y = i % side, i // side
        pz = v[:, 2].to(torch.float64)
        fl = (0.1 * pz + 0.001 * (_u_t(i + m * 0x100000001B3) - 0.5)).to(torch.float32).contiguous()
        u8 = (((x >> 4) + (y >> 4)) & 255).to(torch.uint8).contiguous()
        u16 = ((pz + 5.5) * 5000).clamp(0, 65535).to(torch.int32).to(torch.int16).contiguous()      # bit pattern of the uint16 value
        u64 = ((m << 32) | i).contiguous()
        out.append([("vertices", 1, v, nv), ("triangles", 3, t, t.shape[0]), ("attr_float", 15, fl, nv),
                    ("attr_u8", 17, u8, nv), ("attr_u16", 18, u16, nv), ("attr_u64", 20, u64, nv)])
    return out


def bunny_tiled(device, min_triangles=100_000_000, root=None):
    """the bunny (tests/golden/bunny_full.npz, decoded from the reference's own archive) replicated
    with vertex offsets until it has `min_triangles` triangles: the index planes of a REAL mesh
    (many short sequences) at bench size"""
    import torch
    root = root or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "bunny_full.npz"))
    v0 = torch.from_numpy(z["vertices"].astype(np.float32)).to(device)
    t0 = torch.from_numpy(z["triangles"].astype(np.int64)).to(device)
    nv0, nt0 = v0.shape[0], t0.shape[0]
    tiles = max(1, math.ceil(min_triangles / nt0))
    k = torch.arange(tiles, dtype=torch.int64, device=device)
    t = (t0[None, :, :] + (k * nv0)[:, None, None]).reshape(-1, 3).to(torch.int32).contiguous()
    shift = torch.stack([(k % 64).to(torch.float32) * 0.2, (k // 64).to(torch.float32) * 0.2, torch.zeros_like(k, dtype=torch.float32)], dim=1)
    v = (v0[None, :, :] + shift[:, None, :]).reshape(-1, 3).contiguous()
    return [("vertices", 1, v, v.shape[0]), ("triangles", 3, t, t.shape[0])]
