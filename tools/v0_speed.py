#!/usr/bin/env python
"""Speed of the reference-format (v0) encoders on the device: the three component chains of a float
vec3 array (K3L, tile-parallel; TB200_FPC_V0_TILED=0: one warp per chain) and the four whole-plane LZ4
blocks of its triangle indices.  tools/v0_speed.py [millions of vertices]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trico_b200
from trico_b200.synth import grid_mesh

M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
side = int((M * 1e6) ** 0.5)
v, tri = grid_mesh(side, side, jitter=1.0, seed=5)
nv = v.shape[0]
dev = trico_b200.Device(0)
L = dev.lib
bound = L.tb200_fpc_v0_bound(4, nv)
d_in, d_out, d_nb = dev.upload(v), dev.alloc(3 * bound), dev.alloc(64)
def enc():
    assert L.tb200_fpc_encode_v0(dev.ctx, 4, d_in.ptr, nv, 3, 3, 4, 10, d_out.ptr, bound, d_nb.ptr)
enc(); dev.sync()
nb = dev.download(d_nb.ptr, 12).view(np.uint32)
t0 = time.perf_counter()
reps = 3
for _ in range(reps): enc()
dev.sync()
dt = (time.perf_counter() - t0) / reps
mode = "one warp per chain" if os.environ.get("TB200_FPC_V0_TILED") == "0" else "tile-parallel"
print(f"v0 FPC encode ({mode}): {nv} vertices, {v.nbytes / 1e6:.1f} MB -> {int(nb.sum())} B (ratio {v.nbytes / nb.sum():.3f}) in {dt * 1e3:.2f} ms = {v.nbytes / dt / 1e9:.1f} GB/s")

# whole-plane LZ4 blocks of the index stream
t = np.ascontiguousarray(tri.reshape(-1), dtype=np.uint32)
stride = (L.tb200_lz4_v0_bound(t.size) + 255) & ~255
d_t, d_o, d_n = dev.upload(t), dev.alloc(stride * 4), dev.alloc(64)
import ctypes as C
def enc4():
    assert L.tb200_lz4_encode_v0(dev.ctx, 4, d_t.ptr, t.size, d_o.ptr, stride, d_n.ptr)
enc4(); dev.sync()
nb4 = dev.download(d_n.ptr, 32).view(np.uint64)
t0 = time.perf_counter()
for _ in range(reps): enc4()
dev.sync()
dt = (time.perf_counter() - t0) / reps
print(f"v0 LZ4 planes: {tri.shape[0]} triangles, {t.nbytes / 1e6:.1f} MB -> {int(nb4.sum())} B (ratio {t.nbytes / nb4.sum():.3f}) in {dt * 1e3:.2f} ms = {t.nbytes / dt / 1e9:.1f} GB/s")
