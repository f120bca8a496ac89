#!/usr/bin/env python
"""Speed of the reference-format (v0) FPC encoder on the device: three component chains of a float
vec3 array, tile-parallel (K3L) against the one-warp-per-chain kernel.  tools/v0_speed.py [millions of vertices]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trico_b200
from trico_b200.synth import grid_mesh

M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
side = int((M * 1e6) ** 0.5)
v, _ = grid_mesh(side, side, jitter=1.0, seed=5)
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
