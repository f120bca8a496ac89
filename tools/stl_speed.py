#!/usr/bin/env python
"""STL front-end speed: vertex de-duplication of a binary STL on the GPU (device resident,
wall clock around tb200_stl_dedup, which allocates its scratch and synchronises) next to the reference's
trico_read_stl (oracle/_ref, one host thread, file in the page cache) on the same file.

    python tools/stl_speed.py [million triangles, default 10]
"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import trico_b200
from checkers import REF_SO, c_read_stl, have_ref, stl_facets, stl_file_bytes
from stl_cases import grid_mesh

mt = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
side = int((mt * 1e6 / 2) ** 0.5) + 1
v, t = grid_mesh(side, side, 1)
facets = stl_facets(v, t)
nt = t.shape[0]
dev = trico_b200.Device(0)
d_f, d_v, d_t = dev.upload(facets), dev.alloc(nt * 36), dev.alloc(nt * 12)
best = 1e9
for it in range(4):
    t0 = time.perf_counter()
    nv = dev.stl_dedup_device(d_f.ptr, nt, d_v.ptr, d_t.ptr)
    dt = time.perf_counter() - t0
    if it:
        best = min(best, dt)
passes = dev.lib.tb200_stl_last_sort_passes()
assert nv == v.shape[0], (nv, v.shape)
print(f"GPU de-dup (device resident): {nt} triangles ({nt * 50 / 1e6:.1f} MB of facets) -> {nv} vertices in {best * 1e3:.2f} ms "
      f"= {nt / best / 1e6:.1f} M triangles/s, {nt * 50 / best / 1e9:.2f} GB/s of STL; {passes} of 12 sort passes ran; "
      f"scratch {dev.lib.tb200_stl_dedup_scratch_bytes(nt) / 1e6:.0f} MB")
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "m.stl")
    open(path, "wb").write(stl_file_bytes(facets))
    t0 = time.perf_counter()
    gv, gt = trico_b200.read_stl(path)
    dt = time.perf_counter() - t0
    print(f"trico_read_stl (B200 library, file -> host arrays, context + buffers created per call): {dt * 1e3:.1f} ms = {nt / dt / 1e6:.1f} M triangles/s")
    if have_ref():
        t0 = time.perf_counter()
        rv, rt = c_read_stl(REF_SO, path)
        dr = time.perf_counter() - t0
        same = rv.tobytes() == gv.tobytes() and rt.tobytes() == gt.tobytes()
        print(f"trico_read_stl (reference, one thread): {dr * 1e3:.1f} ms = {nt / dr / 1e6:.2f} M triangles/s; identical mesh: {same}")
