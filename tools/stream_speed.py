#!/usr/bin/env python
"""Device-resident encode/decode GB/s of single v1 streams of the other stream types (C3/C4/C5
shapes: doubles, uv, u64 indices, colours, u8/u16 attribute lists).  Not a bench line: a sanity
check that every kernel instantiation runs at a sensible rate.

    python tools/stream_speed.py [millions of elements]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trico_b200
from trico_b200.synth import grid_mesh

M = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
side = int((M * 1e6) ** 0.5)
v, t = grid_mesh(side, side, jitter=1.0, seed=5)
nv, nt = v.shape[0], t.shape[0]
dev = trico_b200.Device(0)
rng = np.random.default_rng(1)
# SURVEY.md 8(d) generators for the colour / attribute streams of C4 and C5
px, py, pz = v[:, 0].astype(np.float64), v[:, 1].astype(np.float64), v[:, 2].astype(np.float64)
noise = lambda: rng.integers(-4, 5, nv)
col = (np.clip(128 + 100 * np.sin(0.5 * px) + noise(), 0, 255).astype(np.uint32)
       | (np.clip(128 + 100 * np.sin(0.5 * py) + noise(), 0, 255).astype(np.uint32) << 8)
       | (np.clip(128 + 20 * pz + noise(), 0, 255).astype(np.uint32) << 16) | (np.uint32(255) << 24))
ix, iy = np.arange(nv) % side, np.arange(nv) // side
cases = [
    ("vec3 float (type 1)", 1, v, nv),
    ("vec3 double (type 2)", 2, v.astype(np.float64), nv),
    ("triangles u32 (type 3)", 3, t, nt),
    ("triangles u64 (type 4)", 4, t.astype(np.uint64), nt),
    ("uv float (type 5)", 5, np.ascontiguousarray(v[:, :2]), nv),
    ("uv double (type 6)", 6, np.ascontiguousarray(v[:, :2]).astype(np.float64), nv),
    ("colours u32 (type 13)", 13, col, nv),
    ("attr float (type 15)", 15, np.ascontiguousarray(v[:, 2]), nv),
    ("attr double (type 16)", 16, v[:, 2].astype(np.float64), nv),
    ("attr u8 (type 17)", 17, (((ix >> 4) + (iy >> 4)) & 255).astype(np.uint8), nv),
    ("attr u16 (type 18)", 18, np.clip((pz + 5.5) * 5000, 0, 65535).astype(np.uint16), nv),
    ("attr u64 (type 20)", 20, np.arange(nv, dtype=np.uint64) | (np.uint64(7) << 32), nv),
]
only = os.environ.get("ONLY")
for name, ty, data, cnt in cases:
    if only and str(ty) not in only.split(","): continue
    data = np.ascontiguousarray(data, dtype=trico_b200.STREAM_DTYPES[ty])
    log2c = dev.lib.tb200_default_log2_chunk(ty, cnt)
    bound = dev.lib.tb200_v1_stream_bound(ty, cnt, log2c)
    d_in, d_out, d_sz, d_back = dev.upload(data), dev.alloc(bound), dev.alloc(64), dev.alloc(data.nbytes + 64)
    def enc(): dev.encode_stream_device(ty, d_in.ptr, cnt, d_out.ptr, bound, d_sz.ptr, log2c)
    if only: print("encode", name, flush=True)
    enc(); dev.sync()
    if only: print("encoded", flush=True)
    nbytes = int(dev.download(d_sz.ptr, 8).view(np.uint64)[0])
    hdr = dev.download(d_out.ptr, 16).tobytes()
    def dec(): dev.decode_stream_device(hdr, d_out.ptr, nbytes, d_back.ptr)
    dec(); dev.sync()
    if only: print("decoded", nbytes, flush=True)
    if not os.environ.get("NOVERIFY"): assert dev.download(d_back.ptr, data.nbytes).tobytes() == data.tobytes(), name
    res = []
    for f in (enc, dec):
        f(); dev.sync()
        t0 = time.perf_counter()
        for _ in range(5): f()
        dev.sync()
        res.append(5 * data.nbytes / (time.perf_counter() - t0) / 1e9)
    print(f"{name:24s} {data.nbytes / 1e6:8.1f} MB  ratio {data.nbytes / nbytes:6.3f}  encode {res[0]:7.1f} GB/s  decode {res[1]:7.1f} GB/s")
    for b in (d_in, d_out, d_sz, d_back): b.free()
