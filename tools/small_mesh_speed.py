#!/usr/bin/env python
"""C5-shaped batch through the drop-in C API: many small meshes, one archive each, host buffers.
Reports meshes/s and GB/s of uncompressed bytes for encode and decode (wall clock).

    python tools/small_mesh_speed.py [number of meshes] [host threads]
"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trico_b200
from trico_b200.synth import grid_mesh

nm = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 1           # host threads, one archive each at a time (ctypes calls release the GIL)
from concurrent.futures import ThreadPoolExecutor
pool = ThreadPoolExecutor(nthreads)
L = C.CDLL(trico_b200.LIB_PATH)
L.trico_open_archive_for_writing.restype = C.c_void_p
L.trico_open_archive_for_reading.restype = C.c_void_p
L.trico_get_buffer_pointer.restype = C.c_void_p
L.trico_get_size.restype = C.c_uint64
for f in ("trico_write_vertices", "trico_write_triangles", "trico_write_attributes_float", "trico_write_attributes_uint8",
          "trico_write_attributes_uint16", "trico_write_attributes_uint64", "trico_close_archive", "trico_get_size", "trico_get_buffer_pointer",
          "trico_read_vertices", "trico_read_triangles", "trico_read_attributes_float", "trico_read_attributes_uint8",
          "trico_read_attributes_uint16", "trico_read_attributes_uint64"):
    getattr(L, f).argtypes = None
meshes = []
raw = 0
for m in range(nm):
    side = 32 + 8 * (m % 64)
    v, t = grid_mesh(side, side, jitter=1.0, seed=m)
    nv = v.shape[0]
    ix, iy = np.arange(nv) % side, np.arange(nv) // side
    att = dict(f=(0.1 * v[:, 2]).astype(np.float32), u8=(((ix >> 4) + (iy >> 4)) & 255).astype(np.uint8),
               u16=np.clip((v[:, 2] + 5.5) * 5000, 0, 65535).astype(np.uint16), u64=(np.uint64(m) << np.uint64(32)) | np.arange(nv, dtype=np.uint64))
    meshes.append((v, t, att))
    raw += v.nbytes + t.nbytes + sum(a.nbytes for a in att.values())
vp = lambda a: a.ctypes.data_as(C.c_void_p)
def encode(v, t, att):
    a = C.c_void_p(L.trico_open_archive_for_writing(1 << 16))
    assert L.trico_write_vertices(a, vp(v), C.c_uint32(v.shape[0])) == 1
    assert L.trico_write_triangles(a, vp(t), C.c_uint32(t.shape[0])) == 1
    assert L.trico_write_attributes_float(a, vp(att["f"]), C.c_uint32(att["f"].size)) == 1
    assert L.trico_write_attributes_uint8(a, vp(att["u8"]), C.c_uint32(att["u8"].size)) == 1
    assert L.trico_write_attributes_uint16(a, vp(att["u16"]), C.c_uint32(att["u16"].size)) == 1
    assert L.trico_write_attributes_uint64(a, vp(att["u64"]), C.c_uint32(att["u64"].size)) == 1
    n = L.trico_get_size(a)
    blob = C.string_at(L.trico_get_buffer_pointer(a), n)
    L.trico_close_archive(a)
    return blob
encode(*meshes[0])
t0 = time.perf_counter()
blobs = list(pool.map(lambda m: encode(*m), meshes))
te = time.perf_counter() - t0
def decode(blob, v, t, att):
    r = C.c_void_p(L.trico_open_archive_for_reading(blob, C.c_uint64(len(blob))))
    ov, ot = np.empty_like(v), np.empty_like(t)
    p = C.c_void_p(ov.ctypes.data); assert L.trico_read_vertices(r, C.byref(p)) == 1
    p = C.c_void_p(ot.ctypes.data); assert L.trico_read_triangles(r, C.byref(p)) == 1
    outs = {}
    for k, fn in (("f", "trico_read_attributes_float"), ("u8", "trico_read_attributes_uint8"), ("u16", "trico_read_attributes_uint16"), ("u64", "trico_read_attributes_uint64")):
        o = np.empty_like(att[k]); p = C.c_void_p(o.ctypes.data)
        assert getattr(L, fn)(r, C.byref(p)) == 1
        outs[k] = o if k != "f" else np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=att[k].shape).copy()
    L.trico_close_archive(r)
    return ov, ot, outs
decode(blobs[0], *meshes[0])
t0 = time.perf_counter()
outs = list(pool.map(lambda bm: decode(bm[0], *bm[1]), zip(blobs, meshes)))
td = time.perf_counter() - t0
for (ov, ot, oa), (v, t, att) in zip(outs, meshes):
    assert ov.tobytes() == v.tobytes() and ot.tobytes() == t.tobytes() and all(oa[k].tobytes() == att[k].tobytes() for k in att)
arch = sum(len(b) for b in blobs)
print(f"{nm} meshes, {nthreads} host threads, {raw / 1e6:.1f} MB raw, ratio {raw / arch:.3f}: encode {nm / te:.0f} meshes/s {raw / te / 1e9:.2f} GB/s ({te / nm / 6 * 1e6:.0f} us per stream), decode {nm / td:.0f} meshes/s {raw / td / 1e9:.2f} GB/s ({td / nm / 6 * 1e6:.0f} us per stream)")

# ---- the same meshes, device resident, through the BATCHED entry points (tb200_encode_streams /
# tb200_decode_streams): one call for all 6 * nm streams, no per-stream synchronisation ----
dev = trico_b200.Device(0)
order = [("v", 1), ("t", 3), ("f", 15), ("u8", 17), ("u16", 18), ("u64", 20)]
types, arrays, counts = [], [], []
for v, t, att in meshes:
    src = dict(v=v, t=t, **att)
    for key, ty in order:
        a = np.ascontiguousarray(src[key])
        types.append(ty); arrays.append(a); counts.append(a.shape[0])
bufs = [dev.upload(a) for a in arrays]
n = len(types)
batch = dev.Batch(types, [b.ptr for b in bufs], counts)
cap = dev.batch_arena_bytes(batch)
arena, packed = dev.alloc(cap), dev.alloc(cap)
d_sizes, d_prefix, d_status = dev.alloc(8 * n), dev.alloc(8 * (2 * n + 2)), dev.alloc(64)
def enc_batch():
    dev.encode_streams(batch, arena.ptr, cap, packed.ptr, cap, d_sizes.ptr, d_prefix.ptr)
enc_batch(); dev.sync()
h_sizes = dev.download(d_sizes.ptr, 8 * n).view(np.uint64).copy()
h_pref = dev.download(d_prefix.ptr, 8 * (n + 1)).view(np.uint64).copy()
total = int(h_pref[n])
host_packed = dev.download(packed.ptr, total)
# byte-identical to the per-mesh archives: an archive is its 8-byte header + its six streams
for m in range(nm):
    lo, hi = int(h_pref[6 * m]), int(h_pref[6 * m + 6]) if 6 * m + 6 <= n else total
    assert host_packed[lo:hi].tobytes() == blobs[m][8:], f"batched streams of mesh {m} differ from its archive"
headers = b"".join(host_packed[int(o):int(o) + 15].tobytes() for o in h_pref[:n])
outs_d = [dev.alloc(a.nbytes + 64) for a in arrays]
def dec_batch():
    dev.decode_streams(headers, packed.ptr, [int(x) for x in h_pref[:n]], [int(x) for x in h_sizes], [o.ptr for o in outs_d], d_status.ptr)
dec_batch(); dev.sync()
assert int(dev.download(d_status.ptr, 4).view(np.uint32)[0]) == 0
for a, o in list(zip(arrays, outs_d))[:60]:
    assert dev.download(o.ptr, a.nbytes).tobytes() == a.tobytes()
res = []
for f in (enc_batch, dec_batch):
    t0 = time.perf_counter()
    for _ in range(3): f()
    dev.sync()
    res.append(3 * raw / (time.perf_counter() - t0) / 1e9)
print(f"batched entry points, device resident, {n} streams per call: encode {res[0]:.1f} GB/s, decode {res[1]:.1f} GB/s; the packed streams are byte-identical to the {nm} per-mesh archives")
