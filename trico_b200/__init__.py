"""trico_b200 - B200-native encode/decode hot path of the trico mesh-compression library.

The product is ``trico_b200/lib/libtrico_b200.so``: host C (csrc/archive.c) exporting the
reference's C API plus a device-level C ABI (include/trico_b200_device.h), over hand-written
sm_100a CUDA kernels (csrc/*.cuh).  This Python package is only a loader and a thin ctypes mirror
used by the tests and bench.py; it contains no codec logic and no CPU fallback - if the library
cannot be loaded, or no CUDA device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB as LIB_PATH
from .build import build as build_library

__all__ = ["LIB_PATH", "build_library", "load", "Device", "STREAM_DTYPES", "TB200Error"]

_vp = C.c_void_p
_lib = None


class TB200Error(RuntimeError):
    pass


# scalar dtype per stream type (trico/trico.h:11-34)
STREAM_DTYPES = {1: np.float32, 2: np.float64, 3: np.uint32, 4: np.uint64, 5: np.float32, 6: np.float64,
                 7: np.float32, 8: np.float64, 9: np.float32, 10: np.float64, 11: np.float32, 12: np.float64,
                 13: np.uint32, 14: np.uint32, 15: np.float32, 16: np.float64, 17: np.uint8, 18: np.uint16,
                 19: np.uint32, 20: np.uint64}


def load() -> C.CDLL:
    """Load libtrico_b200.so (building it in-tree first if it is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build_library()
    L = C.CDLL(LIB_PATH)
    sig = {
        "tb200_ctx_create": (_vp, [C.c_int, _vp]),
        "tb200_ctx_destroy": (None, [_vp]),
        "tb200_ctx_stream": (_vp, [_vp]),
        "tb200_ctx_sync": (C.c_int, [_vp]),
        "tb200_last_error": (C.c_char_p, []),
        "tb200_ctx_launch_count": (C.c_uint64, [_vp]),
        "tb200_stream_layout": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "tb200_v1_nchunks": (C.c_uint64, [C.c_int, C.c_uint32, C.c_int]),
        "tb200_v1_stream_bound": (C.c_uint64, [C.c_int, C.c_uint32, C.c_int]),
        "tb200_default_log2_chunk": (C.c_int, [C.c_int, C.c_uint32]),
        "tb200_fpc_encode": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_uint64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
        "tb200_fpc_decode": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, _vp]),
        "tb200_fpc_encode_v0": (C.c_int, [_vp, C.c_int, _vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, _vp, C.c_uint64, _vp]),
        "tb200_fpc_v0_bound": (C.c_uint64, [C.c_int, C.c_uint32]),
        "tb200_fpc_decode_v0": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_uint32, _vp, C.c_uint32]),
        "tb200_lz4_encode": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, C.c_int, _vp, _vp, _vp, _vp]),
        "tb200_lz4_decode": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_uint64, C.c_uint64, C.c_int, _vp]),
        "tb200_lz4_decode_v0": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_uint64, _vp]),
        "tb200_lz4_v0_bound": (C.c_uint64, [C.c_uint64]),
        "tb200_lz4_encode_v0": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, _vp, C.c_uint64, _vp]),
        "tb200_encode_stream": (C.c_int, [_vp, C.c_int, _vp, C.c_uint32, C.c_int, _vp, C.c_uint64, _vp]),
        "tb200_decode_stream": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp]),
        "tb200_deinterleave": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_uint64, _vp]),
        "tb200_interleave": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_uint64, _vp]),
        "tb200_device_alloc": (_vp, [C.c_uint64]),
        "tb200_device_free": (None, [_vp]),
        "tb200_host_alloc_pinned": (_vp, [C.c_uint64]),
        "tb200_host_free_pinned": (None, [_vp]),
        "tb200_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
        "tb200_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
        "tb200_device_count": (C.c_int, []),
        "tb200_pointer_is_device": (C.c_int, [_vp]),
        "tb200_event_create": (_vp, []),
        "tb200_event_destroy": (None, [_vp]),
        "tb200_event_record": (C.c_int, [_vp, _vp]),
        "tb200_event_elapsed_ms": (C.c_float, [_vp, _vp]),
        "trico_b200_last_error": (C.c_char_p, []),
        "tb200_batch_arena_bytes": (C.c_uint64, [C.c_int, _vp, _vp]),
        "tb200_encode_streams": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp, _vp]),
        "tb200_decode_streams": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
        "tb200_comm_unique_id": (C.c_int, [_vp]),
        "tb200_comm_create": (_vp, [_vp, C.c_int, C.c_int, _vp]),
        "tb200_comm_destroy": (None, [_vp]),
        "tb200_shard_range": (C.c_int, [C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
        "tb200_encode_stream_sharded": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, _vp, C.c_uint64, _vp, C.POINTER(C.c_float)]),
        "tb200_comm_local_share": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_uint64), C.POINTER(_vp), C.POINTER(C.c_uint64)]),
        "tb200_decode_stream_range": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp]),
        # mesh front-end (include/trico_b200_io.h)
        "tb200_stl_dedup": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint32)]),
        "tb200_stl_dedup_scratch_bytes": (C.c_uint64, [C.c_uint32]),
        "tb200_stl_last_sort_passes": (C.c_int, []),
        "tb200_triangle_normals": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp]),
        "tb200_stl_facets": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp]),
        "trico_write_stl": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp, C.c_char_p]),
        "trico_b200_triangle_normals": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, _vp]),
        "trico_read_stl": (C.c_int, [C.POINTER(C.c_uint32), C.POINTER(_vp), C.POINTER(C.c_uint32), C.POINTER(_vp), C.c_char_p]),
        "trico_read_stl_full": (C.c_int, [C.POINTER(C.c_uint32), C.POINTER(_vp), C.POINTER(C.c_uint32), C.POINTER(_vp),
                                          C.POINTER(_vp), C.POINTER(_vp), C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class DeviceBuffer:
    """A cudaMalloc'd buffer owned by Python."""

    def __init__(self, lib, nbytes: int):
        self._lib = lib
        self.nbytes = int(nbytes)
        self.ptr = lib.tb200_device_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise TB200Error(lib.tb200_last_error().decode())

    def free(self):
        if self.ptr:
            self._lib.tb200_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Device:
    """ctypes mirror of the device-level C ABI on one GPU (one context = one CUDA stream)."""

    def __init__(self, index: int = 0, cuda_stream: int | None = None):
        self.lib = load()
        self.ctx = self.lib.tb200_ctx_create(index, _vp(cuda_stream) if cuda_stream else None)
        if not self.ctx:
            raise TB200Error("no usable CUDA device (trico_b200 has no CPU fallback): " + self.lib.tb200_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.tb200_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------
    def _ck(self, ok):
        if not ok:
            raise TB200Error(self.lib.tb200_last_error().decode())

    def sync(self):
        self._ck(self.lib.tb200_ctx_sync(self.ctx))

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self.lib, nbytes)

    def upload(self, arr) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr)
        buf = self.alloc(arr.nbytes + 64)
        if arr.nbytes:
            self._ck(self.lib.tb200_memcpy_h2d(self.ctx, buf.ptr, arr.ctypes.data_as(_vp), arr.nbytes))
            self.sync()
        return buf

    def download(self, ptr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, np.uint8)
        if nbytes:
            self._ck(self.lib.tb200_memcpy_d2h(self.ctx, out.ctypes.data_as(_vp), _vp(ptr), nbytes))
            self.sync()
        return out

    @property
    def launches(self) -> int:
        return self.lib.tb200_ctx_launch_count(self.ctx)

    def layout(self, stream_type: int):
        w, nc, pc = C.c_int(0), C.c_int(0), C.c_int(0)
        codec = self.lib.tb200_stream_layout(stream_type, C.byref(w), C.byref(nc), C.byref(pc))
        return dict(codec=codec, wordsize=w.value, ncomp=nc.value, per_count=pc.value)

    # -- mesh front-end (include/trico_b200_io.h; reference trico_io/iostl.c:70-138) -------
    def stl_dedup_device(self, d_facets: int, ntriangles: int, d_vertices: int, d_triangles: int, d_normals: int = 0, d_attrs: int = 0) -> int:
        nv = C.c_uint32(0)
        self._ck(self.lib.tb200_stl_dedup(self.ctx, _vp(d_facets), ntriangles, _vp(d_vertices), _vp(d_triangles),
                                          _vp(d_normals) if d_normals else None, _vp(d_attrs) if d_attrs else None, C.byref(nv)))
        return nv.value

    def stl_dedup(self, facets, full: bool = False):
        """host facet records (ntriangles x 50 bytes) -> (vertices [nv,3] f32, triangles [nt,3] u32[, normals, attributes])"""
        facets = np.ascontiguousarray(facets, dtype=np.uint8).reshape(-1, 50)
        nt = facets.shape[0]
        if nt == 0:
            empty = (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
            return empty + ((np.zeros((0, 3), np.float32), np.zeros(0, np.uint16)) if full else ())
        d_f, d_v, d_t = self.upload(facets), self.alloc(nt * 36), self.alloc(nt * 12)
        d_n, d_a = (self.alloc(nt * 12), self.alloc(nt * 2)) if full else (None, None)
        nv = self.stl_dedup_device(d_f.ptr, nt, d_v.ptr, d_t.ptr, d_n.ptr if full else 0, d_a.ptr if full else 0)
        out = (self.download(d_v.ptr, nv * 12).view(np.float32).reshape(-1, 3), self.download(d_t.ptr, nt * 12).view(np.uint32).reshape(-1, 3))
        if full:
            out += (self.download(d_n.ptr, nt * 12).view(np.float32).reshape(-1, 3), self.download(d_a.ptr, nt * 2).view(np.uint16))
        return out

    def triangle_normals(self, vertices, triangles) -> np.ndarray:
        """tools/trico_decoder/main.c:439-470 on the device, bit for bit"""
        vertices = np.ascontiguousarray(vertices, dtype=np.float32)
        triangles = np.ascontiguousarray(triangles, dtype=np.uint32)
        nt = triangles.size // 3
        if nt == 0:
            return np.zeros((0, 3), np.float32)
        d_v, d_t, d_n = self.upload(vertices), self.upload(triangles), self.alloc(nt * 12)
        self._ck(self.lib.tb200_triangle_normals(self.ctx, _vp(d_v.ptr), _vp(d_t.ptr), nt, _vp(d_n.ptr)))
        self.sync()
        return self.download(d_n.ptr, nt * 12).view(np.float32).reshape(-1, 3)

    # -- whole v1 streams (device resident) ------------------------------------------------
    def encode_stream_device(self, stream_type: int, d_data: int, count: int, d_out: int, out_cap: int, d_bytes: int, log2_chunk: int = 0):
        self._ck(self.lib.tb200_encode_stream(self.ctx, stream_type, _vp(d_data), count, log2_chunk, _vp(d_out), out_cap, _vp(d_bytes)))

    def decode_stream_device(self, header: bytes, d_stream: int, stream_bytes: int, d_out: int):
        hb = (C.c_uint8 * 15).from_buffer_copy(header[:15])
        self._ck(self.lib.tb200_decode_stream(self.ctx, hb, _vp(d_stream), stream_bytes, _vp(d_out)))

    def lz4_encode_v0(self, data) -> list:
        """host array of 1/2/4/8-byte integers -> one reference-format LZ4 block per byte plane (LSB first)"""
        data = np.ascontiguousarray(data).reshape(-1)
        w, n = data.dtype.itemsize, data.size
        stride = (self.lib.tb200_lz4_v0_bound(n) + 255) & ~255
        d_in, d_out, d_nb = self.upload(data), self.alloc(stride * w), self.alloc(64)
        self._ck(self.lib.tb200_lz4_encode_v0(self.ctx, w, _vp(d_in.ptr), n, _vp(d_out.ptr), stride, _vp(d_nb.ptr)))
        self.sync()
        nb = self.download(d_nb.ptr, 8 * w).view(np.uint64)
        return [self.download(d_out.ptr + p * stride, int(nb[p])).tobytes() for p in range(w)]

    def encode_stream(self, stream_type: int, data, count: int, log2_chunk: int = 0) -> bytes:
        """host array -> v1 stream bytes (type byte first)."""
        data = np.ascontiguousarray(data, dtype=STREAM_DTYPES[stream_type])
        if log2_chunk <= 0:
            log2_chunk = self.lib.tb200_default_log2_chunk(stream_type, count)
        bound = self.lib.tb200_v1_stream_bound(stream_type, count, log2_chunk)
        d_in = self.upload(data)
        d_out = self.alloc(bound)
        d_sz = self.alloc(64)
        self.encode_stream_device(stream_type, d_in.ptr, count, d_out.ptr, bound, d_sz.ptr, log2_chunk)
        nbytes = int(self.download(d_sz.ptr, 8).view(np.uint64)[0])
        out = self.download(d_out.ptr, nbytes).tobytes()
        for b in (d_in, d_out, d_sz):
            b.free()
        return out

    # -- batches of streams (many small meshes) --------------------------------------------------
    class Batch:
        """host-side description of a batch: arrays the C entry points read (kept alive here)"""

        def __init__(self, types, ptrs, counts):
            self.n = len(types)
            self.types = (C.c_int * self.n)(*types)
            self.counts = (C.c_uint32 * self.n)(*counts)
            self.ptrs = (C.c_void_p * self.n)(*ptrs)

    def batch_arena_bytes(self, batch) -> int:
        return self.lib.tb200_batch_arena_bytes(batch.n, batch.types, batch.counts)

    def encode_streams(self, batch, d_arena: int, arena_cap: int, d_packed: int, packed_cap: int, d_sizes: int, d_prefix: int):
        """asynchronous; d_sizes: n u64, d_prefix: 2n + 2 u64 (offsets of the streams in d_packed, total, scratch)"""
        self._ck(self.lib.tb200_encode_streams(self.ctx, batch.n, batch.types, batch.ptrs, batch.counts, _vp(d_arena), arena_cap,
                                               _vp(d_packed), packed_cap, _vp(d_sizes), _vp(d_prefix)))

    def decode_streams(self, headers: bytes, d_packed: int, offsets, sizes, out_ptrs, d_status: int):
        n = len(offsets)
        hb = (C.c_uint8 * (15 * n)).from_buffer_copy(headers)
        off = (C.c_uint64 * n)(*offsets)
        sz = (C.c_uint64 * n)(*sizes)
        outs = (C.c_void_p * n)(*out_ptrs)
        self._ck(self.lib.tb200_decode_streams(self.ctx, n, hb, _vp(d_packed), off, sz, outs, _vp(d_status)))

    # -- multi-GPU: chunk-sharded streams (one process per GPU) -------------------------------
    def shard_range(self, stream_type: int, count: int, rank: int, world: int, log2_chunk: int = 0):
        """-> (first, n): the units of a stream of `count` units that belong to `rank`"""
        f, n = C.c_uint32(0), C.c_uint32(0)
        self._ck(self.lib.tb200_shard_range(stream_type, count, log2_chunk, rank, world, C.byref(f), C.byref(n)))
        return f.value, n.value

    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._ck(self.lib.tb200_comm_unique_id(buf))
        return bytes(buf)

    def comm_create(self, rank: int, world: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        comm = self.lib.tb200_comm_create(self.ctx, rank, world, buf)
        if not comm:
            raise TB200Error(self.lib.tb200_last_error().decode())
        return comm

    def comm_destroy(self, comm):
        self.lib.tb200_comm_destroy(comm)

    def encode_stream_sharded(self, comm, stream_type: int, d_local: int, count_local: int, count_total: int, root: int = 0,
                              assemble: bool = True, d_out: int = 0, out_cap: int = 0, d_bytes: int = 0, log2_chunk: int = 0, timed: bool = True):
        """collective; -> (ms_encode_and_exchange, ms_assemble) when timed"""
        ms = (C.c_float * 2)()
        self._ck(self.lib.tb200_encode_stream_sharded(self.ctx, comm, stream_type, _vp(d_local), count_local, count_total, log2_chunk, root,
                                                      1 if assemble else 0, _vp(d_out) if d_out else None, out_cap,
                                                      _vp(d_bytes) if d_bytes else None, ms if timed else None))
        return float(ms[0]), float(ms[1])

    def comm_local_share(self, comm):
        """-> (d_sizes, table_bytes, d_payload, payload_bytes) of this rank's share of the last sharded call"""
        ps, pp = _vp(), _vp()
        ts, pb = C.c_uint64(0), C.c_uint64(0)
        self._ck(self.lib.tb200_comm_local_share(comm, C.byref(ps), C.byref(ts), C.byref(pp), C.byref(pb)))
        return ps.value, ts.value, pp.value, pb.value

    def decode_stream_range(self, header: bytes, d_stream: int, stream_bytes: int, first: int, n: int, d_out: int):
        hb = (C.c_uint8 * 15).from_buffer_copy(header[:15])
        self._ck(self.lib.tb200_decode_stream_range(self.ctx, hb, _vp(d_stream), stream_bytes, first, n, _vp(d_out)))

    def decode_stream(self, stream: bytes) -> np.ndarray:
        """v1 stream bytes -> flat host array of the stream's scalars."""
        stream_type = stream[0]
        count = int.from_bytes(stream[1:5], "little")
        lay = self.layout(stream_type)
        nsc = count * lay["per_count"] * (lay["ncomp"] if lay["codec"] == 1 else 1)
        dtype = np.dtype(STREAM_DTYPES[stream_type])
        d_s = self.upload(np.frombuffer(stream, np.uint8))
        d_o = self.alloc(nsc * dtype.itemsize + 64)
        self.decode_stream_device(stream, d_s.ptr, len(stream), d_o.ptr)
        out = self.download(d_o.ptr, nsc * dtype.itemsize).view(dtype)
        d_s.free()
        d_o.free()
        return out


def read_stl(filename: str, full: bool = False):
    """trico_read_stl / trico_read_stl_full of the B200 library (reference trico_io/iostl.c:141, :197):
    binary STL file -> (vertices [nv,3] f32, triangles [nt,3] u32[, normals [nt,3] f32, attributes [nt] u16])."""
    L = load()
    libc = C.CDLL(None)
    libc.free.argtypes = [_vp]
    nv, nt = C.c_uint32(0), C.c_uint32(0)
    pv, pt, pn, pa = _vp(), _vp(), _vp(), _vp()
    if full:
        ok = L.trico_read_stl_full(C.byref(nv), C.byref(pv), C.byref(nt), C.byref(pt), C.byref(pn), C.byref(pa), filename.encode())
    else:
        ok = L.trico_read_stl(C.byref(nv), C.byref(pv), C.byref(nt), C.byref(pt), filename.encode())
    if not ok:
        raise TB200Error(L.tb200_last_error().decode() or "trico_read_stl failed")

    def take(p, count, dtype):
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,)).copy().view(dtype) \
            if count and p.value else np.zeros(0, dtype)
        if p.value:
            libc.free(p)
        return a
    out = (take(pv, nv.value * 3, np.float32).reshape(-1, 3), take(pt, nt.value * 3, np.uint32).reshape(-1, 3))
    if full:
        out += (take(pn, nt.value * 3, np.float32).reshape(-1, 3), take(pa, nt.value, np.uint16))
    return out
