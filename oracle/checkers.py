"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  - oracle/_ref/libtrico_oracle.so, our plain-C restatement (oracle/trico_oracle.c).
* ``Ref``     - oracle/_ref/libtrico_ref.so, the UNMODIFIED reference compiled from
                /root/reference by oracle/Makefile (present wherever that build ran; the .so
                travels to the GPU box, the sources do not).

Nothing in trico_b200/ imports this module; only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs do.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(ORACLE_DIR)
ORACLE_SO = os.path.join(ORACLE_DIR, "_ref", "libtrico_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libtrico_ref.so")

u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(vp)


def build_oracle():
    """Compile the restatement (and the reference when its sources are mounted)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


# numpy dtype for a (type enum) stream's scalars
def stream_dtype(stream_type: int):
    f32 = {1, 5, 7, 9, 11, 15}
    f64 = {2, 6, 8, 10, 12, 16}
    if stream_type in f32:
        return np.float32
    if stream_type in f64:
        return np.float64
    return {3: np.uint32, 4: np.uint64, 13: np.uint32, 14: np.uint32, 17: np.uint8, 18: np.uint16,
            19: np.uint32, 20: np.uint64}[stream_type]


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = self.lib = C.CDLL(ORACLE_SO)
        for name in ("oracle_fpc32_bound", "oracle_fpc64_bound", "oracle_lz4_bound"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_uint64]
        for name in ("oracle_fpc32_compress", "oracle_fpc64_compress"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]
        for name in ("oracle_fpc32_decompress", "oracle_fpc64_decompress"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [vp, vp]
        for name in ("oracle_fpc32_stream_bytes", "oracle_fpc64_stream_bytes"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [vp]
        L.oracle_planes_split.argtypes = [vp, vp, C.c_uint64, C.c_int]
        L.oracle_planes_merge.argtypes = [vp, vp, C.c_uint64, C.c_int]
        L.oracle_deinterleave.argtypes = [vp, vp, C.c_uint64, C.c_int, C.c_int]
        L.oracle_interleave.argtypes = [vp, vp, C.c_uint64, C.c_int, C.c_int]
        L.oracle_lz4_decompress.restype = C.c_int64
        L.oracle_lz4_decompress.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
        L.oracle_lz4_validate.restype = C.c_int64
        L.oracle_lz4_validate.argtypes = [vp, C.c_uint64, C.c_uint64]
        L.oracle_lz4_compress.restype = C.c_uint64
        L.oracle_lz4_compress.argtypes = [vp, vp, C.c_uint64]
        L.oracle_v0_write_header.restype = C.c_uint64
        L.oracle_v0_write_header.argtypes = [vp, C.c_uint32]
        L.oracle_v0_stream_bound.restype = C.c_uint64
        L.oracle_v0_stream_bound.argtypes = [C.c_int, C.c_uint32]
        L.oracle_v0_write_stream.restype = C.c_uint64
        L.oracle_v0_write_stream.argtypes = [vp, C.c_int, vp, C.c_uint32]
        L.oracle_v0_read_stream.restype = C.c_uint64
        L.oracle_v0_read_stream.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
        L.oracle_v1_stream_bound.restype = C.c_uint64
        L.oracle_v1_stream_bound.argtypes = [C.c_int, C.c_uint32, C.c_int]
        L.oracle_v1_write_stream.restype = C.c_uint64
        L.oracle_v1_write_stream.argtypes = [vp, C.c_int, vp, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.oracle_v1_read_stream.restype = C.c_uint64
        L.oracle_v1_read_stream.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]

    # ---- codecs -------------------------------------------------------------------------
    def fpc_compress(self, values: np.ndarray, e1: int, e2: int) -> bytes:
        values = np.ascontiguousarray(values)
        w = values.dtype.itemsize
        assert w in (4, 8)
        n = values.size
        bound = (self.lib.oracle_fpc32_bound if w == 4 else self.lib.oracle_fpc64_bound)(n)
        out = np.empty(bound, np.uint8)
        fn = self.lib.oracle_fpc32_compress if w == 4 else self.lib.oracle_fpc64_compress
        nb = fn(_ptr(out), _ptr(values), n, e1, e2)
        return out[:nb].tobytes()

    def fpc_decompress(self, stream: bytes, wordsize: int) -> np.ndarray:
        buf = np.frombuffer(stream, np.uint8)
        n = int.from_bytes(stream[1:5], "big")
        out = np.empty(n, np.uint32 if wordsize == 4 else np.uint64)
        fn = self.lib.oracle_fpc32_decompress if wordsize == 4 else self.lib.oracle_fpc64_decompress
        got = fn(_ptr(out), _ptr(buf))
        assert got == n
        return out

    def fpc_stream_bytes(self, stream: bytes, wordsize: int) -> int:
        buf = np.frombuffer(stream, np.uint8)
        fn = self.lib.oracle_fpc32_stream_bytes if wordsize == 4 else self.lib.oracle_fpc64_stream_bytes
        return fn(_ptr(buf))

    def planes_split(self, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a)
        out = np.empty((a.dtype.itemsize, a.size), np.uint8)
        self.lib.oracle_planes_split(_ptr(out), _ptr(a), a.size, a.dtype.itemsize)
        return out

    def planes_merge(self, planes: np.ndarray, dtype) -> np.ndarray:
        planes = np.ascontiguousarray(planes)
        w, n = planes.shape
        out = np.empty(n, dtype)
        assert out.dtype.itemsize == w
        self.lib.oracle_planes_merge(_ptr(out), _ptr(planes), n, w)
        return out

    def lz4_compress(self, raw: bytes) -> bytes:
        src = np.frombuffer(raw, np.uint8) if len(raw) else np.zeros(1, np.uint8)
        out = np.empty(self.lib.oracle_lz4_bound(len(raw)), np.uint8)
        nb = self.lib.oracle_lz4_compress(_ptr(out), _ptr(src), len(raw))
        return out[:nb].tobytes()

    def lz4_decompress(self, block: bytes, raw_len: int) -> bytes:
        src = np.frombuffer(block, np.uint8)
        out = np.empty(max(raw_len, 1), np.uint8)
        r = self.lib.oracle_lz4_decompress(_ptr(out), raw_len, _ptr(src), len(block))
        if r < 0:
            raise ValueError(f"malformed LZ4 block ({r})")
        return out[:r].tobytes()

    def lz4_validate(self, block: bytes, raw_len: int) -> int:
        src = np.frombuffer(block, np.uint8)
        return self.lib.oracle_lz4_validate(_ptr(src), len(block), raw_len)

    # ---- containers ---------------------------------------------------------------------
    def header(self, version: int) -> bytes:
        out = np.empty(8, np.uint8)
        self.lib.oracle_v0_write_header(_ptr(out), version)
        return out.tobytes()

    def v0_write_stream(self, stream_type: int, data: np.ndarray, count: int) -> bytes:
        data = np.ascontiguousarray(data)
        out = np.empty(self.lib.oracle_v0_stream_bound(stream_type, count) + 64, np.uint8)
        nb = self.lib.oracle_v0_write_stream(_ptr(out), stream_type, _ptr(data), count)
        return out[:nb].tobytes()

    def v1_write_stream(self, stream_type: int, data: np.ndarray, count: int, log2_chunk: int, e1=4, e2=4) -> bytes:
        data = np.ascontiguousarray(data)
        out = np.empty(self.lib.oracle_v1_stream_bound(stream_type, count, log2_chunk) + 64, np.uint8)
        nb = self.lib.oracle_v1_write_stream(_ptr(out), stream_type, _ptr(data), count, log2_chunk, e1, e2)
        return out[:nb].tobytes()

    def _read_stream(self, fn, blob: bytes, offset: int):
        buf = np.frombuffer(blob, np.uint8)
        t, cnt = C.c_int(0), C.c_uint32(0)
        base = buf.ctypes.data + offset
        used = fn(None, vp(base), len(blob) - offset, C.byref(t), C.byref(cnt))
        if used == 0:
            raise ValueError("cannot parse stream")
        lay = self.layout(t.value)
        nsc = cnt.value * lay["per_count"] * (lay["ncomp"] if lay["codec"] == 1 else 1)
        out = np.empty(nsc, stream_dtype(t.value))
        used = fn(_ptr(out), vp(base), len(blob) - offset, C.byref(t), C.byref(cnt))
        if used == 0:
            raise ValueError("cannot decode stream")
        return t.value, cnt.value, out, used

    def v0_read_stream(self, blob: bytes, offset: int):
        return self._read_stream(self.lib.oracle_v0_read_stream, blob, offset)

    def v1_read_stream(self, blob: bytes, offset: int):
        return self._read_stream(self.lib.oracle_v1_read_stream, blob, offset)

    def read_archive(self, blob: bytes):
        """Decode a whole archive (either version) -> (version, [(type, count, ndarray)])."""
        assert blob[:4] == b"Trco"
        version = int.from_bytes(blob[4:8], "little")
        off, out = 8, []
        while off < len(blob):
            t, cnt, arr, used = (self.v0_read_stream if version == 0 else self.v1_read_stream)(blob, off)
            out.append((t, cnt, arr))
            off += used
        return version, out

    def layout(self, stream_type: int):
        class Lay(C.Structure):
            _fields_ = [("codec", C.c_int), ("wordsize", C.c_int), ("ncomp", C.c_int), ("per_count", C.c_int)]
        lay = Lay()
        self.lib.oracle_stream_layout.argtypes = [C.c_int, C.POINTER(Lay)]
        ok = self.lib.oracle_stream_layout(stream_type, C.byref(lay))
        assert ok
        return dict(codec=lay.codec, wordsize=lay.wordsize, ncomp=lay.ncomp, per_count=lay.per_count)


# names of the reference's writer / reader / counter per stream type (trico/trico.h:40-93)
REF_API = {
    1: ("trico_write_vertices", "trico_read_vertices", "trico_get_number_of_vertices"),
    2: ("trico_write_vertices_double", "trico_read_vertices_double", "trico_get_number_of_vertices"),
    3: ("trico_write_triangles", "trico_read_triangles", "trico_get_number_of_triangles"),
    4: ("trico_write_triangles_long", "trico_read_triangles_long", "trico_get_number_of_triangles"),
    5: ("trico_write_uv_per_vertex", "trico_read_uv_per_vertex", "trico_get_number_of_uvs"),
    6: ("trico_write_uv_per_vertex_double", "trico_read_uv_per_vertex_double", "trico_get_number_of_uvs"),
    7: ("trico_write_uv_per_triangle", "trico_read_uv_per_triangle", "trico_get_number_of_uvs"),
    8: ("trico_write_uv_per_triangle_double", "trico_read_uv_per_triangle_double", "trico_get_number_of_uvs"),
    9: ("trico_write_vertex_normals", "trico_read_vertex_normals", "trico_get_number_of_normals"),
    10: ("trico_write_vertex_normals_double", "trico_read_vertex_normals_double", "trico_get_number_of_normals"),
    11: ("trico_write_triangle_normals", "trico_read_triangle_normals", "trico_get_number_of_normals"),
    12: ("trico_write_triangle_normals_double", "trico_read_triangle_normals_double", "trico_get_number_of_normals"),
    13: ("trico_write_vertex_colors", "trico_read_vertex_colors", "trico_get_number_of_colors"),
    14: ("trico_write_triangle_colors", "trico_read_triangle_colors", "trico_get_number_of_colors"),
    15: ("trico_write_attributes_float", "trico_read_attributes_float", "trico_get_number_of_attributes"),
    16: ("trico_write_attributes_double", "trico_read_attributes_double", "trico_get_number_of_attributes"),
    17: ("trico_write_attributes_uint8", "trico_read_attributes_uint8", "trico_get_number_of_attributes"),
    18: ("trico_write_attributes_uint16", "trico_read_attributes_uint16", "trico_get_number_of_attributes"),
    19: ("trico_write_attributes_uint32", "trico_read_attributes_uint32", "trico_get_number_of_attributes"),
    20: ("trico_write_attributes_uint64", "trico_read_attributes_uint64", "trico_get_number_of_attributes"),
}


class TricoCApi:
    """The reference's C API surface over any shared library exporting it (the compiled
    reference, or - in the parity tests - our own drop-in library)."""

    def __init__(self, path: str):
        L = self.lib = C.CDLL(path)
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [vp]
        L.trico_open_archive_for_writing.restype = vp
        L.trico_open_archive_for_writing.argtypes = [C.c_uint64]
        L.trico_open_archive_for_reading.restype = vp
        L.trico_open_archive_for_reading.argtypes = [vp, C.c_uint64]
        L.trico_close_archive.argtypes = [vp]
        L.trico_get_buffer_pointer.restype = vp
        L.trico_get_buffer_pointer.argtypes = [vp]
        L.trico_get_size.restype = C.c_uint64
        L.trico_get_size.argtypes = [vp]
        L.trico_get_version.restype = C.c_uint32
        L.trico_get_version.argtypes = [vp]
        L.trico_get_next_stream_type.restype = C.c_int
        L.trico_get_next_stream_type.argtypes = [vp]
        L.trico_skip_next_stream.restype = C.c_int
        L.trico_skip_next_stream.argtypes = [vp]
        for w, r, cfn in REF_API.values():
            getattr(L, w).restype = C.c_int
            getattr(L, w).argtypes = [vp, vp, C.c_uint32]
            getattr(L, r).restype = C.c_int
            getattr(L, r).argtypes = [vp, C.POINTER(vp)]
            getattr(L, cfn).restype = C.c_uint32
            getattr(L, cfn).argtypes = [vp]
        L.trico_compress.argtypes = [C.POINTER(C.c_uint32), C.POINTER(vp), vp, C.c_uint32, C.c_uint32, C.c_uint32]
        L.trico_decompress.argtypes = [C.POINTER(C.c_uint32), C.POINTER(vp), vp]
        L.trico_compress_double_precision.argtypes = [C.POINTER(C.c_uint32), C.POINTER(vp), vp, C.c_uint32, C.c_uint64, C.c_uint64]
        L.trico_decompress_double_precision.argtypes = [C.POINTER(C.c_uint32), C.POINTER(vp), vp]

    # ---- archive level ------------------------------------------------------------------
    def encode(self, streams, initial=1024) -> bytes:
        """streams: [(type, ndarray, count_argument)] -> archive bytes."""
        L = self.lib
        a = L.trico_open_archive_for_writing(initial)
        assert a
        try:
            for t, data, count in streams:
                data = np.ascontiguousarray(data, dtype=stream_dtype(t))
                ok = getattr(L, REF_API[t][0])(a, _ptr(data), count)
                if ok != 1:
                    raise RuntimeError(f"{REF_API[t][0]} failed")
            n = L.trico_get_size(a)
            return C.string_at(L.trico_get_buffer_pointer(a), n)
        finally:
            L.trico_close_archive(a)

    def decode(self, blob: bytes, oracle: "Oracle"):
        """-> (version, [(type, count, ndarray)]) through the read API."""
        L = self.lib
        buf = np.frombuffer(blob, np.uint8)
        a = L.trico_open_archive_for_reading(_ptr(buf), len(blob))
        if not a:
            raise ValueError("not a trico archive")
        out = []
        try:
            version = L.trico_get_version(a)
            while True:
                t = L.trico_get_next_stream_type(a)
                if t == 0:
                    break
                lay = oracle.layout(t)
                cnt = getattr(L, REF_API[t][2])(a)
                nsc = cnt * lay["per_count"] * (lay["ncomp"] if lay["codec"] == 1 else 1)
                arr = np.zeros(max(nsc, 1), stream_dtype(t))
                p = vp(arr.ctypes.data)
                ok = getattr(L, REF_API[t][1])(a, C.byref(p))
                if ok != 1:
                    raise RuntimeError(f"{REF_API[t][1]} failed")
                if p.value != arr.ctypes.data:
                    # float/double attribute readers hand back a malloc'd buffer (trico.c:1377,:1408)
                    got = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nsc * arr.dtype.itemsize,)).copy()
                    self.libc.free(p)
                    arr = got.view(arr.dtype)
                out.append((t, cnt, arr[:nsc]))
            return version, out
        finally:
            L.trico_close_archive(a)

    # ---- raw codec level ----------------------------------------------------------------
    def compress(self, values: np.ndarray, e1: int, e2: int) -> bytes:
        values = np.ascontiguousarray(values)
        nb, out = C.c_uint32(0), vp()
        if values.dtype.itemsize == 4:
            self.lib.trico_compress(C.byref(nb), C.byref(out), _ptr(values), values.size, e1, e2)
        else:
            self.lib.trico_compress_double_precision(C.byref(nb), C.byref(out), _ptr(values), values.size, e1, e2)
        data = C.string_at(out, nb.value)
        self.libc.free(out)
        return data

    def decompress(self, stream: bytes, wordsize: int) -> np.ndarray:
        buf = np.frombuffer(stream, np.uint8)
        n, out = C.c_uint32(0), vp()
        if wordsize == 4:
            self.lib.trico_decompress(C.byref(n), C.byref(out), _ptr(buf))
        else:
            self.lib.trico_decompress_double_precision(C.byref(n), C.byref(out), _ptr(buf))
        data = C.string_at(out, n.value * wordsize)
        self.libc.free(out)
        return np.frombuffer(data, np.uint32 if wordsize == 4 else np.uint64).copy()


class Ref(TricoCApi):
    """The compiled, unmodified reference."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        super().__init__(REF_SO)
        L = self.lib
        L.LZ4_compress_default.restype = C.c_int
        L.LZ4_compress_default.argtypes = [vp, vp, C.c_int, C.c_int]
        L.LZ4_decompress_safe.restype = C.c_int
        L.LZ4_decompress_safe.argtypes = [vp, vp, C.c_int, C.c_int]
        L.trico_read_stl.restype = C.c_int
        L.trico_read_stl.argtypes = [C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(vp), C.c_char_p]

    def lz4_compress(self, raw: bytes) -> bytes:
        src = np.frombuffer(raw, np.uint8) if len(raw) else np.zeros(1, np.uint8)
        cap = len(raw) + len(raw) // 255 + 16
        out = np.empty(cap, np.uint8)
        nb = self.lib.LZ4_compress_default(_ptr(src), _ptr(out), len(raw), cap)
        assert nb > 0
        return out[:nb].tobytes()

    def lz4_decompress(self, block: bytes, raw_len: int) -> bytes:
        src = np.frombuffer(block, np.uint8)
        out = np.empty(max(raw_len, 1), np.uint8)
        r = self.lib.LZ4_decompress_safe(_ptr(src), _ptr(out), len(block), raw_len)
        if r < 0:
            raise ValueError("LZ4_decompress_safe failed")
        return out[:r].tobytes()

    def read_stl(self, path: str):
        nv, nt, pv, pt = C.c_uint32(0), C.c_uint32(0), vp(), vp()
        ok = self.lib.trico_read_stl(C.byref(nv), C.byref(pv), C.byref(nt), C.byref(pt), path.encode())
        assert ok == 1
        v = np.frombuffer(C.string_at(pv, nv.value * 12), np.float32).reshape(-1, 3).copy()
        t = np.frombuffer(C.string_at(pt, nt.value * 12), np.uint32).reshape(-1, 3).copy()
        self.libc.free(pv)
        self.libc.free(pt)
        return v, t


def have_ref() -> bool:
    return os.path.exists(REF_SO)


# ---- mesh front-end: STL facets, de-duplication, normals --------------------------------

def stl_facets(vertices: np.ndarray, triangles: np.ndarray, normals=None, attrs=None) -> np.ndarray:
    """facet records of a binary STL file (iostl.c:261-320 layout): [nt, 50] bytes"""
    vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    triangles = np.ascontiguousarray(triangles, np.uint32).reshape(-1, 3)
    nt = triangles.shape[0]
    rec = np.zeros((nt, 50), np.uint8)
    if normals is not None:
        rec[:, 0:12] = np.ascontiguousarray(normals, np.float32).reshape(nt, 3).view(np.uint8).reshape(nt, 12)
    rec[:, 12:48] = vertices[triangles.reshape(-1)].reshape(nt, 9).view(np.uint8).reshape(nt, 36)
    if attrs is not None:
        rec[:, 48:50] = np.ascontiguousarray(attrs, np.uint16).reshape(nt, 1).view(np.uint8).reshape(nt, 2)
    return rec


def stl_file_bytes(facets: np.ndarray, header: bytes = b"binary stl written by the trico_b200 tests") -> bytes:
    facets = np.ascontiguousarray(facets, np.uint8).reshape(-1, 50)
    return header.ljust(80, b" ")[:80] + np.uint32(facets.shape[0]).tobytes() + facets.tobytes()


def oracle_stl_dedup(oracle: "Oracle", facets: np.ndarray):
    L = oracle.lib
    L.oracle_stl_dedup.restype = C.c_uint32
    L.oracle_stl_dedup.argtypes = [vp, C.c_uint32, vp, vp]
    facets = np.ascontiguousarray(facets, np.uint8).reshape(-1, 50)
    nt = facets.shape[0]
    v = np.zeros((max(nt, 1) * 3, 3), np.float32)
    t = np.zeros((nt, 3), np.uint32)
    nv = L.oracle_stl_dedup(_ptr(facets), nt, _ptr(v), _ptr(t))
    return v[:nv].copy(), t


def oracle_triangle_normals(oracle: "Oracle", vertices: np.ndarray, triangles: np.ndarray) -> np.ndarray:
    L = oracle.lib
    L.oracle_triangle_normals.restype = None
    L.oracle_triangle_normals.argtypes = [vp, vp, C.c_uint32, vp]
    vertices = np.ascontiguousarray(vertices, np.float32)
    triangles = np.ascontiguousarray(triangles, np.uint32).reshape(-1, 3)
    out = np.zeros((triangles.shape[0], 3), np.float32)
    L.oracle_triangle_normals(_ptr(vertices), _ptr(triangles), triangles.shape[0], _ptr(out))
    return out


def c_read_stl(lib_path: str, filename: str, full: bool = False):
    """trico_read_stl / trico_read_stl_full of a library with the reference's trico_io signatures"""
    L = C.CDLL(lib_path)
    libc = C.CDLL(None)
    libc.free.argtypes = [vp]
    nv, nt = C.c_uint32(0), C.c_uint32(0)
    pv, pt, pn, pa = vp(), vp(), vp(), vp()
    if full:
        L.trico_read_stl_full.argtypes = [C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.c_char_p]
        ok = L.trico_read_stl_full(C.byref(nv), C.byref(pv), C.byref(nt), C.byref(pt), C.byref(pn), C.byref(pa), filename.encode())
    else:
        L.trico_read_stl.argtypes = [C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(vp), C.c_char_p]
        ok = L.trico_read_stl(C.byref(nv), C.byref(pv), C.byref(nt), C.byref(pt), filename.encode())
    if not ok:
        return None

    def take(p, count, dtype):
        if not (count and p.value):
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,)).copy().view(dtype)
    out = (take(pv, nv.value * 3, np.float32).reshape(-1, 3), take(pt, nt.value * 3, np.uint32).reshape(-1, 3))
    if full:
        out += (take(pn, nt.value * 3, np.float32).reshape(-1, 3), take(pa, nt.value, np.uint16))
    return out
