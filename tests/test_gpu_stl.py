"""GPU: the STL front-end (trico_io/iostl.c:70-138, trico_decoder/main.c:439-470) through the
C ABI of include/trico_b200_io.h against the oracle, the reference-generated bunny fixture and - when its
prebuilt library travelled along - the reference's own trico_read_stl.  Bit-exact."""
import os

import numpy as np
import pytest

from checkers import (c_read_stl, have_ref, oracle_stl_dedup, oracle_triangle_normals, stl_facets, stl_file_bytes, REF_SO)
from stl_cases import cases, grid_mesh, soup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from trico_b200 import Device
    d = Device(0)
    yield d
    d.close()


def _same(a, b):
    return a.shape == b.shape and a.view(np.uint32).tobytes() == b.view(np.uint32).tobytes()


@pytest.mark.parametrize("name", list(cases()))
def test_dedup_matches_oracle(dev, oracle, name):
    v, t = cases()[name]
    rng = np.random.default_rng(5)
    normals = rng.standard_normal((t.shape[0], 3)).astype(np.float32)
    attrs = rng.integers(0, 65536, t.shape[0]).astype(np.uint16)
    facets = stl_facets(v, t, normals, attrs)
    gv, gt, gn, ga = dev.stl_dedup(facets, full=True)
    ov, ot = oracle_stl_dedup(oracle, facets)
    assert _same(gv, ov), name
    assert np.array_equal(gt, ot), name
    assert _same(gn, normals) and np.array_equal(ga, attrs)
    gv2, gt2 = dev.stl_dedup(facets)                         # without the optional outputs
    assert _same(gv2, ov) and np.array_equal(gt2, ot)


def test_one_point_runs_no_sort_pass(dev):
    v, t = cases()["one_point"]
    dev.stl_dedup(stl_facets(v, t))
    assert dev.lib.tb200_stl_last_sort_passes() == 0
    v, t = cases()["grid"]
    dev.stl_dedup(stl_facets(v, t))
    assert 0 < dev.lib.tb200_stl_last_sort_passes() <= 12


def test_bunny_fixture(dev, golden):
    b = golden["bunny_full"]
    gv, gt = dev.stl_dedup(stl_facets(b["vertices"], b["triangles"]))
    assert _same(gv, b["vertices"]) and np.array_equal(gt, b["triangles"])


def test_unaligned_facets(dev, oracle):
    """facets at an odd device address (a whole STL file uploaded as it is: facets start at byte 84, but a
    caller may hand over any address)"""
    v, t = soup(3000, 21, 400)
    facets = stl_facets(v, t)
    ov, ot = oracle_stl_dedup(oracle, facets)
    for shift in (1, 2, 84):
        blob = np.concatenate([np.zeros(shift, np.uint8), facets.reshape(-1)])
        d_f, d_v, d_t = dev.upload(blob), dev.alloc(t.shape[0] * 36), dev.alloc(t.shape[0] * 12)
        nv = dev.stl_dedup_device(d_f.ptr + shift, t.shape[0], d_v.ptr, d_t.ptr)
        assert _same(dev.download(d_v.ptr, nv * 12).view(np.float32).reshape(-1, 3), ov)
        assert np.array_equal(dev.download(d_t.ptr, t.shape[0] * 12).view(np.uint32).reshape(-1, 3), ot)


def test_large_mesh(dev, oracle):
    """2 M triangles (many sort tiles, multi-chunk scans): against the oracle, and the round trip property"""
    v, t = grid_mesh(1001, 1000, 9)
    facets = stl_facets(v, t)
    gv, gt = dev.stl_dedup(facets)
    ov, ot = oracle_stl_dedup(oracle, facets)
    assert _same(gv, ov) and np.array_equal(gt, ot)
    assert gv.shape[0] == v.shape[0]
    assert np.array_equal(gv[gt.reshape(-1)].view(np.uint32), v[t.reshape(-1)].view(np.uint32))


def test_read_stl_files(oracle, tmp_path):
    import trico_b200
    for name in ("grid", "soup4097", "extremes"):
        v, t = cases()[name]
        rng = np.random.default_rng(2)
        normals = rng.standard_normal((t.shape[0], 3)).astype(np.float32)
        attrs = rng.integers(0, 65536, t.shape[0]).astype(np.uint16)
        facets = stl_facets(v, t, normals, attrs)
        path = os.path.join(tmp_path, name + ".stl")
        open(path, "wb").write(stl_file_bytes(facets))
        gv, gt, gn, ga = trico_b200.read_stl(path, full=True)
        ov, ot = oracle_stl_dedup(oracle, facets)
        assert _same(gv, ov) and np.array_equal(gt, ot) and _same(gn, normals) and np.array_equal(ga, attrs)
        gv, gt = trico_b200.read_stl(path)
        assert _same(gv, ov) and np.array_equal(gt, ot)
        if have_ref():
            rv, rt = c_read_stl(REF_SO, path)
            assert _same(gv, rv) and np.array_equal(gt, rt)
    # the reference's refusals (iostl.c:148-161, :189-190)
    L = trico_b200.load()
    assert c_read_stl(trico_b200.LIB_PATH, os.path.join(tmp_path, "missing.stl")) is None
    open(os.path.join(tmp_path, "ascii.stl"), "wb").write(b"solid x\n" + b" " * 200)
    assert c_read_stl(trico_b200.LIB_PATH, os.path.join(tmp_path, "ascii.stl")) is None
    whole = stl_file_bytes(stl_facets(*cases()["grid"]))
    open(os.path.join(tmp_path, "cut.stl"), "wb").write(whole[:-7])
    assert c_read_stl(trico_b200.LIB_PATH, os.path.join(tmp_path, "cut.stl")) is None
    open(os.path.join(tmp_path, "empty.stl"), "wb").write(stl_file_bytes(np.zeros((0, 50), np.uint8)))
    ev, et = c_read_stl(trico_b200.LIB_PATH, os.path.join(tmp_path, "empty.stl"))
    assert ev.shape[0] == 0 and et.shape[0] == 0
    assert L is not None


def test_triangle_normals(dev, oracle, golden):
    b = golden["bunny_full"]
    got = dev.triangle_normals(b["vertices"], b["triangles"])
    assert _same(got, oracle_triangle_normals(oracle, b["vertices"], b["triangles"]))
    v, t = cases()["soup_dense"]                             # many degenerate triangles: zero length
    assert _same(dev.triangle_normals(v, t), oracle_triangle_normals(oracle, v, t))
    v, t = cases()["extremes"]                               # overflow to inf / nan follows the same operations;
    g, o = dev.triangle_normals(v, t), oracle_triangle_normals(oracle, v, t)   # the bits of a GENERATED nan are the hardware's
    assert np.array_equal(np.isnan(g), np.isnan(o)) and g[~np.isnan(g)].tobytes() == o[~np.isnan(o)].tobytes()
    # host-buffer entry point
    import ctypes as C
    out = np.zeros((b["triangles"].shape[0], 3), np.float32)
    vv, tt = np.ascontiguousarray(b["vertices"]), np.ascontiguousarray(b["triangles"])
    assert dev.lib.trico_b200_triangle_normals(vv.ctypes.data_as(C.c_void_p), vv.shape[0], tt.ctypes.data_as(C.c_void_p), tt.shape[0], out.ctypes.data_as(C.c_void_p))
    assert _same(out, got)


def test_write_stl_is_the_reference_file(tmp_path, golden):
    """trico_write_stl (iostl.c:261-320): facets gathered on the device, the same file byte for byte"""
    import ctypes as C
    import trico_b200
    L = trico_b200.load()
    b = golden["bunny_full"]
    v, t = np.ascontiguousarray(b["vertices"]), np.ascontiguousarray(b["triangles"])
    rng = np.random.default_rng(4)
    normals = rng.standard_normal((t.shape[0], 3)).astype(np.float32)
    attrs = rng.integers(0, 65536, t.shape[0]).astype(np.uint16)
    vp = C.c_void_p
    ref = None
    if have_ref():
        ref = C.CDLL(REF_SO)
        ref.trico_write_stl.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_char_p]
    for k, (nrm, att) in enumerate(((normals, attrs), (None, None), (normals, None))):
        path = os.path.join(tmp_path, f"w{k}.stl").encode()
        args = (v.ctypes.data_as(vp), t.ctypes.data_as(vp), t.shape[0], nrm.ctypes.data_as(vp) if nrm is not None else None,
                att.ctypes.data_as(vp) if att is not None else None)
        assert L.trico_write_stl(*args, path)
        mine = open(path, "rb").read()
        assert mine[84:] == stl_facets(v, t, nrm, att).tobytes() and mine[80:84] == np.uint32(t.shape[0]).tobytes()
        if ref is not None:
            rpath = os.path.join(tmp_path, f"r{k}.stl").encode()
            assert ref.trico_write_stl(*args, rpath)
            assert open(rpath, "rb").read() == mine
    # and back through the reader: the same mesh
    gv, gt = trico_b200.read_stl(os.path.join(tmp_path, "w1.stl"))
    assert _same(gv, v) and np.array_equal(gt, t)
