"""CPU: the oracle's restatement of the STL front-end against the reference-generated
bunny fixture and, where its prebuilt library is present, against the reference's own trico_read_stl."""
import os

import numpy as np
import pytest

from checkers import (c_read_stl, have_ref, oracle_stl_dedup, oracle_triangle_normals, stl_facets, stl_file_bytes, REF_SO)
from stl_cases import cases


def test_bunny_dedup_matches_the_reference_fixture(oracle, golden):
    """tests/golden/bunny_full.npz holds what the reference's trico_read_stl made of StanfordBunny.stl;
    expanding it back to facets and de-duplicating again must give the same mesh, bit for bit."""
    b = golden["bunny_full"]
    v, t = oracle_stl_dedup(oracle, stl_facets(b["vertices"], b["triangles"]))
    assert v.shape == b["vertices"].shape
    assert v.view(np.uint32).tobytes() == b["vertices"].view(np.uint32).tobytes()
    assert np.array_equal(t, b["triangles"])


@pytest.mark.parametrize("name", [k for k in cases() if k != "signed_zero"])
def test_oracle_dedup_against_the_reference_reader(oracle, name, tmp_path):
    if not have_ref():
        pytest.skip("oracle/_ref/libtrico_ref.so not built (needs /root/reference)")
    v, t = cases()[name]
    rng = np.random.default_rng(11)
    normals = rng.standard_normal((t.shape[0], 3)).astype(np.float32)
    attrs = rng.integers(0, 65536, t.shape[0]).astype(np.uint16)
    facets = stl_facets(v, t, normals, attrs)
    path = os.path.join(tmp_path, "m.stl")
    open(path, "wb").write(stl_file_bytes(facets))
    rv, rt, rn, ra = c_read_stl(REF_SO, path, full=True)
    ov, ot = oracle_stl_dedup(oracle, facets)
    assert rv.view(np.uint32).tobytes() == ov.view(np.uint32).tobytes()
    assert np.array_equal(rt, ot)
    assert rn.view(np.uint32).tobytes() == normals.view(np.uint32).tobytes() and np.array_equal(ra, attrs)
    # the mesh is the same mesh: every corner still has its position
    assert np.array_equal(ov[ot.reshape(-1)].view(np.uint32), v[t.reshape(-1)].view(np.uint32))


def test_oracle_dedup_properties(oracle):
    for name, (v, t) in cases().items():
        ov, ot = oracle_stl_dedup(oracle, stl_facets(v, t))
        assert np.array_equal(ov[ot.reshape(-1)], v[t.reshape(-1)]), name           # float equality (+0 == -0)
        key = ov.astype(np.float64)
        order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
        assert np.array_equal(order, np.arange(ov.shape[0])), name                   # sorted by x, y, z
        assert ov.shape[0] == np.unique(key + 0.0, axis=0).shape[0], name            # and unique


def test_oracle_normals_against_float32_numpy(oracle, golden):
    """main.c:441-469 restated a second time, in numpy float32 (every operation rounded on its own)"""
    b = golden["bunny_full"]
    v, t = b["vertices"], b["triangles"]
    f = np.float32
    p0, p1, p2 = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    a, c = (p1 - p0).astype(f), (p2 - p0).astype(f)
    nx = (a[:, 1] * c[:, 2]).astype(f) - (a[:, 2] * c[:, 1]).astype(f)
    ny = (a[:, 2] * c[:, 0]).astype(f) - (a[:, 0] * c[:, 2]).astype(f)
    nz = (a[:, 0] * c[:, 1]).astype(f) - (a[:, 1] * c[:, 0]).astype(f)
    s = ((nx * nx).astype(f) + (ny * ny).astype(f)).astype(f) + (nz * nz).astype(f)
    length = np.sqrt(s.astype(np.float64)).astype(f)
    safe = np.where(length != 0, length, f(1))
    want = np.stack([np.where(length != 0, nx / safe, nx), np.where(length != 0, ny / safe, ny), np.where(length != 0, nz / safe, nz)], 1).astype(f)
    got = oracle_triangle_normals(oracle, v, t)
    assert got.view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    # degenerate triangle: the vector is left as it is (main.c:466-468)
    d = oracle_triangle_normals(oracle, np.zeros((3, 3), f), np.array([[0, 1, 2]], np.uint32))
    assert np.array_equal(d, np.zeros((1, 3), f))
