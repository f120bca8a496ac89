"""Pins the CPU oracle (oracle/trico_oracle.c) to the reference.

* against tests/golden/ - bytes produced by the unmodified reference (make_golden.py), and
* against oracle/_ref/libtrico_ref.so side by side on fresh random inputs, when it is present.

CPU only.
"""
import numpy as np
import pytest


def _fromhex(h, dtype):
    return np.frombuffer(bytes.fromhex(h), dtype)


# ---------------------------------------------------------------------------------- golden
def test_fpc32_known_answers(oracle, golden):
    kat = golden["kat"]["fpc32"]
    assert len(kat) > 100
    for case in kat:
        vals = _fromhex(case["in"], np.uint32)
        want = bytes.fromhex(case["out"])
        got = oracle.fpc_compress(vals, case["e1"], case["e2"])
        assert got == want, (case["e1"], case["e2"], vals.size)
        assert np.array_equal(oracle.fpc_decompress(want, 4), vals)
        assert oracle.fpc_stream_bytes(want, 4) == len(want)


def test_fpc64_known_answers(oracle, golden):
    kat = golden["kat"]["fpc64"]
    assert len(kat) > 80
    for case in kat:
        vals = _fromhex(case["in"], np.uint64)
        want = bytes.fromhex(case["out"])
        assert oracle.fpc_compress(vals, case["e1"], case["e2"]) == want
        assert np.array_equal(oracle.fpc_decompress(want, 8), vals)
        assert oracle.fpc_stream_bytes(want, 8) == len(want)


def test_survey_hand_vectors(oracle):
    # SURVEY.md section 8(c), produced by the compiled reference
    one = np.array([1.0], np.float32)
    assert oracle.fpc_compress(one, 4, 10).hex() == "25" "00000001" "24924c" "3f800000" + "00" * 7
    eight = np.array([1, 1, 2, 3, 4, 5, 6, 7], np.float32)
    assert oracle.fpc_compress(eight, 4, 10).hex() == ("25" "00000008" "6dbf2c" "3f800000" "00" "7f800000" "400000"
                                                       "c00000" "200000" "600000" "200000")
    assert oracle.fpc_compress(np.array([1.0]), 20, 20).hex() == "aa" "00000001" "18" "3ff0000000000000" "00"
    assert oracle.fpc_compress(np.array([1.0, 1.0]), 20, 20).hex() == "aa" "00000002" "98" "3ff0000000000000" "00"


def test_lz4_decoder_on_reference_blocks(oracle, golden):
    for case in golden["kat"]["lz4"]:
        raw, block = bytes.fromhex(case["in"]), bytes.fromhex(case["out"])
        assert oracle.lz4_decompress(block, len(raw)) == raw
        assert oracle.lz4_validate(block, len(raw)) == len(raw)


def test_lz4_own_compressor_is_valid(oracle, golden):
    for case in golden["kat"]["lz4"]:
        raw = bytes.fromhex(case["in"])
        block = oracle.lz4_compress(raw)
        assert oracle.lz4_validate(block, len(raw)) == len(raw)
        assert oracle.lz4_decompress(block, len(raw)) == raw


def test_v0_streams_from_reference(oracle, golden):
    b = golden["bunny"]
    for ty in range(1, 21):
        if ty in (6, 8):
            continue  # the reference cannot write these tags (trico.c:622,:627)
        blob = b[f"v0_{ty}"].tobytes()
        version, streams = oracle.read_archive(blob)
        assert version == 0 and len(streams) == 1
        t, cnt, arr = streams[0]
        want = b[f"in_{ty}"].reshape(-1)
        assert t == ty
        assert cnt == (int(b[f"cnt_{ty}"]) * (3 if ty == 7 else 1))
        assert arr.tobytes() == want.tobytes(), ty
        lay = oracle.layout(ty)
        if lay["codec"] == 1:
            # FPC streams are deterministic: the oracle's writer must reproduce the reference bytes
            assert oracle.header(0) + oracle.v0_write_stream(ty, want, cnt) == blob, ty


def test_reference_double_uv_quirk(oracle, golden):
    # trico_write_uv_per_vertex_double tags its stream 5 (float uv): trico.c:622
    blob = golden["bunny"]["v0_6_as_written_by_reference"].tobytes()
    assert blob[8] == 5
    # decoding the payload as what it really is (double uv, tag 6) recovers the data
    fixed = blob[:8] + bytes([6]) + blob[9:]
    _, streams = oracle.read_archive(fixed)
    assert streams[0][2].tobytes() == golden["bunny"]["in_6"].tobytes()


def test_v0_multi_stream_archive(oracle, golden):
    b = golden["bunny"]
    version, streams = oracle.read_archive(b["v0_multi"].tobytes())
    assert [s[0] for s in streams] == [1, 3, 9, 13]
    assert streams[0][2].tobytes() == b["vertices"].tobytes()
    assert streams[1][2].tobytes() == b["triangles"].tobytes()


def test_v1_container_roundtrip(oracle, golden):
    b = golden["bunny"]
    for ty in range(1, 21):
        if ty in (6, 8):
            data, cnt = b["in_6"].reshape(-1), int(b["cnt_6"])
        else:
            data, cnt = b[f"in_{ty}"].reshape(-1), int(b[f"cnt_{ty}"]) * (3 if ty == 7 else 1)
        for log2c in (5, 7, 9, 12):
            blob = oracle.header(1) + oracle.v1_write_stream(ty, data, cnt, log2c, 4, 4)
            version, streams = oracle.read_archive(blob)
            assert version == 1
            assert streams[0][0] == ty and streams[0][1] == cnt
            assert streams[0][2].tobytes() == data.tobytes()


# --------------------------------------------------------------------------- side by side
@pytest.mark.parametrize("seed", range(4))
def test_fpc_matches_reference_side_by_side(oracle, ref, seed):
    rng = np.random.default_rng(seed)
    for n in (1, 7, 8, 9, 100, 4097):
        walk = np.cumsum(rng.standard_normal(n) * 0.01) + 1.5
        noise = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
        for vals in (walk.astype(np.float32), noise, np.zeros(n, np.float32)):
            for e in ((4, 10), (4, 4), (2, 2), (12, 14)):
                s = ref.compress(vals, *e)
                assert oracle.fpc_compress(vals.view(np.uint32), *e) == s
                assert np.array_equal(oracle.fpc_decompress(s, 4), ref.decompress(s, 4))
        for vals in (walk, rng.integers(0, 2**63, n, dtype=np.uint64).view(np.float64)):
            for e in ((20, 20), (4, 4)):
                s = ref.compress(vals, *e)
                assert oracle.fpc_compress(vals.view(np.uint64), *e) == s
                assert np.array_equal(oracle.fpc_decompress(s, 8), ref.decompress(s, 8))


def test_lz4_matches_reference_side_by_side(oracle, ref):
    rng = np.random.default_rng(5)
    for n in (0, 1, 12, 13, 100, 5000, 70000, 200000):
        for alphabet in (2, 16, 256):
            raw = bytes(rng.integers(0, alphabet, n, dtype=np.uint8))
            blk = ref.lz4_compress(raw)
            assert oracle.lz4_decompress(blk, n) == raw
            assert oracle.lz4_validate(blk, n) == n
            mine = oracle.lz4_compress(raw)
            assert ref.lz4_decompress(mine, n) == raw


def test_planes_match_reference(oracle, ref):
    import ctypes as C
    rng = np.random.default_rng(6)
    a = rng.integers(0, 2**32, 1001, dtype=np.uint64).astype(np.uint32)
    planes = oracle.planes_split(a)
    bufs = [np.empty(a.size, np.uint8) for _ in range(4)]
    ptrs = [C.c_void_p(b.ctypes.data) for b in bufs]
    ref.lib.trico_transpose_uint32_aos_to_soa(C.byref(ptrs[0]), C.byref(ptrs[1]), C.byref(ptrs[2]), C.byref(ptrs[3]),
                                              C.c_void_p(a.ctypes.data), C.c_uint32(a.size))
    for k in range(4):
        assert np.array_equal(planes[k], bufs[k])
    assert np.array_equal(oracle.planes_merge(planes, np.uint32), a)


def test_whole_bunny_through_reference_and_oracle(oracle, ref, golden):
    import hashlib
    import os
    path = "/root/reference/trico.tests/data/StanfordBunny.stl"
    if not os.path.exists(path):
        pytest.skip("reference test data not mounted")
    v, t = ref.read_stl(path)
    facts = golden["facts"]
    assert (v.shape[0], t.shape[0]) == (facts["nv"], facts["nt"])
    blob = ref.encode([(1, v, v.shape[0]), (3, t, t.shape[0])])
    assert len(blob) == facts["archive_bytes"] == 584613
    assert hashlib.md5(blob).hexdigest() == facts["archive_md5"]
    _, streams = oracle.read_archive(blob)
    assert streams[0][2].tobytes() == v.tobytes() and streams[1][2].tobytes() == t.tobytes()
    # FPC part of the oracle's own v0 writer is byte-identical to the reference
    mine = oracle.v0_write_stream(1, v.reshape(-1), v.shape[0])
    assert blob[8:8 + len(mine)] == mine
