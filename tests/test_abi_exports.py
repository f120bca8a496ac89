"""CPU-only: libtrico_b200.so builds, loads, and exports every symbol include/*.h declares.
No compute calls (there is no GPU here and the library has no CPU fallback)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#\s*define[^\n]*", "", text)
    names = re.findall(r"(?:TRICO_API|TB200_API)[^;(]*?\b(\w+)\s*\(", text)
    return sorted(set(names))


def test_headers_declare_the_reference_surface():
    api = _declared("trico_b200.h")
    # trico.h:36-94 (54) + transpose_aos_to_soa.h:12-38 (14) + floating_point_stream_compression.h:11-17 (4)
    ref = [n for n in api if not n.startswith("trico_b200_")]
    assert len(ref) == 72, len(ref)
    for must in ("trico_open_archive_for_writing", "trico_read_attributes_uint64", "trico_skip_next_stream",
                 "trico_transpose_uint64_soa_to_aos", "trico_decompress_double_precision"):
        assert must in ref


def test_library_exports_every_declared_symbol():
    import trico_b200
    lib = trico_b200.load()
    for header in ("trico_b200.h", "trico_b200_device.h", "trico_b200_io.h"):
        names = _declared(header)
        assert names
        for n in names:
            assert hasattr(lib, n), f"{n} declared in {header} but not exported"


def test_no_cuda_means_loud_failure_not_fallback():
    """Without a CUDA device the data path refuses to work; only the framing still does."""
    import trico_b200
    lib = trico_b200.load()
    if lib.tb200_device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    assert not lib.tb200_ctx_create(0, None)
    assert b"no CUDA device" in lib.tb200_last_error() or lib.tb200_last_error()
    L = C.CDLL(trico_b200.LIB_PATH)
    L.trico_open_archive_for_writing.restype = C.c_void_p
    L.trico_open_archive_for_writing.argtypes = [C.c_uint64]
    L.trico_get_size.restype = C.c_uint64
    L.trico_get_size.argtypes = [C.c_void_p]
    L.trico_write_vertices.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.trico_close_archive.argtypes = [C.c_void_p]
    a = L.trico_open_archive_for_writing(64)
    assert a and L.trico_get_size(a) == 8            # header only (trico.tests test_header)
    import numpy as np
    v = np.zeros(30, np.float32)
    assert L.trico_write_vertices(a, v.ctypes.data_as(C.c_void_p), 10) == 0   # no device -> failure, not a CPU codec
    assert L.trico_get_size(a) == 8
    L.trico_close_archive(a)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "trico_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".inc", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("the CPU oracle", "").lower() or f in ("build.py",), (dirpath, f)


def test_dynamic_shared_memory_is_declared_alike_everywhere():
    """All `extern __shared__` arrays of the one CUDA translation unit are ONE symbol: a larger alignment on any
    of them moves the dynamic shared memory of every kernel ("A shared-memory hazard")."""
    csrc = os.path.join(ROOT, "trico_b200", "csrc")
    seen = []
    for f in sorted(os.listdir(csrc)):
        text = open(os.path.join(csrc, f), errors="ignore").read()
        seen += [(f, m) for m in re.findall(r"extern\s+__shared__\s+(?:__align__\((\w+)\))?", text)]
    assert len(seen) >= 10
    assert all(a == "16" for _, a in seen), seen


def test_library_carries_tma_bulk_copies():
    """the float encoder's slab staging and the assembly's stage ring are cp.async.bulk + mbarrier: UBLKCP / SYNCS in SASS"""
    import shutil
    import subprocess
    import trico_b200
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        import pytest
        pytest.skip("cuobjdump not available")
    trico_b200.load()
    sass = subprocess.run([tool, "-sass", trico_b200.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    assert sass.count("UBLKCP") >= 32 and "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK" in sass


def test_stl_front_end_without_a_gpu(tmp_path):
    """include/trico_b200_io.h: the refusals of the reference's reader need no device; a real file fails loudly
    without one (no CPU de-duplication hides behind the API)."""
    import numpy as np
    import trico_b200
    sys_path_oracle = os.path.join(ROOT, "oracle")
    import sys
    sys.path.insert(0, sys_path_oracle)
    from checkers import c_read_stl, stl_facets, stl_file_bytes
    lib = trico_b200.load()
    assert lib.tb200_stl_dedup_scratch_bytes(0) > 0
    assert lib.tb200_stl_dedup_scratch_bytes(1000000) >= 2 * 16 * 3000000 + 4 * 3000000      # two record arrays + the flags
    assert c_read_stl(trico_b200.LIB_PATH, os.path.join(tmp_path, "missing.stl")) is None
    open(os.path.join(tmp_path, "ascii.stl"), "wb").write(b"solid x\n" + b" " * 200)
    assert c_read_stl(trico_b200.LIB_PATH, os.path.join(tmp_path, "ascii.stl")) is None           # iostl.c:157-161
    open(os.path.join(tmp_path, "empty.stl"), "wb").write(stl_file_bytes(np.zeros((0, 50), np.uint8)))
    ev, et = c_read_stl(trico_b200.LIB_PATH, os.path.join(tmp_path, "empty.stl"))                # iostl.c:72-73: nothing to do
    assert ev.shape[0] == 0 and et.shape[0] == 0
    if lib.tb200_device_count() > 0:
        return
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    open(os.path.join(tmp_path, "one.stl"), "wb").write(stl_file_bytes(stl_facets(v, np.array([[0, 1, 2]], np.uint32))))
    assert c_read_stl(trico_b200.LIB_PATH, os.path.join(tmp_path, "one.stl")) is None
    assert lib.tb200_last_error()
