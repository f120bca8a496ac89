"""CPU: the DEVICE SOURCE of the lane-parallel LZ4 parser (trico_b200/csrc/lz4_lanes.cuh) compiled
with g++ against a 32-thread emulation of the warp (tools/sim/warp_emu.hpp, one std::thread per
lane, every *_sync intrinsic a barrier that also checks that all lanes arrived at the same call
site).  Every block must be a valid LZ4 block (checked with the oracle's validator and decoder)
that decodes to the input, and the sizes must equal those of the executable Python model of the
parse (tools/sim/lz4_lanes_model.py) within a few percent."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("emu") / "lanes_emu")
    cmd = [gxx, "-O1", "-std=c++20", "-pthread", "-DTB200_HOST_EMU", "-I", os.path.join(ROOT, "tools", "sim"),
           "-I", os.path.join(ROOT, "trico_b200", "csrc"), os.path.join(ROOT, "tools", "sim", "lanes_emu.cpp"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def _planes(golden):
    tri = golden["bunny_full"]["triangles"].astype(np.uint32).reshape(-1)
    p1 = np.ascontiguousarray(tri.view(np.uint8).reshape(-1, 4)[:, 1])
    rng = np.random.default_rng(3)
    runs = np.repeat(rng.integers(0, 255, 4000).astype(np.uint8), rng.integers(1, 40, 4000))
    noise = rng.integers(0, 256, 40000).astype(np.uint8)
    lowent = rng.integers(0, 6, 50000).astype(np.uint8)
    return {"bunny_plane1": p1[:16384 * 6 + 777], "runs": runs[:50000], "noise": noise, "low_entropy": lowent,
            "tiny": p1[:11], "short": p1[:300]}


@pytest.mark.parametrize("name", ["bunny_plane1", "runs", "noise", "low_entropy", "tiny", "short"])
def test_device_source_of_the_lane_parser_on_the_cpu(emu, golden, oracle, tmp_path, name):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools", "sim"))
    from lz4_lanes_model import parse_lanes, encode
    plane = _planes(golden)[name]
    f = tmp_path / "plane.bin"
    plane.tofile(f)
    sizes_file = tmp_path / "sizes.txt"
    r = subprocess.run([emu, str(f), "16384", str(sizes_file)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr          # round trip of every block, no lane left the common path
    sizes = [int(x) for x in open(sizes_file).read().split()]
    B = 16384
    model = []
    for b0 in range(0, len(plane), B):
        raw = plane[b0:b0 + B].tobytes()
        blk = encode(parse_lanes(raw, HLOG=11, S=64, longest=True, merge=False), raw)
        assert oracle.lz4_validate(blk, len(raw)) >= 0
        model.append(len(blk))
    assert len(sizes) == len(model)
    assert abs(sum(sizes) - sum(model)) <= 0.03 * sum(model) + 16, (sizes, model)     # (ties are broken differently in a few places)
