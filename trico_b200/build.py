"""In-tree build of libtrico_b200.so (host C archive layer + sm_100a CUDA kernels).

    python -m trico_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but travels to the
GPU box with the source snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# TRICO_B200_LIB selects another build of the same sources (kernel experiments: see build_variant)
LIB = os.environ.get("TRICO_B200_LIB") or os.path.join(LIBDIR, "libtrico_b200.so")
ROOT = os.path.dirname(HERE)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"] + os.environ.get("TB200_NVCC_EXTRA", "").split()
CC_FLAGS = ["-O2", "-fPIC", "-std=c11", "-Wall", "-Wextra", "-Wno-unused-parameter", "-fvisibility=hidden",
            "-D_POSIX_C_SOURCE=200809L"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 library cannot be built (there is no CPU fallback)")


def _sources():
    deps = []
    for d in (CSRC, os.path.join(ROOT, "include"), os.path.join(ROOT, "include", "trico")):
        if not os.path.isdir(d):
            continue
        for f in os.listdir(d):
            p = os.path.join(d, f)
            if os.path.isfile(p):
                deps.append(p)
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    inc = ["-I", os.path.join(ROOT, "include")]
    obj_c = os.path.join(LIBDIR, "archive.o")
    obj_cu = os.path.join(LIBDIR, "device_api.o")
    cmds = [
        [os.environ.get("CC", "gcc"), *CC_FLAGS, *inc, "-c", os.path.join(CSRC, "archive.c"), "-o", obj_c],
        [nvcc, *NVCC_FLAGS, *inc, "-c", os.path.join(CSRC, "device_api.cu"), "-o", obj_cu],
        [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, obj_c, obj_cu, "-lcudart", "-ldl"],
    ]
    for cmd in cmds:
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return LIB


def build_variant(name: str, defs, verbose: bool = False) -> str:
    """Experiment build: the same sources with extra -D definitions -> lib/libtrico_b200_<name>.so
    (load it with TRICO_B200_LIB=<path>)."""
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    inc = ["-I", os.path.join(ROOT, "include")]
    obj_c = os.path.join(LIBDIR, "archive.o")
    obj_cu = os.path.join(LIBDIR, f"device_api_{name}.o")
    out = os.path.join(LIBDIR, f"libtrico_b200_{name}.so")
    dd = [f"-D{d}" for d in defs]
    cmds = [
        [os.environ.get("CC", "gcc"), *CC_FLAGS, *inc, "-c", os.path.join(CSRC, "archive.c"), "-o", obj_c],
        [nvcc, *NVCC_FLAGS, *dd, *inc, "-c", os.path.join(CSRC, "device_api.cu"), "-o", obj_cu],
        [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, obj_c, obj_cu, "-lcudart", "-ldl"],
    ]
    for cmd in cmds:
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    os.remove(obj_cu)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:], verbose=True))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
