"""Compression-ratio report: our chunked archives vs the reference's whole-stream archives.

    python tools/ratio_report.py            (needs a GPU; run under gpurun)

Prints, for the bundled bunny (config C1) and a synthetic C2-style mesh, the per-stream sizes of
the reference archive (oracle/_ref, CPU) and of ours for several chunk geometries.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import trico_b200  # noqa: E402
from checkers import Ref, have_ref  # noqa: E402
from trico_b200.synth import grid_mesh  # noqa: E402


def main():
    dev = trico_b200.Device(0)
    ref = Ref() if have_ref() else None
    full = dict(np.load(os.path.join(ROOT, "tests", "golden", "bunny_full.npz")))
    meshes = {"bunny": (full["vertices"], full["triangles"]), "grid 1500x1400 j=1": grid_mesh(1500, 1400, 1.0, 1)}
    for name, (v, t) in meshes.items():
        nv, nt = v.shape[0], t.shape[0]
        print(f"== {name}: {nv} vertices, {nt} triangles")
        if ref:
            rv = len(ref.encode([(1, v, nv)])) - 8
            rt = len(ref.encode([(3, t, nt)])) - 8
            print(f"   reference: vertices {rv} B (ratio {v.nbytes / rv:.4f}), triangles {rt} B (ratio {t.nbytes / rt:.4f}), total ratio {(v.nbytes + t.nbytes) / (rv + rt):.4f}")
        else:
            rv = rt = None
        for l2 in (7, 8, 9, 10):
            s = len(dev.encode_stream(1, v.reshape(-1), nv, l2))
            extra = "" if rv is None else f" ({100 * (rv / s - 1):+.2f} % vs reference)"
            print(f"   ours FPC chunk 2^{l2}: {s} B ratio {v.nbytes / s:.4f}{extra}")
        for l2 in (12, 13, 14, 15):
            s = len(dev.encode_stream(3, t.reshape(-1), nt, l2))
            extra = "" if rt is None else f" ({100 * (rt / s - 1):+.2f} % vs reference)"
            print(f"   ours LZ4 block 2^{l2}: {s} B ratio {t.nbytes / s:.4f}{extra}")


if __name__ == "__main__":
    main()
