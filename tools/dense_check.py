#!/usr/bin/env python
"""Round trip + oracle validation of the tiled bunny's index stream (dense LZ4 planes); run under
compute-sanitizer when hunting memory errors:  tools/dense_check.py [triangles]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import trico_b200
from checkers import Oracle
want = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
z = np.load(os.path.join(ROOT, "tests", "golden", "bunny_full.npz"))
t0 = z["triangles"].astype(np.uint32)
nv0 = int(z["vertices"].shape[0])
tiles = max(1, -(-want // t0.shape[0]))
t = np.concatenate([t0 + np.uint32(k * nv0) for k in range(tiles)])
dev = trico_b200.Device(0)
s = dev.encode_stream(3, t.reshape(-1), t.shape[0])
print("triangles", t.shape[0], "stream bytes", len(s), "ratio", t.nbytes / len(s))
back = dev.decode_stream(s)
assert back.tobytes() == t.tobytes(), "GPU round trip mismatch"
orc = Oracle()
log2 = s[6]; n = t.size
nr = (n + (1 << log2) - 1) >> log2
sizes = np.frombuffer(s[15:15 + 8 * nr], np.uint16).astype(np.int64)
offs = np.concatenate([[0], np.cumsum(sizes)]); pay = s[15 + 8 * nr:]
planes = t.reshape(-1).view(np.uint8).reshape(-1, 4)
for k in range(min(nr, 40)):
    for p in range(4):
        g = k * 4 + p
        raw = np.ascontiguousarray(planes[k << log2:(k + 1) << log2, p]).tobytes()
        blk = pay[offs[g]:offs[g + 1]]
        assert orc.lz4_validate(blk, len(raw)) >= 0, (k, p)
        assert orc.lz4_decompress(blk, len(raw)) == raw, (k, p)
print("ok: first 40 ranges validated by the oracle; plane sizes of range 0:", sizes[:4])

# ---- per-block comparison with the executable model of the lane parser ----
sys.path.insert(0, os.path.join(ROOT, "tools", "sim"))
from lz4_lanes_model import parse_lanes, encode as model_encode

def parse_block(blk):
    ip = 0; seqs = []; n = len(blk)
    while ip < n:
        tok = blk[ip]; ip += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = blk[ip]; ip += 1; lit += b
                if b != 255: break
        ip += lit
        if ip >= n:
            seqs.append((lit, 0, 0)); break
        off = blk[ip] | (blk[ip + 1] << 8); ip += 2
        m = tok & 15
        if m == 15:
            while True:
                b = blk[ip]; ip += 1; m += b
                if b != 255: break
        seqs.append((lit, m + 4, off))
    return seqs

shown = 0
for k in range(min(nr, 6)):
    g = k * 4 + 1
    raw = np.ascontiguousarray(planes[k << log2:(k + 1) << log2, 1]).tobytes()
    blk = pay[offs[g]:offs[g + 1]]
    ms = parse_lanes(raw, HLOG=11, S=64, longest=True, merge=False)
    mb = model_encode(ms, raw)
    print(f"range {k} plane 1: gpu {len(blk)} B, model {len(mb)} B", "IDENTICAL" if mb == bytes(blk) else "")
    if mb != bytes(blk) and shown < 2:
        shown += 1
        gs = parse_block(blk)
        pos_g = pos_m = 0
        for i, (a, b) in enumerate(zip(gs, ms)):
            if a != tuple(int(x) for x in b):
                print(f"  first difference at sequence {i}: input position gpu {pos_g} model {pos_m}: gpu {a} model {tuple(int(x) for x in b)}")
                for j in range(max(0, i - 2), min(len(gs), i + 6)):
                    print("     gpu", gs[j], "   model", tuple(int(x) for x in ms[j]) if j < len(ms) else None)
                break
            pos_g += a[0] + a[1]; pos_m += b[0] + b[1]
