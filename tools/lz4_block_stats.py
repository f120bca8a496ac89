#!/usr/bin/env python
"""Sequence statistics of the GPU compressor's LZ4 plane blocks (needs a GPU): tools/lz4_block_stats.py [W H]"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trico_b200
from trico_b200.synth import grid_mesh

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1500, 1000)
v, t = grid_mesh(W, H, jitter=1.0, seed=1)
d = trico_b200.Device(0)
s = d.encode_stream(3, t.reshape(-1), t.shape[0])
log2 = s[6]
n = t.size
nr = (n + (1 << log2) - 1) >> log2
nch = nr * 4
sizes = np.frombuffer(s[15:15 + 2 * nch], np.uint16).astype(np.int64)
pay = s[15 + 2 * nch:]
offs = np.concatenate([[0], np.cumsum(sizes)])

def parse(blk):
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

for p in range(4):
    ll, ml, of, nseq, tot, nb = [], [], [], 0, 0, 0
    for k in range(min(nr, 60)):
        g = k * 4 + p
        sq = parse(pay[offs[g]:offs[g + 1]])
        nseq += len(sq); tot += sizes[g]; nb += 1
        ll += [a for a, b, c in sq]; ml += [b for a, b, c in sq if b]; of += [c for a, b, c in sq if b]
    ll, ml, of = np.array(ll), np.array(ml), np.array(of)
    print(f"plane {p}: ratio {nb * (1 << log2) / tot:.2f} seq/block {nseq / nb:.1f} lit avg {ll.mean():.1f} med {np.median(ll)} max {ll.max()} | match avg {ml.mean() if len(ml) else 0:.1f} med {np.median(ml) if len(ml) else 0} | offset med {np.median(of) if len(of) else 0} <16: {np.mean(of < 16) if len(of) else 0:.2f}")
    print("   match len hist (<=20):", np.bincount(np.minimum(ml, 20))[:21] if len(ml) else [])
    print("   offsets hist (<=16):", np.bincount(np.minimum(of, 17))[:18] if len(of) else [])
