#!/usr/bin/env python
"""Executable model of the lane-parallel LZ4 block parser of trico_b200/csrc/lz4.cuh
(lz4_compress_lanes): same waves, same per-lane state machine, same tables and the same stitch, in
plain Python, so that the parse policy (ratio) can be studied on the CPU.  It produces the sequence
list the kernel produces; `encode` turns it into block bytes.  Not part of the product.

    python tools/sim/lz4_lanes_model.py            # ratios on the bunny's index plane 1 and synthetic planes
"""
import sys, os
import numpy as np

S = 64           # bytes per lane and wave
NL = 32
WAVE = S * NL
MINMATCH, LASTLITERALS, MFLIMIT = 4, 5, 12


def parse_lanes(data: bytes, HLOG=10, OWNBITS=5, lazy=True, use_rep=True, S=64, longest=False, merge=True):
    WAVE = S * NL
    n = len(data)
    d = np.frombuffer(data, np.uint8)
    if n < MFLIMIT + 1:
        return [(n, 0, 0)]
    pad = np.concatenate([d, np.zeros(80, np.uint8)]).astype(np.uint32)
    w32 = pad[:-3] | (pad[1:-2] << 8) | (pad[2:-1] << 16) | (pad[3:] << 24)
    hs = (((w32.astype(np.uint64) * 2654435761) & 0xffffffff) >> (32 - HLOG)).astype(np.int64)
    dz = np.concatenate([d, np.zeros(80, np.uint8)])
    mflimit, matchlimit = n - MFLIMIT, n - LASTLITERALS
    T = [np.zeros(1 << HLOG, np.int64), np.zeros(1 << HLOG, np.int64)]
    OWN = np.zeros((NL, 1 << OWNBITS), np.int64)       # relative positions (stale entries are verified by content)
    seqs = []                 # emitted (lit, ml, off)
    emitted_to = 0            # input position up to which the emitted sequences reach (= anchor of the stitch)
    held = None               # (q, ml, off): the last match of the previous wave, open at the wave boundary; its literals start at emitted_to
    rep = 0
    misses = 0

    def mlen(q, c, lim):
        m = 0
        while q + m < lim and dz[c + m] == dz[q + m]:
            m += 1
        return m

    for wi, w0 in enumerate(range(0, n, WAVE)):
        Tc, To = T[wi & 1], T[(wi & 1) ^ 1]
        sub = [w0 + l * S for l in range(NL)]
        end = [min(s + S, n) for s in sub]
        pos = list(sub)
        anchor = list(sub)
        lst = [[] for _ in range(NL)]          # per lane: (q, ml, off) matches in order
        pend = [None] * NL
        wave_rep = rep
        while True:
            stride = 1 + (misses >> 6)
            probes = []
            for l in range(NL):
                if sub[l] >= n:
                    continue
                q = pos[l]
                can = q <= mflimit and q + MINMATCH <= end[l]
                if not can:
                    if pend[l]:
                        lst[l].append(pend[l]); anchor[l] = pend[l][0] + pend[l][1]; pend[l] = None
                        pos[l] = anchor[l]
                        q = pos[l]
                        can = q <= mflimit and q + MINMATCH <= end[l]
                    if not can:
                        continue
                probes.append(l)
            if not probes:
                break
            res = {}
            for l in probes:
                q = pos[l]; h = hs[q]; ho = h >> (HLOG - OWNBITS)
                cands = []
                if use_rep and wave_rep and q >= wave_rep and w32[q - wave_rep] == w32[q]:
                    cands.append(q - wave_rep)
                co = sub[l] + OWN[l][ho]
                if co < q and w32[co] == w32[q]: cands.append(co)
                c1 = Tc[h]
                if c1 < q and w32[c1] == w32[q]: cands.append(c1)
                c0 = To[h]
                if c0 < q and w32[c0] == w32[q]: cands.append(c0)
                if not cands:
                    res[l] = -1
                elif longest and len(cands) > 1:
                    lim = min(end[l], matchlimit)
                    res[l] = max(cands, key=lambda c: (mlen(q, c, lim), c))
                else:
                    res[l] = cands[0]
            for l in probes:                      # inserts after all look-ups of the step; the highest position of a bucket wins
                q = pos[l]; h = hs[q]
                OWN[l][h >> (HLOG - OWNBITS)] = q - sub[l]
                if q > Tc[h]: Tc[h] = q
            anyhit = False
            for l in probes:
                q = pos[l]; c = res[l]
                ml = 0
                if c >= 0:
                    qq, cc = q, c
                    if not pend[l]:
                        while qq > anchor[l] and cc > 0 and dz[qq - 1] == dz[cc - 1]:
                            qq -= 1; cc -= 1
                    ml = mlen(qq, cc, min(end[l], matchlimit))
                if pend[l]:
                    if ml > pend[l][1]:
                        pend[l] = (qq, ml, qq - cc); pos[l] = q + 1
                    else:
                        lst[l].append(pend[l]); anchor[l] = pend[l][0] + pend[l][1]; pend[l] = None
                        pos[l] = anchor[l]
                    anyhit = True
                    continue
                if ml >= MINMATCH:
                    anyhit = True
                    if lazy and qq == q and q + 1 <= mflimit and q + 1 + MINMATCH <= end[l]:
                        pend[l] = (qq, ml, qq - cc); pos[l] = q + 1
                    else:
                        lst[l].append((qq, ml, qq - cc)); pos[l] = anchor[l] = qq + ml
                else:
                    pos[l] = q + stride
            misses = 0 if anyhit else misses + len(probes)
        best_ml = 15          # the next wave tries the offset of this wave's longest match first (if it is a long one)
        # ---- stitch: lanes in order; a first match that starts at its lane's first byte with the
        # offset of the open match ending there is absorbed into it ----
        for l in range(NL):
            if sub[l] >= n:
                break
            for i, (q, ml, off) in enumerate(lst[l]):
                if held is not None:
                    hq, hml, hoff = held
                    if merge and i == 0 and q == sub[l] and hq + hml == q and hoff == off:
                        held = (hq, hml + ml, hoff)
                        continue
                    seqs.append((hq - emitted_to, hml, hoff)); emitted_to = hq + hml; held = None
                held = (q, ml, off)
                if ml > best_ml: best_ml, rep = ml, off
            # a match that does not reach the end of its lane can no longer grow
            if held is not None and held[0] + held[1] != end[l]:
                hq, hml, hoff = held
                seqs.append((hq - emitted_to, hml, hoff)); emitted_to = hq + hml; held = None
    if held is not None:
        hq, hml, hoff = held
        seqs.append((hq - emitted_to, hml, hoff)); emitted_to = hq + hml
    seqs.append((n - emitted_to, 0, 0))
    return seqs


def encode(seqs, data):
    out = bytearray(); p = 0
    for lit, ml, off in seqs:
        m = ml - 4 if ml else 0
        out.append((min(lit, 15) << 4) | (min(m, 15) if ml else 0))
        if lit >= 15:
            r = lit - 15
            while r >= 255: out.append(255); r -= 255
            out.append(r)
        out += data[p:p + lit]; p += lit
        if ml:
            out += bytes([off & 255, off >> 8])
            if m >= 15:
                r = m - 15
                while r >= 255: out.append(255); r -= 255
                out.append(r)
            p += ml
    return bytes(out)


if __name__ == "__main__":
    ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
    from checkers import Ref, Oracle
    from trico_b200.synth import grid_mesh
    ref, orc = Ref(), Oracle()

    def run(name, plane, B=16384, **kw):
        n = len(plane); tot = 0; nseq = 0
        for b0 in range(0, n, B):
            raw = plane[b0:b0 + B].tobytes()
            sq = parse_lanes(raw, **kw)
            blk = encode(sq, raw)
            assert orc.lz4_decompress(blk, len(raw)) == raw, name
            assert orc.lz4_validate(blk, len(raw)) >= 0, name
            tot += len(blk) + 2; nseq += len(sq)
        whole = len(ref.lz4_compress(plane.tobytes()))
        print(f"{name:10s} B={B} {kw}: {n / tot:.4f} (reference, whole plane: {n / whole:.4f}) {100 * (whole / tot - 1):+.2f}% B/seq {n / nseq:.1f}", flush=True)

    bun = np.load(os.path.join(ROOT, "tests", "golden", "bunny_full.npz"))
    pl = np.ascontiguousarray(bun["triangles"].reshape(-1).view(np.uint8).reshape(-1, 4)[:, 1])
    which = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    if which == "sweep":
        for kw in (dict(S=64), dict(S=64, longest=True), dict(S=128, longest=True), dict(S=256, longest=True), dict(S=512, longest=True), dict(S=512, longest=True, HLOG=11), dict(S=512, longest=True, HLOG=12),
                   dict(S=512, longest=True, HLOG=12, merge=False), dict(S=256, longest=True, HLOG=12)):
            run("bunny p1", pl, **kw)
