#!/usr/bin/env python
"""Stress of the lane-parallel LZ4 second pass (lz4_encode_dense_kernel) on the GPU: many kinds of
planes made of short sequences, tens of megabytes each, every one encoded twice (the two streams
must be identical: the parse may not depend on scheduling), decoded and compared.

    python tools/lanes_stress.py [millions of elements per case] [seeds]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trico_b200

M = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 24_000_000
SEEDS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = trico_b200.Device(0)


def runs(rng, n, alphabet, maxrun):
    """values from a small alphabet in runs of 1..maxrun"""
    k = n // max(1, (1 + maxrun) // 2) + 16
    vals = rng.integers(0, alphabet, k)
    lens = rng.integers(1, maxrun + 1, k)
    return np.repeat(vals, lens)[:n]


def cases(seed):
    rng = np.random.default_rng(seed)
    n = M
    yield "u8 runs 1..4 of 3 values", 17, runs(rng, n, 3, 4).astype(np.uint8)
    yield "u8 runs 1..12 of 16 values", 17, runs(rng, n, 16, 12).astype(np.uint8)
    yield "u8 two values, runs 1..3", 17, runs(rng, n, 2, 3).astype(np.uint8)
    per = rng.integers(0, 256, 37).astype(np.uint8)
    x = np.tile(per, n // 37 + 1)[:n].copy()
    hit = rng.random(n) < 0.03
    x[hit] = rng.integers(0, 256, int(hit.sum()))
    yield "u8 period 37 with 3 % noise", 17, x
    walk = np.cumsum(rng.integers(-1, 2, n)).astype(np.int64)
    yield "u16 random walk", 18, (walk & 0xffff).astype(np.uint16)
    yield "u16 heights", 18, np.clip(30000 + 3000 * np.sin(np.arange(n) * 0.001) + rng.normal(0, 600, n), 0, 65535).astype(np.uint16)
    yield "u32 ids near the diagonal", 19, (np.arange(n, dtype=np.int64) // 3 + rng.integers(-40, 41, n)).clip(0).astype(np.uint32)
    yield "u32 ids, wide jitter", 19, (np.arange(n, dtype=np.int64) // 2 + rng.integers(-3000, 3001, n)).clip(0).astype(np.uint32)
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bunny_full.npz"))
    t0 = z["triangles"].astype(np.uint32)
    nv0 = int(z["vertices"].shape[0])
    tiles = max(1, n // (3 * t0.shape[0]))
    base = int(rng.integers(0, 60000))
    t = np.concatenate([t0 + np.uint32((base + k) * nv0) for k in range(tiles)])
    yield f"bunny tiles {base}..{base + tiles}", 3, t
    yield "u64 ids near the diagonal", 20, (np.arange(n // 2, dtype=np.int64) // 3 + rng.integers(-40, 41, n // 2)).clip(0).astype(np.uint64)


bad = 0
for seed in range(SEEDS):
    for name, ty, data in cases(seed):
        data = np.ascontiguousarray(data)
        cnt = data.shape[0]
        t0 = time.perf_counter()
        s1 = dev.encode_stream(ty, data.reshape(-1), cnt)
        s2 = dev.encode_stream(ty, data.reshape(-1), cnt)
        back = dev.decode_stream(s1)
        same = bytes(s1) == bytes(s2)
        ok = back.tobytes() == data.tobytes()
        print(f"seed {seed} {name:34s} {data.nbytes / 1e6:7.1f} MB ratio {data.nbytes / len(s1):7.3f} "
              f"{'ok' if ok else 'ROUND TRIP MISMATCH'} {'deterministic' if same else 'NOT DETERMINISTIC'} ({time.perf_counter() - t0:.1f} s)", flush=True)
        bad += (not ok) + (not same)
print("FAILED" if bad else "all ok")
sys.exit(1 if bad else 0)
