"""Generate the golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference for the bunny STL and
oracle/_ref/libtrico_ref.so, which oracle/Makefile compiles from the reference's own sources):

    make -C oracle all && python tests/golden/make_golden.py

Writes tests/golden/kat.json.gz (known-answer vectors, hex) and tests/golden/bunny_head.npz
(a 4096-vertex / 4096-triangle slice of the reference's StanfordBunny.stl after its own STL
de-duplication, with the reference's v0 archives of that slice) plus whole-bunny facts.
Every byte in these files was produced by reference code, none by this repository's codecs.
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
from checkers import Ref  # noqa: E402

BUNNY = "/root/reference/trico.tests/data/StanfordBunny.stl"


def special_floats():
    return np.array([0.0, -0.0, 1.0, -1.0, np.nan, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, 1.17549435e-38,
                     0.1, 0.2, 0.30000001, 123456.789, -7.5], np.float32)


def main():
    ref = Ref()
    rng = np.random.default_rng(20261018)
    kat = {"fpc32": [], "fpc64": [], "lz4": []}

    def add_fpc(values, e1, e2):
        values = np.ascontiguousarray(values)
        key = "fpc32" if values.dtype.itemsize == 4 else "fpc64"
        kat[key].append({"e1": e1, "e2": e2, "in": values.tobytes().hex(), "out": ref.compress(values, e1, e2).hex()})

    # the survey's hand vectors, regenerated
    add_fpc(np.array([1.0], np.float32), 4, 10)
    add_fpc(np.array([1, 1, 2, 3, 4, 5, 6, 7], np.float32), 4, 10)
    add_fpc(np.array([1.0], np.float64), 20, 20)
    add_fpc(np.array([1.0, 1.0], np.float64), 20, 20)
    add_fpc(special_floats(), 4, 10)
    add_fpc(special_floats().astype(np.float64), 20, 20)
    # every tail length, several exponent pairs, smooth + noisy data
    for n in list(range(1, 20)) + [31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 512, 513, 1000]:
        t = np.arange(n, dtype=np.float64)
        smooth = (np.sin(t * 0.01) * 3 + 0.001 * rng.standard_normal(n))
        for (e1, e2) in ((4, 10), (4, 4), (2, 2), (6, 8)):
            add_fpc(smooth.astype(np.float32), e1, e2)
        for (e1, e2) in ((20, 20), (4, 4), (4, 10)):
            add_fpc(smooth, e1, e2)
    add_fpc(rng.integers(0, 2**32, 300, dtype=np.uint64).astype(np.uint32).view(np.float32), 4, 10)
    add_fpc(rng.integers(0, 2**63, 300, dtype=np.uint64).view(np.float64), 20, 20)
    add_fpc(np.repeat(np.float32(2.5), 100), 4, 10)

    def add_lz4(raw: bytes):
        kat["lz4"].append({"in": raw.hex(), "out": ref.lz4_compress(raw).hex()})

    add_lz4(b"")
    add_lz4(b"abcabcabcabc")
    add_lz4(b"a" * 32)
    add_lz4(b"a" * 70000)
    add_lz4(bytes(rng.integers(0, 256, 5000, dtype=np.uint8)))
    add_lz4(bytes(rng.integers(0, 4, 5000, dtype=np.uint8)))
    add_lz4((b"0123456789abcdef" * 300) + bytes(rng.integers(0, 256, 100, dtype=np.uint8)) + b"xyz" * 1000)
    for n in (1, 4, 5, 11, 12, 13, 14, 20, 64, 65, 300):
        add_lz4(bytes(rng.integers(0, 3, n, dtype=np.uint8)))

    with gzip.open(os.path.join(HERE, "kat.json.gz"), "wt") as f:
        json.dump(kat, f)

    # ---- bunny ----
    v, t = ref.read_stl(BUNNY)
    facts = {"nv": int(v.shape[0]), "nt": int(t.shape[0])}
    full = ref.encode([(1, v, v.shape[0]), (3, t, t.shape[0])], initial=1024 * 1024)
    facts["archive_bytes"] = len(full)
    facts["archive_md5"] = hashlib.md5(full).hexdigest()
    # sub-stream sizes
    off, sizes = 8 + 5, []
    for _ in range(3):
        nb = int.from_bytes(full[off:off + 4], "little"); sizes.append(nb); off += 4 + nb
    off += 5
    for _ in range(4):
        nb = int.from_bytes(full[off:off + 4], "little"); sizes.append(nb); off += 4 + nb
    facts["substream_bytes"] = sizes
    facts["vertices_sha1"] = hashlib.sha1(v.tobytes()).hexdigest()
    facts["triangles_sha1"] = hashlib.sha1(t.tobytes()).hexdigest()

    # the whole bunny as the reference sees it, with the reference's own archive (config C1)
    np.savez_compressed(os.path.join(HERE, "bunny_full.npz"), vertices=v, triangles=t, v0_archive=np.frombuffer(full, np.uint8))

    hv, ht = v[:4096].copy(), t[20000:24096].copy()
    n = 4096
    streams = {
        1: hv, 2: hv.astype(np.float64), 3: ht, 4: ht.astype(np.uint64),
        5: hv[:, :2].copy(), 7: np.repeat(hv[:1365, :2], 3, axis=0).copy(),
        9: hv[::-1].copy(), 10: hv[::-1].astype(np.float64), 11: hv * np.float32(0.5), 12: (hv * np.float32(0.5)).astype(np.float64),
        13: (ht[:, 0] * np.uint32(2654435761)) >> np.uint32(8), 14: ht[:, 1].copy(),
        15: hv[:, 2].copy(), 16: hv[:, 1].astype(np.float64),
        17: (ht[:, 0] & 0xff).astype(np.uint8), 18: (ht[:, 0] & 0xffff).astype(np.uint16),
        19: ht[:, 2].copy(), 20: ht[:, 0].astype(np.uint64) << np.uint64(20),
    }
    out = {"vertices": hv, "triangles": ht}
    for ty, data in streams.items():
        cnt = 1365 if ty == 7 else (n if data.ndim == 1 or ty not in (7,) else n)
        cnt = 1365 if ty == 7 else data.shape[0]
        out[f"in_{ty}"] = np.ascontiguousarray(data)
        out[f"cnt_{ty}"] = np.array(cnt)
        out[f"v0_{ty}"] = np.frombuffer(ref.encode([(ty, data, cnt)]), np.uint8)
    # double-uv: the reference writes the float tags 5/7 for these (trico.c:622,:627); keep its bytes
    duv = hv[:, :2].astype(np.float64)
    out["in_6"] = duv
    out["cnt_6"] = np.array(n)
    out["v0_6_as_written_by_reference"] = np.frombuffer(ref.encode([(6, duv, n)]), np.uint8)
    # one multi-stream archive: vertices + triangles + normals + colours
    multi = ref.encode([(1, hv, n), (3, ht, n), (9, hv[::-1].copy(), n), (13, streams[13], n)])
    out["v0_multi"] = np.frombuffer(multi, np.uint8)
    np.savez_compressed(os.path.join(HERE, "bunny_head.npz"), **out)
    with open(os.path.join(HERE, "bunny_facts.json"), "w") as f:
        json.dump(facts, f, indent=1)
    print(facts)


if __name__ == "__main__":
    main()
