#!/usr/bin/env bash
# Runs the binaries built by build_reference_clients.sh (needs a GPU): the reference's unmodified
# test suite against our library, then encoder -> decoder on the bundled bunny.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/build/refclients"
cd "$OUT"
./trico.tests | tail -n 12
./trico_encoder -i data/StanfordBunny.stl -o bunny.trc
./trico_decoder -i bunny.trc -o bunny_out.stl
ls -l data/StanfordBunny.stl bunny.trc bunny_out.stl
python - <<'PY'
import struct, sys
def tris(path):
    b = open(path, 'rb').read()
    n = struct.unpack_from('<I', b, 80)[0]
    out = []
    for i in range(n):
        f = struct.unpack_from('<12f', b, 84 + 50 * i)
        out.append(f[3:12])            # the three vertices; normals are recomputed by the decoder
    return out
a, b = tris('data/StanfordBunny.stl'), tris('bunny_out.stl')
assert len(a) == len(b), (len(a), len(b))
bad = sum(1 for x, y in zip(a, b) if x != y)
print(f"STL round trip through trico_encoder/trico_decoder on libtrico_b200: {len(a)} triangles, {bad} differing")
sys.exit(1 if bad else 0)
PY
