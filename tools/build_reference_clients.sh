#!/usr/bin/env bash
# Builds the reference's OWN clients - trico_encoder, trico_decoder and its trico.tests suite -
# UNMODIFIED, from the sources where they lie under $REF, against libtrico_b200.so.
# This is the drop-in acceptance check of SURVEY.md 4/8b: the reference's callers compile against
# include/trico/*.h (our shims) and link against our library instead of the reference's libtrico.
# Only possible where /root/reference is mounted; outputs go to build/refclients (git-ignored,
# but carried to the GPU box by gpurun, where `run_reference_clients.sh` executes them).
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${REF:-/root/reference}"
OUT="$ROOT/build/refclients"
LIBDIR="$ROOT/trico_b200/lib"
[ -d "$REF/trico.tests" ] || { echo "reference sources not found at $REF"; exit 1; }
[ -f "$LIBDIR/libtrico_b200.so" ] || python -m trico_b200.build
mkdir -p "$OUT/data"
cp "$REF/trico.tests/data/StanfordBunny.stl" "$OUT/data/"
INC=(-I"$ROOT/include" -I"$REF")          # our trico/*.h shims shadow the reference's headers
CFLAGS=(-O2 -w)
# pieces of the reference that are NOT on the hot path and stay CPU code: file I/O, PLY parser, and
# the CPU lz4 that trico.tests/int_compression.cpp calls directly
for f in trico_io/iostl.c trico_io/ioply.c rply/rply.c lz4/lz4.c; do
  gcc "${CFLAGS[@]}" "${INC[@]}" -c "$REF/$f" -o "$OUT/$(basename "${f%.c}").o"
done
AUX=("$OUT/iostl.o" "$OUT/ioply.o" "$OUT/rply.o" "$OUT/lz4.o")
LINK=(-L"$LIBDIR" -ltrico_b200 -Wl,-rpath,"$LIBDIR" -lm)
gcc "${CFLAGS[@]}" "${INC[@]}" "$REF/tools/trico_encoder/main.c" "${AUX[@]}" -o "$OUT/trico_encoder" "${LINK[@]}"
gcc "${CFLAGS[@]}" "${INC[@]}" "$REF/tools/trico_decoder/main.c" "${AUX[@]}" -o "$OUT/trico_decoder" "${LINK[@]}"
# the same two tools with the STL front-end taken from OUR library as well (trico_read_stl, trico_read_stl_full,
# trico_write_stl: include/trico_b200_io.h) - the reference's iostl.c is simply left out of the link
AUX_GPUIO=("$OUT/ioply.o" "$OUT/rply.o" "$OUT/lz4.o")
gcc "${CFLAGS[@]}" "${INC[@]}" "$REF/tools/trico_encoder/main.c" "${AUX_GPUIO[@]}" -o "$OUT/trico_encoder_gpuio" "${LINK[@]}"
gcc "${CFLAGS[@]}" "${INC[@]}" "$REF/tools/trico_decoder/main.c" "${AUX_GPUIO[@]}" -o "$OUT/trico_decoder_gpuio" "${LINK[@]}"
g++ -std=c++17 "${CFLAGS[@]}" "${INC[@]}" -I"$REF/trico.tests" "$REF"/trico.tests/*.cpp "${AUX[@]}" -o "$OUT/trico.tests" "${LINK[@]}"
echo "built: $OUT/trico_encoder $OUT/trico_decoder $OUT/trico.tests"
