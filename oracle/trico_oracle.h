/*
 * trico_oracle.h - TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference (janm31415/trico) encode/decode hot path, used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the *checker*.  Nothing in
 * the product path (trico_b200/) may include, link or call this.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle.py) against
 *   (a) the known-answer vectors and bunny-derived fixtures under tests/golden/, which were
 *       produced by the unmodified reference compiled into oracle/_ref/libtrico_ref.so
 *       (tests/golden/make_golden.py is the generating script), and
 *   (b) oracle/_ref/libtrico_ref.so itself, side by side on random inputs, whenever that
 *       library is present.
 *
 * Citations are to files under /root/reference (upstream janm31415/trico):
 *   fpc.c = trico/floating_point_stream_compression.c
 */
#ifndef TRICO_ORACLE_H
#define TRICO_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- FPC-style float / double stream codec (fpc.c:86-417, fpc.c:576-1164) ---- */

/* worst-case output size of oracle_fpc{32,64}_compress, including the 5-byte stream header */
uint64_t oracle_fpc32_bound(uint64_t n);
uint64_t oracle_fpc64_bound(uint64_t n);

/* Reference stream format: u8 hash_info, u32 BE n, groups.  Returns bytes written. */
uint64_t oracle_fpc32_compress(uint8_t* out, const uint32_t* in, uint32_t n, uint32_t e1, uint32_t e2);
uint64_t oracle_fpc64_compress(uint8_t* out, const uint64_t* in, uint32_t n, uint32_t e1, uint32_t e2);
/* Returns the value count from the header; writes that many values to out (out may be NULL to query). */
uint32_t oracle_fpc32_decompress(uint32_t* out, const uint8_t* in);
uint32_t oracle_fpc64_decompress(uint64_t* out, const uint8_t* in);
/* Returns the number of input bytes a stream occupies (walks the code words). */
uint64_t oracle_fpc32_stream_bytes(const uint8_t* in);
uint64_t oracle_fpc64_stream_bytes(const uint8_t* in);

/* ---- AoS <-> SoA and byte planes (trico/transpose_aos_to_soa.c:8-147) ---- */
void oracle_deinterleave(void* soa, const void* aos, uint64_t n, int ncomp, int wordsize); /* soa = ncomp arrays of n, back to back */
void oracle_interleave(void* aos, const void* soa, uint64_t n, int ncomp, int wordsize);
void oracle_planes_split(uint8_t* planes, const void* in, uint64_t n, int wordsize);       /* planes = wordsize arrays of n bytes, LSB plane first */
void oracle_planes_merge(void* out, const uint8_t* planes, uint64_t n, int wordsize);

/* ---- LZ4 block format (lz4/lz4.c:1657-2072 decoder; format constants lz4.c:189-196) ---- */
/* Safe decoder: returns decoded size, or -1 on malformed input / overflow. */
int64_t oracle_lz4_decompress(uint8_t* dst, uint64_t dst_cap, const uint8_t* src, uint64_t src_len);
/* Checks the end-of-block rules an LZ4 *encoder* must honour (last 5 bytes literal, last match
 * starts >= 12 bytes before the end, offsets in 1..65535 and inside the block).  Returns decoded
 * size or a negative error code. */
int64_t oracle_lz4_validate(const uint8_t* src, uint64_t src_len, uint64_t expect_raw);
/* A plain greedy single-hash compressor producing valid blocks; NOT byte-identical to LZ4 1.9.2
 * (only used to make test inputs for the GPU block decoder and as ratio yardstick). */
uint64_t oracle_lz4_bound(uint64_t n);
uint64_t oracle_lz4_compress(uint8_t* dst, const uint8_t* src, uint64_t n);

/* ---- stream layout table (trico/trico.h:11-34, trico/trico.c:215-858) ---- */
typedef struct
  {
  int codec;      /* 0 = none, 1 = FPC, 2 = LZ4 byte planes */
  int wordsize;   /* bytes per scalar: 1,2,4,8 */
  int ncomp;      /* FPC: components per element (3,2,1); LZ4: always 1 */
  int per_count;  /* scalars per counted element in each plane/component: 3 for triangle index streams, else 1 */
  } oracle_layout;
int oracle_stream_layout(int type, oracle_layout* lay);

/* ---- v0 (reference wire format) archive: README.md:251-296, trico.c:90-98 ---- */
/* Appends the 8-byte header. Returns bytes written. */
uint64_t oracle_v0_write_header(uint8_t* out, uint32_t version);
/* Appends one stream of `type` holding `count` elements from `data`; returns bytes written.
 * The caller provides room (oracle_v0_stream_bound). */
uint64_t oracle_v0_stream_bound(int type, uint32_t count);
uint64_t oracle_v0_write_stream(uint8_t* out, int type, const void* data, uint32_t count);
/* Decodes the stream starting at `in` (pointing at the type byte). Writes count*ncomp*per_count
 * scalars to out (NULL = skip). Returns bytes consumed, 0 on error. */
uint64_t oracle_v0_read_stream(void* out, const uint8_t* in, uint64_t avail, int* type, uint32_t* count);

/* ---- v1 (chunked, B200) stream body: DESIGN.md "v1 wire format" ---- */
/* chunk payload = reference FPC stream of the chunk's values minus its 5-byte header, or one
 * LZ4 block of a byte-plane slice.  These restate the *container*; the codecs are the ones above. */
uint64_t oracle_v1_stream_bound(int type, uint32_t count, int log2_chunk);
uint64_t oracle_v1_write_stream(uint8_t* out, int type, const void* data, uint32_t count, int log2_chunk, int e1, int e2);
uint64_t oracle_v1_read_stream(void* out, const uint8_t* in, uint64_t avail, int* type, uint32_t* count);

/* ---- mesh front-end: STL vertex de-duplication and triangle normals ---- */
/* iostl.c:70-138 (trico_remove_duplicate_vertices) on the facets of a binary STL file (50 bytes
 * each: normal, three corners, attribute word - iostl.c:171-186).  Writes the unique vertices in
 * (x, y, z) float order (comparator iostl.c:8-19, equality :21-26) and 3 indices per facet;
 * vertices must hold 9*ntriangles floats.  Among corners that compare equal (+0 / -0) the lowest
 * corner id supplies the bits (the reference's unstable quicksort leaves this open).
 * Returns the number of vertices. */
uint32_t oracle_stl_dedup(const uint8_t* facets, uint32_t ntriangles, float* vertices, uint32_t* triangles);
/* tools/trico_decoder/main.c:441-469 */
void oracle_triangle_normals(const float* vertices, const uint32_t* triangles, uint32_t ntriangles, float* normals);

#ifdef __cplusplus
}
#endif
#endif
