// stl.cuh - the mesh front-end of the tools on the GPU:
//   * binary-STL facet parse + vertex de-duplication = trico_remove_duplicate_vertices
//     (/root/reference/trico_io/iostl.c:70-138, called from trico_read_stl :141 and trico_read_stl_full :197)
//   * triangle normals as trico_decoder recomputes them (/root/reference/tools/trico_decoder/main.c:439-470)
//
// The reference sorts the 3T facet corners with a recursive quicksort under the comparator
// "x, then y, then z as floats" (iostl.c:8-19), walks the sorted list and starts a new vertex wherever
// the float comparison says "different" (iostl.c:21-26, :112-137).  Its RESULT is order independent:
// the unique vertices in lexicographic order, every corner mapped to its vertex.  Here that is
//   parse -> 16-byte records (key x, key y, key z, corner id)            stl_parse_kernel
//   LSD radix sort, 8 bits a pass, stable, passes whose digit is the same everywhere skipped
//                                                                        radix_count / scan / radix_scatter
//   head flags -> scan -> vertex ids; heads write the vertex, every corner its index
//                                                                        stl_heads_kernel / stl_emit_kernel
// Keys are the float bits mapped to unsigned order with -0 folded onto +0 (the reference compares
// floats, so the two are one coordinate); the vertex that is WRITTEN carries the original bits of the
// lowest corner of its run (the reference's unstable quicksort leaves that choice open).  NaN
// coordinates: the reference's comparator is not an order on them; here equal NaN bits are one vertex.
#pragma once
#include "common.cuh"

namespace tb200 {

constexpr int STL_FACET_BYTES = 50;
constexpr int STL_THREADS = 256;
constexpr int STL_SORT_ROUNDS = 16;                         // a sort tile = 16 rounds of 256 records
constexpr int STL_SORT_TILE = STL_THREADS * STL_SORT_ROUNDS;
constexpr int STL_SCAN_ITEMS = 16;                          // a scan chunk = 256 threads x 16 entries
constexpr int STL_SCAN_CHUNK = STL_THREADS * STL_SCAN_ITEMS;

__device__ __forceinline__ uint32_t stl_key(uint32_t bits)
  {
  if (bits == 0x80000000u) bits = 0;                        // -0 == +0 (iostl.c:10, :23)
  return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
  }
__device__ __forceinline__ uint32_t stl_digit(const uint4 r, int pass)
  { // pass 0..3: bytes of z (least significant key), 4..7: y, 8..11: x
  const uint32_t w = pass < 4 ? r.z : pass < 8 ? r.y : r.x;
  return (w >> (8 * (pass & 3))) & 255u;
  }
// a 4-byte field of a facet (facets start at byte 84 of the file and are 50 bytes long: 2-byte aligned)
__device__ __forceinline__ uint32_t stl_load_u32(const uint8_t* p, bool aligned2)
  {
  if (aligned2)
    {
    const uint16_t* q = reinterpret_cast<const uint16_t*>(p);
    return (uint32_t)q[0] | ((uint32_t)q[1] << 16);
    }
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  }

// ---------------------------------------------------------------------------------------------
// facets -> corner records (+ triangle normals and attribute words for trico_read_stl_full) and the
// twelve digit histograms of the whole input (they do not depend on the order: the host reads them
// once and drops every pass whose digit is the same in all records).
__global__ void __launch_bounds__(STL_THREADS)
stl_parse_kernel(const uint8_t* __restrict__ facets, uint32_t ntri, uint4* __restrict__ rec,
                 float* __restrict__ normals, uint16_t* __restrict__ attrs, uint32_t* __restrict__ ghist)
  {
  __shared__ uint32_t hist[12 * 256];
  __shared__ __align__(16) uint8_t tile[STL_THREADS * STL_FACET_BYTES];
  for (int i = threadIdx.x; i < 12 * 256; i += STL_THREADS) hist[i] = 0;
  const bool aligned4 = (reinterpret_cast<uintptr_t>(facets) & 3u) == 0;
  const uint32_t ntiles = (ntri + STL_THREADS - 1) / STL_THREADS;
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
    __syncthreads();
    const uint32_t f0 = t * STL_THREADS;
    const uint32_t nf = min((uint32_t)STL_THREADS, ntri - f0);
    const uint8_t* src = facets + (size_t)f0 * STL_FACET_BYTES;           // 12800 bytes a tile: as aligned as `facets`
    const uint32_t nbytes = nf * STL_FACET_BYTES;
    if (aligned4)
      {
      const uint32_t nw = nbytes >> 2;
      for (uint32_t i = threadIdx.x; i < nw; i += STL_THREADS)
        reinterpret_cast<uint32_t*>(tile)[i] = __ldcs(reinterpret_cast<const uint32_t*>(src) + i);
      for (uint32_t i = (nw << 2) + threadIdx.x; i < nbytes; i += STL_THREADS) tile[i] = src[i];
      }
    else
      for (uint32_t i = threadIdx.x; i < nbytes; i += STL_THREADS) tile[i] = src[i];
    __syncthreads();
    if (threadIdx.x < nf)
      {
      const uint32_t f = f0 + threadIdx.x;
      const uint8_t* p = tile + threadIdx.x * STL_FACET_BYTES;
      uint32_t w[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) w[i] = stl_load_u32(p + 4 * i, true);
      if (normals)
        {
        normals[(size_t)f * 3] = __uint_as_float(w[0]);
        normals[(size_t)f * 3 + 1] = __uint_as_float(w[1]);
        normals[(size_t)f * 3 + 2] = __uint_as_float(w[2]);
        }
      if (attrs) attrs[f] = *reinterpret_cast<const uint16_t*>(p + 48);
#pragma unroll
      for (int j = 0; j < 3; ++j)
        {
        const uint4 r = make_uint4(stl_key(w[3 + 3 * j]), stl_key(w[4 + 3 * j]), stl_key(w[5 + 3 * j]), f * 3u + (uint32_t)j);
        rec[(size_t)f * 3 + j] = r;
#pragma unroll
        for (int pass = 0; pass < 12; ++pass) atomicAdd(&hist[pass * 256 + stl_digit(r, pass)], 1u);
        }
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 12 * 256; i += STL_THREADS) if (hist[i]) atomicAdd(&ghist[i], hist[i]);
  }

// ---------------------------------------------------------------------------------------------
// one pass of the sort.  counts[digit * ntiles + tile]: an exclusive scan over that array (digit
// major) is where each tile's records of each digit go.
__global__ void __launch_bounds__(STL_THREADS)
radix_count_kernel(const uint4* __restrict__ in, uint64_t n, int pass, uint32_t* __restrict__ counts, uint32_t ntiles)
  {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * STL_SORT_TILE;
#pragma unroll 4
  for (int r = 0; r < STL_SORT_ROUNDS; ++r)
    {
    const uint64_t i = base + (uint64_t)r * STL_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&hist[stl_digit(in[i], pass)], 1u);
    }
  __syncthreads();
  counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
  }

// stable scatter: a record's place = start of (digit, tile) + records of that digit earlier in the tile.
// Per round of 256 consecutive records: match.any gives a lane its rank among the warp's records of
// the same digit, the warps' counts per digit sit in shared memory for one round.
__global__ void __launch_bounds__(STL_THREADS)
radix_scatter_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint64_t n, int pass,
                     const uint32_t* __restrict__ offsets, uint32_t ntiles)
  {
  constexpr int NW = STL_THREADS / 32;
  __shared__ uint32_t running[256];                         // next free place of every digit
  __shared__ uint32_t wcnt[NW][256];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  running[threadIdx.x] = offsets[(size_t)threadIdx.x * ntiles + blockIdx.x];
  for (int w = 0; w < NW; ++w) wcnt[w][threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * STL_SORT_TILE;
  for (int r = 0; r < STL_SORT_ROUNDS; ++r)
    {
    const uint64_t i = base + (uint64_t)r * STL_THREADS + threadIdx.x;
    if (base + (uint64_t)r * STL_THREADS >= n) break;       // uniform
    const bool live = i < n;
    uint4 rcd = make_uint4(0, 0, 0, 0);
    uint32_t d = 256;                                       // dead lanes match only each other
    if (live) { rcd = in[i]; d = stl_digit(rcd, pass); }
    const unsigned peers = __match_any_sync(FULL, d);
    const uint32_t rank = __popc(peers & lanemask_lt());
    if (live && rank == 0) wcnt[warp][d] = __popc(peers);
    __syncthreads();
    if (live)
      {
      uint32_t pos = running[d] + rank;
#pragma unroll
      for (int w = 0; w < NW; ++w) if (w < (int)warp) pos += wcnt[w][d];
      out[pos] = rcd;
      }
    __syncthreads();
      {
      uint32_t sum = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) { sum += wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = 0; }
      running[threadIdx.x] += sum;
      }
    __syncthreads();
    }
  }

// ---------------------------------------------------------------------------------------------
// exclusive scan of a u32 array in place, three small kernels: chunks of 4096 entries, the chunk
// totals (one CTA), the add.  Totals stay below 2^32 (3T corners, the reference's own corner ids are u32).
__global__ void __launch_bounds__(STL_THREADS)
scan_chunks_kernel(uint32_t* __restrict__ a, uint64_t n, uint32_t* __restrict__ sums)
  {
  __shared__ uint32_t wsum[STL_THREADS / 32];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const uint64_t base = (uint64_t)blockIdx.x * STL_SCAN_CHUNK + (uint64_t)threadIdx.x * STL_SCAN_ITEMS;
  uint32_t v[STL_SCAN_ITEMS];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < STL_SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? a[base + k] : 0u; mine += v[k]; }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t up = __shfl_up_sync(FULL, incl, o); if (lane >= (unsigned)o) incl += up; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t before = incl - mine, total = 0;
#pragma unroll
  for (int w = 0; w < STL_THREADS / 32; ++w) { const uint32_t t = wsum[w]; if (w < (int)warp) before += t; total += t; }
#pragma unroll
  for (int k = 0; k < STL_SCAN_ITEMS; ++k) { if (base + k < n) a[base + k] = before; before += v[k]; }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
  }

__global__ void __launch_bounds__(1024)
scan_sums_kernel(uint32_t* __restrict__ sums, uint32_t nsums)
  {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t b = 0; b < nsums; b += 1024)
    {
    const uint32_t i = b + threadIdx.x;
    const uint32_t mine = i < nsums ? sums[i] : 0u;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t up = __shfl_up_sync(FULL, incl, o); if (lane >= (unsigned)o) incl += up; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t before = carry_s + incl - mine, total = 0;
    for (int w = 0; w < 32; ++w) { const uint32_t t = wsum[w]; if (w < (int)warp) before += t; total += t; }
    if (i < nsums) sums[i] = before;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
    }
  }

__global__ void __launch_bounds__(STL_THREADS)
scan_add_kernel(uint32_t* __restrict__ a, uint64_t n, const uint32_t* __restrict__ sums)
  {
  const uint32_t add = sums[blockIdx.x];
  if (!add) return;
  const uint64_t base = (uint64_t)blockIdx.x * STL_SCAN_CHUNK;
#pragma unroll 4
  for (int k = 0; k < STL_SCAN_ITEMS; ++k)
    {
    const uint64_t i = base + (uint64_t)k * STL_THREADS + threadIdx.x;
    if (i < n) a[i] += add;
    }
  }

// ---------------------------------------------------------------------------------------------
// sorted records -> 1 where a new vertex starts (iostl.c:112-114: the walk compares neighbours)
__device__ __forceinline__ bool stl_same_vertex(const uint4 a, const uint4 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

__global__ void __launch_bounds__(STL_THREADS)
stl_heads_kernel(const uint4* __restrict__ rec, uint64_t n, uint32_t* __restrict__ flags)
  {
  const uint64_t i = (uint64_t)blockIdx.x * STL_THREADS + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i == 0 || !stl_same_vertex(rec[i], rec[i - 1])) ? 1u : 0u;
  }

// heads write their vertex (original bits, fetched from the facet of the run's lowest corner - the
// sort is stable), every corner writes its vertex index (iostl.c:107-137)
__global__ void __launch_bounds__(STL_THREADS)
stl_emit_kernel(const uint4* __restrict__ rec, uint64_t n, const uint32_t* __restrict__ before,
                const uint8_t* __restrict__ facets, float* __restrict__ vertices, uint32_t* __restrict__ triangles,
                uint32_t* __restrict__ nvertices)
  {
  const uint64_t i = (uint64_t)blockIdx.x * STL_THREADS + threadIdx.x;
  if (i >= n) return;
  const uint4 r = rec[i];
  const bool head = i == 0 || !stl_same_vertex(r, rec[i - 1]);
  const uint32_t vid = before[i] + (head ? 1u : 0u) - 1u;
  triangles[r.w] = vid;
  if (head)
    {
    const uint32_t f = r.w / 3u, j = r.w - 3u * f;
    const uint8_t* p = facets + (size_t)f * STL_FACET_BYTES + 12 + 12 * j;
    const bool a2 = (reinterpret_cast<uintptr_t>(p) & 1u) == 0;
    vertices[(size_t)vid * 3] = __uint_as_float(stl_load_u32(p, a2));
    vertices[(size_t)vid * 3 + 1] = __uint_as_float(stl_load_u32(p + 4, a2));
    vertices[(size_t)vid * 3 + 2] = __uint_as_float(stl_load_u32(p + 8, a2));
    }
  if (i == n - 1) *nvertices = vid + 1u;
  }

// ---------------------------------------------------------------------------------------------
// triangle normals, operation for operation as trico_decoder computes them (main.c:441-469):
// every product and sum rounded on its own (the reference is compiled for baseline x86-64: no fused
// multiply-add), the length as the float square root (main.c:465 takes the double root of a float:
// the same value), a zero length leaves the vector as it is.
__global__ void __launch_bounds__(STL_THREADS)
triangle_normals_kernel(const float* __restrict__ vertices, const uint32_t* __restrict__ triangles, uint32_t ntri,
                        float* __restrict__ normals)
  {
  const uint32_t t = blockIdx.x * STL_THREADS + threadIdx.x;
  if (t >= ntri) return;
  const uint32_t v0 = triangles[(size_t)t * 3], v1 = triangles[(size_t)t * 3 + 1], v2 = triangles[(size_t)t * 3 + 2];
  const float x0 = vertices[(size_t)v0 * 3], y0 = vertices[(size_t)v0 * 3 + 1], z0 = vertices[(size_t)v0 * 3 + 2];
  const float x1 = vertices[(size_t)v1 * 3], y1 = vertices[(size_t)v1 * 3 + 1], z1 = vertices[(size_t)v1 * 3 + 2];
  const float x2 = vertices[(size_t)v2 * 3], y2 = vertices[(size_t)v2 * 3 + 1], z2 = vertices[(size_t)v2 * 3 + 2];
  const float ax = __fsub_rn(x1, x0), ay = __fsub_rn(y1, y0), az = __fsub_rn(z1, z0);
  const float bx = __fsub_rn(x2, x0), by = __fsub_rn(y2, y0), bz = __fsub_rn(z2, z0);
  const float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
  const float ny = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
  const float nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
  const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
  const bool nz_len = len != 0.0f;
  normals[(size_t)t * 3] = nz_len ? __fdiv_rn(nx, len) : nx;
  normals[(size_t)t * 3 + 1] = nz_len ? __fdiv_rn(ny, len) : ny;
  normals[(size_t)t * 3 + 2] = nz_len ? __fdiv_rn(nz, len) : nz;
  }

// ---------------------------------------------------------------------------------------------
// indexed mesh -> the 50-byte facet records of a binary STL file (trico_write_stl, iostl.c:261-320):
// normal (zeros when there are none), the three corner positions, the attribute word (zero when none)
__global__ void __launch_bounds__(STL_THREADS)
stl_facets_kernel(const float* __restrict__ vertices, const uint32_t* __restrict__ triangles, uint32_t ntri,
                  const float* __restrict__ normals, const uint16_t* __restrict__ attrs, uint8_t* __restrict__ out)
  {
  const uint32_t t = blockIdx.x * STL_THREADS + threadIdx.x;
  if (t >= ntri) return;
  uint32_t w[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) w[i] = normals ? __float_as_uint(normals[(size_t)t * 3 + i]) : 0u;
#pragma unroll
  for (int j = 0; j < 3; ++j)
    {
    const uint32_t v = triangles[(size_t)t * 3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[3 + 3 * j + i] = __float_as_uint(vertices[(size_t)v * 3 + i]);
    }
  uint16_t* q = reinterpret_cast<uint16_t*>(out + (size_t)t * STL_FACET_BYTES);      // 50 t: 2-byte aligned
#pragma unroll
  for (int i = 0; i < 12; ++i) { q[2 * i] = (uint16_t)w[i]; q[2 * i + 1] = (uint16_t)(w[i] >> 16); }
  q[24] = attrs ? attrs[t] : (uint16_t)0;
  }

}  // namespace tb200
