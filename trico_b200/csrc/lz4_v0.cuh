// lz4_v0.cuh - reference-format (v0) LZ4 planes from the GPU: ONE LZ4 block per whole byte plane
// (trico.c:323-378: `LZ4_compress_default` over the plane, u32 size in front), built from the
// independent per-block results of K5 / K5b.
//
// An LZ4 block is a list of sequences (literals + match) that ends with a literals-only sequence
// (lz4.c:189-196).  Offsets are distances, so the sequences of independently compressed blocks
// stay valid when the blocks are laid end to end - except for every block's closing literals-only
// sequence, which may only be the last thing in a block.  Those literals are moved into the NEXT
// sequence that has a match: its literal run grows by the carried bytes (a new token and new
// length bytes), everything behind its literal run is untouched.  A block without any match (a
// stored noise block) is all carry.
//   1. lz4_v0_meta_kernel   thread per chunk: walks its sequences (first literal run, closing run)
//   2. lz4_v0_scan_kernel   warp per plane: the carried literal count in front of every chunk (a scan
//                           over "constant or add" functions), the bytes every chunk's sequences
//                           take in the merged block, their offsets, the base of every literal
//                           region; writes the closing sequence's header and the plane's size
//   3. lz4_v0_place_kernel  warp per chunk: new header of the first sequence, the chunk's body, its
//                           closing literals (into the literal region of the next sequence)
// The merged block decodes with LZ4_decompress_safe (lz4.c:2078) to the plane, which is all the
// reference's reader asks of it (trico.c:1101).
#pragma once

#include "lz4.cuh"

namespace tb200 {

struct Lz4V0Meta
  {
  uint16_t tail;           // literals of the closing sequence (== cnt for a chunk without a match)
  uint16_t first;          // literals of the first sequence
  uint8_t h_first;         // bytes of the first sequence's token + literal length bytes
  uint8_t h_tail;          // the same of the closing sequence
  uint8_t has_match;
  uint8_t pad;
  };

struct Lz4V0Chunk
  {
  uint64_t out_start;      // has_match: where the chunk's (re-headed) first sequence starts
  uint32_t carry_in;       // literals carried into the chunk's first sequence / offset of its bytes in the literal region
  uint32_t region;         // index of the literal region the chunk's first sequence owns (= matches before it)
  };

struct Lz4V0Args
  {
  const uint8_t* scratch;  // chunk g = range * planes + plane at scratch + g * slot
  uint32_t slot;
  const uint8_t* sizes;    // u16 LE per chunk
  uint64_t n;              // elements (= bytes per plane)
  uint32_t nranges;
  int log2B;
  int planes;
  Lz4V0Meta* meta;         // [nranges * planes]
  Lz4V0Chunk* chunk;       // [nranges * planes]
  uint64_t* region_base;   // [planes][nranges + 1]
  uint8_t* out;            // plane p at out + p * out_stride
  uint64_t out_stride;
  uint64_t* nbytes;        // [planes]
  };

__device__ __forceinline__ uint32_t lz4_lit_header_bytes(uint32_t lit) { return 1u + (lit >= 15u ? (lit - 15u) / 255u + 1u : 0u); }

__global__ void __launch_bounds__(128)
lz4_v0_meta_kernel(const Lz4V0Args a)
  {
  const uint64_t nchunks = (uint64_t)a.nranges * a.planes;
  for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < nchunks; g += (uint64_t)gridDim.x * blockDim.x)
    {
    const uint8_t* b = a.scratch + g * a.slot;
    const uint32_t size = (uint32_t)a.sizes[2 * g] | ((uint32_t)a.sizes[2 * g + 1] << 8);
    Lz4V0Meta m; m.tail = 0; m.first = 0; m.h_first = 0; m.h_tail = 0; m.has_match = 0; m.pad = 0;
    uint32_t ip = 0, nseq = 0;
    while (ip < size)
      {
      const uint32_t tok = __ldcg(b + ip);
      uint32_t lit = tok >> 4, p = ip + 1u;
      if (lit == 15u) { uint32_t x; do { x = __ldcg(b + p); ++p; lit += x; } while (x == 255u && p < size); }
      if (nseq == 0) { m.first = (uint16_t)lit; m.h_first = (uint8_t)(p - ip); }
      const uint32_t hdr = p - ip;
      p += lit;
      if (p >= size) { m.tail = (uint16_t)lit; m.h_tail = (uint8_t)hdr; break; }      // the closing sequence
      p += 2u;
      if ((tok & 15u) == 15u) { uint32_t x; do { x = __ldcg(b + p); ++p; } while (x == 255u && p < size); }
      ip = p; ++nseq;
      }
    m.has_match = nseq != 0;
    a.meta[g] = m;
    }
  }

// carry functions: (is_const, val): x -> is_const ? val : x + val; `b after a`
__device__ __forceinline__ void lz4_v0_compose(uint32_t& ac, uint32_t& av, uint32_t bc, uint32_t bv)
  {
  if (bc) { ac = 1u; av = bv; } else av += bv;
  }

__global__ void __launch_bounds__(32)
lz4_v0_scan_kernel(const Lz4V0Args a)
  {
  const unsigned lane = lane_id();
  const uint32_t p = blockIdx.x;
  const uint32_t B = 1u << a.log2B;
  uint64_t* rbase = a.region_base + (size_t)p * (a.nranges + 1u);
  uint8_t* out = a.out + (size_t)p * a.out_stride;
  uint32_t carry = 0;            // literals pending in front of the step
  uint64_t at = 0;               // bytes of the merged block so far
  uint32_t regions = 0;          // matches so far
  // the loads of a step are requested one step ahead (a step is a dozen shuffles: the round trip to
  // memory would be all of its time)
  auto fetch = [&](uint32_t k, Lz4V0Meta& m, uint32_t& size)
    {
    m.tail = 0; m.first = 0; m.h_first = 0; m.h_tail = 0; m.has_match = 0; m.pad = 0; size = 0;
    if (k < a.nranges)
      {
      const uint64_t g = (uint64_t)k * a.planes + p;
      m = a.meta[g];
      size = (uint32_t)a.sizes[2 * g] | ((uint32_t)a.sizes[2 * g + 1] << 8);
      }
    };
  Lz4V0Meta m_next; uint32_t size_next;
  fetch(lane, m_next, size_next);
  for (uint32_t k0 = 0; k0 < a.nranges; k0 += 32)
    {
    const uint32_t k = k0 + lane;
    const bool act = k < a.nranges;
    const uint64_t g = (uint64_t)k * a.planes + p;
    const Lz4V0Meta m = m_next;
    const uint32_t size = size_next;
    fetch(k + 32u, m_next, size_next);
    uint32_t cnt = 0;
    if (act)
      {
      const uint64_t lo = (uint64_t)k << a.log2B;
      cnt = (uint32_t)(a.n - lo < B ? a.n - lo : B);
      }
    // carry behind every chunk: inclusive scan of the chunks' functions
    uint32_t fc = act && m.has_match ? 1u : 0u, fv = act ? (m.has_match ? (uint32_t)m.tail : cnt) : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
      {
      const uint32_t pc = __shfl_up_sync(FULL, fc, d), pv = __shfl_up_sync(FULL, fv, d);
      if (lane >= (unsigned)d)
        { // (prev ... ) then (mine)
        uint32_t nc = pc, nv = pv;
        lz4_v0_compose(nc, nv, fc, fv);
        fc = nc; fv = nv;
        }
      }
    const uint32_t carry_out = fc ? fv : carry + fv;
    uint32_t carry_in = __shfl_up_sync(FULL, carry_out, 1);
    if (lane == 0) carry_in = carry;
    // bytes of the chunk's sequences in the merged block (its closing literals belong to the next region)
    const uint32_t hdr = lz4_lit_header_bytes(carry_in + m.first);
    const uint64_t e = act && m.has_match ? (uint64_t)hdr + carry_in + (size - m.h_first - m.h_tail - m.tail) : 0ull;
    uint64_t incl = e;
    uint32_t rinc = act && m.has_match ? 1u : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
      {
      const uint64_t pe = __shfl_up_sync(FULL, incl, d);
      const uint32_t pr = __shfl_up_sync(FULL, rinc, d);
      if (lane >= (unsigned)d) { incl += pe; rinc += pr; }
      }
    if (act)
      {
      Lz4V0Chunk c;
      c.out_start = at + incl - e;
      c.carry_in = carry_in;
      c.region = regions + rinc - (m.has_match ? 1u : 0u);
      a.chunk[g] = c;
      if (m.has_match) rbase[c.region] = c.out_start + hdr;
      }
    at += __shfl_sync(FULL, incl, 31);
    regions += __shfl_sync(FULL, rinc, 31);
    carry = __shfl_sync(FULL, carry_out, 31);
    }
  // the closing sequence: everything still carried, literals only
  const uint32_t hdr = lz4_lit_header_bytes(carry);
  if (lane == 0)
    {
    rbase[regions] = at + hdr;
    a.nbytes[p] = at + hdr + carry;
    out[at] = (uint8_t)((carry >= 15u ? 15u : carry) << 4);
    }
  if (carry >= 15u)
    {
    const uint32_t next = hdr - 1u;
    for (uint32_t i = lane; i < next; i += 32) out[at + 1u + i] = (i + 1u == next) ? (uint8_t)((carry - 15u) % 255u) : (uint8_t)255;
    }
  }

__device__ __forceinline__ void lz4_v0_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n)
  { // warp copy, any alignment on both sides: bytes up to the destination's 16-byte boundary, then 16-byte
    // stores assembled from five aligned source words each (the source may read up to 7 bytes past its
    // end: the scratch slots have that slack)
  const unsigned lane = lane_id();
  uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
  if (head > n) head = n;
  if (lane < head) dst[lane] = __ldcg(src + lane);
  const uint32_t nvec = (n - head) >> 4;
  const uint8_t* s = src + head;
  const uint32_t sh = (uint32_t)((uintptr_t)s & 3u);
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(s - sh);
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
  for (uint32_t i = lane; i < nvec; i += 32)
    {
    const uint32_t* q = sw + 4u * i;
    const uint32_t w0 = __ldcg(q), w1 = __ldcg(q + 1), w2 = __ldcg(q + 2), w3 = __ldcg(q + 3), w4 = __ldcg(q + 4);
    dv[i] = make_uint4(__funnelshift_r(w0, w1, 8u * sh), __funnelshift_r(w1, w2, 8u * sh), __funnelshift_r(w2, w3, 8u * sh), __funnelshift_r(w3, w4, 8u * sh));
    }
  const uint32_t done = head + (nvec << 4);
  if (done + lane < n) dst[done + lane] = __ldcg(src + done + lane);
  }

__global__ void __launch_bounds__(256)
lz4_v0_place_kernel(const Lz4V0Args a)
  {
  const unsigned lane = lane_id();
  const uint64_t nchunks = (uint64_t)a.nranges * a.planes;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < nchunks; g += warps)
    {
    const uint32_t p = (uint32_t)(g % a.planes);
    const uint8_t* b = a.scratch + g * a.slot;
    const uint32_t size = (uint32_t)a.sizes[2 * g] | ((uint32_t)a.sizes[2 * g + 1] << 8);
    const Lz4V0Meta m = a.meta[g];
    const Lz4V0Chunk c = a.chunk[g];
    const uint64_t* rbase = a.region_base + (size_t)p * (a.nranges + 1u);
    uint8_t* out = a.out + (size_t)p * a.out_stride;
    if (m.has_match)
      {
      // the first sequence with its longer literal run: token (match nibble kept), length bytes
      const uint32_t lit = c.carry_in + m.first;
      const uint32_t hdr = lz4_lit_header_bytes(lit);
      uint8_t* h = out + c.out_start;
      if (lane == 0) h[0] = (uint8_t)(((lit >= 15u ? 15u : lit) << 4) | (__ldcg(b) & 15u));
      for (uint32_t i = lane; i + 1u < hdr; i += 32) h[1u + i] = (i + 2u == hdr) ? (uint8_t)((lit - 15u) % 255u) : (uint8_t)255;
      // its own bytes from the first literal on, up to the closing sequence
      const uint32_t body = size - m.h_first - m.h_tail - m.tail;
      lz4_v0_copy(out + rbase[c.region] + c.carry_in, b + m.h_first, body);
      // the closing literals open the next literal region
      lz4_v0_copy(out + rbase[c.region + 1u], b + size - m.tail, m.tail);
      }
    else
      lz4_v0_copy(out + rbase[c.region] + c.carry_in, b + m.h_first, m.tail);       // all literals
    }
  }

} // namespace tb200
