// lz4.cuh - LZ4 *block format* compressor / decompressor for sm_100a, one warp per block.
//
// Replaces the bundled LZ4 v1.9.2 calls of the reference:
//   LZ4_compress_default  /root/reference/lz4/lz4.c:1271 (LZ4_compress_generic :793-1181)
//   LZ4_decompress_safe   /root/reference/lz4/lz4.c:2078 (LZ4_decompress_generic :1657-2072)
// The compressor is a greedy single-pass hash matcher like the reference's, but tests 32
// consecutive positions per step (one per lane) and inserts every scanned position; its output is
// a valid LZ4 block (last 5 bytes literal, last match starts >= 12 bytes before the end,
// lz4.c:189-196) but not byte-identical to the CPU library - the format, not the parse, is the
// contract.  The decompressor accepts any valid block, including the reference's.
#pragma once

#include "common.cuh"

namespace tb200 {

constexpr uint32_t LZ4_MINMATCH = 4;
constexpr uint32_t LZ4_LASTLITERALS = 5;     // lz4.c:192
constexpr uint32_t LZ4_MFLIMIT = 12;         // lz4.c:193

__host__ __device__ constexpr uint32_t lz4_block_bound(uint32_t n) { return n + n / 255u + 16u; }   // lz4.h:171

// unaligned little-endian 32-bit read from shared memory
__device__ __forceinline__ uint32_t smem_read32(const uint8_t* base, uint32_t pos)
  {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (pos >> 2);
  return __funnelshift_r(w[0], w[1], (pos & 3u) * 8u);
  }

// 16 bytes at an arbitrary shared-memory position (the word after the last byte must be readable)
__device__ __forceinline__ uint4 smem_read128(const uint8_t* base, uint32_t pos)
  {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (pos >> 2);
  const unsigned sh = (pos & 3u) * 8u;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
  }

// The same for positions whose low four bits are equal in every lane of the warp (lanes 16 bytes
// apart) in a 16-byte aligned buffer: two aligned 16-byte reads - no bank conflicts, where the five
// word reads above collide four ways - and the byte shift resolved by a warp-uniform branch.  Up to
// 15 bytes past the vector must be readable.
__device__ __forceinline__ uint4 smem_read128u(const uint8_t* base, uint32_t pos)
  {
  const uint4* v = reinterpret_cast<const uint4*>(base + (pos & ~15u));
  const uint4 a = v[0];
  const uint32_t m = pos & 15u;
  if (m == 0) return a;
  const uint4 b = v[1];
  const unsigned sh = (m & 3u) * 8u;
  switch (m >> 2)
    {
    case 0: return make_uint4(__funnelshift_r(a.x, a.y, sh), __funnelshift_r(a.y, a.z, sh), __funnelshift_r(a.z, a.w, sh), __funnelshift_r(a.w, b.x, sh));
    case 1: return make_uint4(__funnelshift_r(a.y, a.z, sh), __funnelshift_r(a.z, a.w, sh), __funnelshift_r(a.w, b.x, sh), __funnelshift_r(b.x, b.y, sh));
    case 2: return make_uint4(__funnelshift_r(a.z, a.w, sh), __funnelshift_r(a.w, b.x, sh), __funnelshift_r(b.x, b.y, sh), __funnelshift_r(b.y, b.z, sh));
    default: return make_uint4(__funnelshift_r(a.w, b.x, sh), __funnelshift_r(b.x, b.y, sh), __funnelshift_r(b.y, b.z, sh), __funnelshift_r(b.z, b.w, sh));
    }
  }

// warp copy of n literal bytes src[s..s+n) (shared memory) to dst (any alignment, any space);
// long runs move as 16-byte vectors (512 bytes per warp instruction), four in flight per lane
template <typename DstPtr>
__device__ __forceinline__ void lz4_copy_from_smem(DstPtr dst, const uint8_t* src, uint32_t s, uint32_t n)
  {
  const unsigned lane = lane_id();
  if (n <= 64)
    {
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[s + i];
    return;
    }
  const uint32_t head = (16u - ((uint32_t)reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (lane < head) dst[lane] = src[s + lane];
  const uint32_t nv = (n - head) >> 4;
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
  constexpr int UN = 4;
  for (uint32_t i0 = 0; i0 < nv; i0 += 32 * UN)
    {
    uint4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u)
      {
      const uint32_t i = i0 + lane + 32 * u;
      if (i < nv) v[u] = smem_read128u(src, s + head + 16u * i);
      }
#pragma unroll
    for (int u = 0; u < UN; ++u)
      {
      const uint32_t i = i0 + lane + 32 * u;
      if (i < nv) dv[i] = v[u];
      }
    }
  const uint32_t done = head + (nv << 4);
  if (done + lane < n) dst[done + lane] = src[s + done + lane];
  }

// Emits one sequence (warp-cooperative).  `mlen` = 0 means "last sequence, literals only".
// Returns the new output position.
template <typename DstPtr>
__device__ __forceinline__ uint32_t lz4_emit(DstPtr dst, uint32_t op, const uint8_t* src, uint32_t lit_start,
                                             uint32_t nlit, uint32_t offset, uint32_t mlen)
  {
  const unsigned lane = lane_id();
  const uint32_t m = mlen ? mlen - LZ4_MINMATCH : 0;
  const uint32_t next = nlit >= 15 ? (nlit - 15) / 255 + 1 : 0;       // literal-length extension bytes
    { // a sequence of at most 32 bytes (short literal run, match length continuation included) leaves
      // as one warp store: lane i produces byte i
    const uint32_t mext = (mlen && m >= 15) ? (m - 15) / 255 + 1 : 0;
    const uint32_t total = 1u + next + nlit + (mlen ? 2u + mext : 0u);
    if (total <= 32u)
      {
      const uint32_t l0 = 1u + next, o0 = l0 + nlit;                  // index of the first literal, of the offset's low byte
      uint32_t byte = ((nlit >= 15 ? 15u : nlit) << 4) | (m >= 15 ? 15u : m);
      if (lane >= 1u && lane < l0) byte = (lane == next) ? (nlit - 15) % 255 : 255u;
      if (lane >= l0 && lane < o0) byte = src[lit_start + lane - l0];
      if (lane == o0) byte = offset & 0xffu;
      if (lane == o0 + 1u) byte = offset >> 8;
      if (lane >= o0 + 2u) byte = (lane + 1u == total) ? (m - 15) % 255 : 255u;
      if (lane < total) dst[op + lane] = (uint8_t)byte;
      return op + total;
      }
    }
  if (lane == 0) dst[op] = (uint8_t)(((nlit >= 15 ? 15u : nlit) << 4) | (m >= 15 ? 15u : m));
  for (uint32_t i = lane; i < next; i += 32) dst[op + 1 + i] = (i + 1 == next) ? (uint8_t)((nlit - 15) % 255) : (uint8_t)255;
  op += 1 + next;
  lz4_copy_from_smem(dst + op, src, lit_start, nlit);
  op += nlit;
  if (mlen)
    {
    const uint32_t mext = m >= 15 ? (m - 15) / 255 + 1 : 0;
    if (lane == 0) { dst[op] = (uint8_t)offset; dst[op + 1] = (uint8_t)(offset >> 8); }
    for (uint32_t i = lane; i < mext; i += 32) dst[op + 2 + i] = (i + 1 == mext) ? (uint8_t)((m - 15) % 255) : (uint8_t)255;
    op += 2 + mext;
    }
  return op;
  }

// The block the matcher below produces for n >= 25 copies of byte v: one literal, a match at
// offset 1 up to the last five bytes, five literals (lz4.c:189-196).  Returns its size.
template <typename DstPtr>
__device__ __forceinline__ uint32_t lz4_emit_run(DstPtr dst, uint32_t n, uint32_t v)
  {
  const unsigned lane = lane_id();
  const uint32_t m = n - 10u;                           // match length n - 6, minus LZ4_MINMATCH
  const uint32_t mext = (m - 15u) / 255u + 1u;          // length continuation bytes
  const uint32_t total = 4u + mext + 6u;
  for (uint32_t i = lane; i < total; i += 32)
    {
    uint32_t byte;
    if (i == 0) byte = 0x1fu;                           // 1 literal, match length continued
    else if (i == 1) byte = v;
    else if (i == 2) byte = 1u;                         // offset 1, little-endian
    else if (i == 3) byte = 0u;
    else if (i < 4u + mext) byte = (i + 1u == 4u + mext) ? (m - 15u) % 255u : 255u;
    else if (i == 4u + mext) byte = 0x50u;              // last sequence: 5 literals
    else byte = v;
    dst[i] = (uint8_t)byte;
    }
  return total;
  }

#ifndef TB200_LZ4_DENSE_MIN
#define TB200_LZ4_DENSE_MIN 512u       // bytes of a block that are parsed before the decision
#endif
#ifndef TB200_LZ4_DENSE_SEQ
#define TB200_LZ4_DENSE_SEQ 24u        // bytes per sequence below which a block counts as dense (0: never)
#endif
#ifndef TB200_LZ4_DENSE_OUT
#define TB200_LZ4_DENSE_OUT 9u         // ... while the output stays above TB200_LZ4_DENSE_OUT tenths of the input
#endif
constexpr uint32_t LZ4_SRC_PAD = 576;    // zeroed bytes the compressor may read past the block end (512-byte match extension steps)
// A block that saves less than 1/64 of its bytes is written STORED (one literal run): noise planes
// (the low bytes of colours, heights, jittered indices) otherwise come out as a few hundred useless
// sequences that cost the decoder a thousand cycles each - and a stored plane needs no decoder at
// all (K6M reads it straight from the payload).  Costs at most 1.6 % of ratio on such a plane.
__host__ __device__ __forceinline__ bool lz4_not_worth(uint32_t nbytes, uint32_t cnt)
  {
  return cnt >= 64u && nbytes + (cnt >> 6) >= cnt;
  }
__host__ __device__ __forceinline__ uint32_t lz4_literal_run_bytes(uint32_t nlit)
  { // token, length bytes, literals
  return 1u + (nlit >= 15u ? (nlit - 15u) / 255u + 1u : 0u) + nlit;
  }

constexpr uint32_t LZ4_ENC_STAGE = 128;  // per-warp staging bytes for the sequences of one window (lz4_compress_warp)
constexpr uint32_t LZ4_WIN_CAP = 36;     // match lengths are measured up to this inside a window; longer ones take the warp-wide extension

// Compresses src[0..n) (shared memory, LZ4_SRC_PAD zero bytes readable past n) into dst.
// `table` = (1 << HLOG) u16 entries, `stage` = LZ4_ENC_STAGE bytes of shared memory private to
// the warp.  n <= 65535.
//
// Greedy single-pass matcher in the spirit of LZ4_compress_generic (lz4.c:793-1181); every lane
// tests one position per step.  Like the reference (lz4.c:879-909) the scan accelerates over data
// that does not match: after every 64 failed attempts the distance between tested positions grows.
//
// Dense scan (stride 1: 32 consecutive positions per step).  A step does not stop at its first
// match: every lane measures its own match (two candidates: the hash table's and the nearest
// earlier lane of the window holding the same four bytes - the longer one wins) up to LZ4_WIN_CAP
// bytes, a walk over the window picks the matches a left-to-right parse takes (with one position of
// look-ahead: a longer match starting one byte later is preferred, as in lz4hc's lazy evaluation),
// and ALL of them are emitted by their own lanes in one go through `stage`.  Data made of many
// short sequences (real index planes, colour planes, attribute lists) thus costs one step per 32+
// input bytes instead of one step per sequence.  A first match that reaches the cap (or sits
// behind a long literal run) takes the warp-cooperative path: backward extension, 128 + 512 bytes
// per step forward extension, one-store emission.
// Table inserts follow the reference: literal positions and match starts, not match interiors
// (lz4.c:1118 adds one position near the end of a match) - the interior positions of a run would
// replace the candidate at the run's start, which is the one that extends across the run.
// `handover`: a block that averages less than LZ4_HANDOVER_SEQ bytes per sequence once
// LZ4_HANDOVER_MIN bytes have been parsed is not finished: the function returns 0xffffffff and the
// caller passes the block to the lane-parallel parser (lz4_lanes.cuh).
constexpr uint32_t LZ4_HANDOVER_MIN = 1024, LZ4_HANDOVER_SEQ = 16;
template <int HLOG, typename DstPtr>
__device__ __forceinline__ uint32_t lz4_compress_warp(const uint8_t* src, uint32_t n, DstPtr dst, uint16_t* table, uint8_t* stage, bool handover = false, unsigned long long* dbg = nullptr)
  {
#define TB200_EPH(i) do { if (dbg) { const long long t__ = clock64(); acc_ph[i] += (uint32_t)(t__ - t_ph); t_ph = t__; } } while (0)
  long long t_ph = dbg ? clock64() : 0;
  uint32_t acc_ph[5] = {0, 0, 0, 0, 0};
  const unsigned lane = lane_id();
  const unsigned lt = lanemask_lt();
  uint32_t op = 0, anchor = 0;
  bool give_up = false;
  if (n >= LZ4_MFLIMIT + 1)
    {
    for (uint32_t i = lane; i < (1u << HLOG); i += 32) table[i] = 0;
    __syncwarp();
    const uint32_t mflimit = n - LZ4_MFLIMIT;        // last position where a match may start
    const uint32_t matchlimit = n - LZ4_LASTLITERALS;
    uint32_t p = 0, attempts = 0, nseq = 0;
    bool quarter_seen = false;
    // Dense mode: data that yields a match every few bytes WITHOUT getting smaller for it (noisy
    // planes: colours, quantised heights - four equal bytes turn up by chance all the time) pays
    // for every sequence twice, here and in the decoder.  Once the sequences of the block average
    // less than TB200_LZ4_DENSE_SEQ bytes and the output so far is above 90 % of the input, a
    // match has to be 8 bytes long to be taken.  (Blocks that do compress with short matches -
    // index planes of real meshes - never get here.)
    bool dense = false;
    // Deterministic insert of the lanes flagged `ins` (position q, bucket h): every lane stores, then
    // the losers of a bucket (they read back a lower position) store again - the highest position wins.
    auto insert = [&](bool ins, uint32_t h, uint32_t q)
      {
      __syncwarp();
      if (ins) table[h] = (uint16_t)q;
      __syncwarp();
      for (;;)
        {
        const bool lost = ins && table[h] < (uint16_t)q;
        if (!__any_sync(FULL, lost)) break;
        if (lost) table[h] = (uint16_t)q;
        __syncwarp();
        }
      };
    while (p <= mflimit)
      {
      if (!quarter_seen && p >= (n >> 2) && p >= 1024u)
        { // Nothing gained on the first quarter of the block: noise (the low planes of colours, heights,
          // jittered indices).  The block will be stored; searching the rest of it costs this warp
          // ~2000 cycles per 32 bytes for nothing.
        quarter_seen = true;
        if (lz4_not_worth(op + lz4_literal_run_bytes(p - anchor), p)) { give_up = true; break; }
        }
      const uint32_t stride = 1u + (attempts >> 6);
      uint32_t mq, mc;
      if (stride >= 2u && (attempts & 32u) == 0)
        { // Accelerated scan (nothing matched for 64+ positions): two positions per lane and step - the
          // 64 positions of two consecutive steps with this stride.  The two look-ups are independent,
          // which is what a lone warp is short of.  (The second half does not see the first half's
          // inserts; at this distance that costs nothing measurable.)
        const uint32_t qa = p + lane * stride, qb = qa + 32u * stride;
        const bool va = qa <= mflimit, vb = qb <= mflimit;
        const uint32_t sa = va ? smem_read32(src, qa) : 0u, sb = vb ? smem_read32(src, qb) : 0u;
        const uint32_t ha = (sa * 2654435761u) >> (32 - HLOG), hb = (sb * 2654435761u) >> (32 - HLOG);
        const uint32_t ca = table[ha], cb = table[hb];
        const uint32_t ra = smem_read32(src, ca), rb = smem_read32(src, cb);     // table entries are positions inside the block
        const bool oka = va && ca < qa && ra == sa, okb = vb && cb < qb && rb == sb;
        const unsigned maska = __ballot_sync(FULL, oka), maskb = __ballot_sync(FULL, okb);
        const int fa = maska ? __ffs((int)maska) - 1 : 32, fb = maska ? -1 : (maskb ? __ffs((int)maskb) - 1 : 32);
        // insert the positions up to the chosen match; the highest position of a bucket wins
        const bool ia = va && (int)lane <= fa, ib = vb && (int)lane <= fb;
        __syncwarp();
        if (ia) table[ha] = (uint16_t)qa;
        __syncwarp();
        if (ib) table[hb] = (uint16_t)qb;
        __syncwarp();
        for (;;)
          {
          const bool la = ia && table[ha] < (uint16_t)qa, lb = ib && table[hb] < (uint16_t)qb;
          if (!__any_sync(FULL, la || lb)) break;
          if (la) table[ha] = (uint16_t)qa;
          __syncwarp();
          if (lb && table[hb] < (uint16_t)qb) table[hb] = (uint16_t)qb;
          __syncwarp();
          }
        if ((maska | maskb) == 0) { p += 64u * stride; attempts += 64; TB200_EPH(0); continue; }
        const int f = maska ? fa : fb;
        mq = (maska ? p : p + 32u * stride) + (uint32_t)f * stride;
        mc = __shfl_sync(FULL, maska ? ca : cb, f);
        }
      else if (stride >= 2u)
        { // accelerated, one position per lane (the odd steps of the schedule above)
        const uint32_t q = p + lane * stride;
        const bool valid = q <= mflimit;
        const uint32_t seq = valid ? smem_read32(src, q) : 0u;
        const uint32_t h = (seq * 2654435761u) >> (32 - HLOG);
        const uint32_t cand = table[h];
        const bool ok = valid && cand < q && smem_read32(src, cand) == seq;
        const unsigned mask = __ballot_sync(FULL, ok);
        const int f = mask ? __ffs((int)mask) - 1 : 31;
        insert(valid && (int)lane <= f, h, q);
        if (mask == 0) { p += 32u * stride; attempts += 32; TB200_EPH(0); continue; }
        mq = p + (uint32_t)f * stride;
        mc = __shfl_sync(FULL, cand, f);
        }
      else
        { // ---- dense scan: positions p .. p+31 ----
        const uint32_t q = p + lane;
        const bool valid = q <= mflimit;
        const uint32_t seq = valid ? smem_read32(src, q) : 0u;
        const uint32_t h = (seq * 2654435761u) >> (32 - HLOG);
        const uint32_t c1 = table[h];                                              // candidate 1: the table's
        bool run1 = valid && c1 < q && smem_read32(src, c1) == seq;
        // candidate 2: the nearest earlier lane of the window with the same four bytes (the table is
        // read before this window is inserted, so repeats inside the window are invisible to it)
        const unsigned twins = __match_any_sync(FULL, valid ? seq : (0x5a000000u ^ lane)) & lt & __ballot_sync(FULL, valid);
        const uint32_t c2 = p + (31u - (uint32_t)__clz((int)(twins | 1u)));
        bool run2 = valid && twins != 0;
        if (dense)
          { // noisy data: a candidate has to match 8 bytes to be looked at at all
          const uint32_t nxt = smem_read32(src, q + 4u);
          run1 = run1 && smem_read32(src, c1 + 4u) == nxt;
          run2 = run2 && smem_read32(src, c2 + 4u) == nxt;
          }
        const unsigned anyok = __ballot_sync(FULL, run1 || run2);
        if (anyok == 0)
          {
          insert(valid, h, q);
          p += 32u; attempts += 32; TB200_EPH(0);
          continue;
          }
        // Long-match data (runs, periodic index planes) must not pay for the window machinery: the
        // first candidate of the window is measured by the whole warp (128 bytes in one step); if it
        // reaches the cap it takes the warp-wide path at once.
          {
          const int f0 = __ffs((int)anyok) - 1;
          const uint32_t c0 = __shfl_sync(FULL, run1 ? c1 : c2, f0);
          const uint32_t q0 = p + (uint32_t)f0;
          const uint32_t x = smem_read32(src, q0 + LZ4_MINMATCH + 4u * lane) ^ smem_read32(src, c0 + LZ4_MINMATCH + 4u * lane);
          const unsigned ne = __ballot_sync(FULL, x != 0);
          if (ne == 0 || 4u * ((uint32_t)__ffs((int)ne) - 1u) + LZ4_MINMATCH >= LZ4_WIN_CAP)
            {
            insert(valid && (int)lane <= f0, h, q);
            mq = q0; mc = c0;
            goto long_match;
            }
          }
        // match lengths up to LZ4_WIN_CAP, four bytes per step, both candidates side by side
        uint32_t l1 = run1 ? LZ4_MINMATCH : 0u, l2 = run2 ? LZ4_MINMATCH : 0u;
#pragma unroll 1
        for (uint32_t i = LZ4_MINMATCH; i < LZ4_WIN_CAP; i += 4)
          {
          const unsigned going = __ballot_sync(FULL, run1 || run2);
          if (going == 0) break;
          // every match of the window is still growing after 12 bytes: long-match data (runs, periodic
          // index planes) - no point in measuring further, the first one takes the warp-wide path
          if (i == 12u && going == anyok) break;
          const uint32_t a = smem_read32(src, q + i);
          if (run1)
            {
            const uint32_t x = a ^ smem_read32(src, c1 + i);
            if (x) { l1 += ((uint32_t)__ffs((int)x) - 1u) >> 3; run1 = false; } else l1 += 4u;
            }
          if (run2)
            {
            const uint32_t x = a ^ smem_read32(src, c2 + i);
            if (x) { l2 += ((uint32_t)__ffs((int)x) - 1u) >> 3; run2 = false; } else l2 += 4u;
            }
          }
        // the better candidate: one that is still growing beats one that ended (the table's if both
        // are: a far candidate that long is a real repeat, the near one a run); else the longer, the
        // nearer on a tie
        const bool take1 = run1 || (!run2 && l1 > l2);
        uint32_t cand = take1 ? c1 : c2;
        uint32_t L = take1 ? l1 : l2;
        bool capped = take1 ? run1 : run2;
        const uint32_t maxlen = matchlimit - (valid ? q : mflimit);               // >= 7
        if (L >= maxlen) { L = maxlen; capped = false; }                           // ends at the last position a match may cover
        const unsigned mask = __ballot_sync(FULL, L >= (dense ? 8u : LZ4_MINMATCH)), capmask = __ballot_sync(FULL, capped);
        if (mask == 0)
          { // (dense mode: nothing of 8+ bytes in this window)
          insert(valid, h, q);
          p += 32u; attempts += 32; TB200_EPH(0);
          continue;
          }
        // walk over the window: the matches of a greedy left-to-right parse
        unsigned sel = 0;
        int longf = -1;
          {
          uint32_t cur = 0;
          while (cur < 32u)
            {
            const unsigned m = mask & (0xffffffffu << cur);
            if (m == 0) break;
            const uint32_t f = (uint32_t)__ffs((int)m) - 1u;
            const bool capf = (capmask >> f) & 1u;
            const uint32_t Lf = __shfl_sync(FULL, L, f);
            if (!capf && f < 31u && ((mask >> (f + 1u)) & 1u))
              { // look-ahead: a longer match starts one byte later - position f becomes a literal
              const uint32_t Ln = __shfl_sync(FULL, L, f + 1u);
              if (((capmask >> (f + 1u)) & 1u) || Ln > Lf) { cur = f + 1u; continue; }
              }
            if (capf) { if (sel == 0) longf = (int)f; break; }
            sel |= 1u << f;
            cur = f + Lf;
            }
          }
        if (sel != 0)
          { // a first sequence behind a long literal run is emitted by the whole warp
          const uint32_t f0 = (uint32_t)__ffs((int)sel) - 1u;
          if (p + f0 - anchor >= 32u) { longf = (int)f0; sel = 0; }
          }
        attempts = 0;
        if (sel != 0)
          { // ---- all sequences of the window at once ----
          const bool chosen = (sel >> lane) & 1u;
          const uint32_t e_rel = lane + L;                                         // end of this lane's match, relative to p
          const unsigned below = sel & lt;
          const int jprev = below ? 31 - __clz((int)below) : 0;
          int prev_end = __shfl_sync(FULL, (int)e_rel, jprev);
          if (!below) prev_end = (int)anchor - (int)p;                             // <= 0: literals pending from before the window
          uint32_t lit = chosen ? (uint32_t)((int)lane - prev_end) : 0u;
          uint32_t mqi = q, mci = cand, Li = L;
          if (chosen)                                                              // backward extension over this sequence's own literals (lz4.c:947-950)
            while (lit > 0u && mci > 0u && src[mqi - 1u] == src[mci - 1u]) { --mqi; --mci; ++Li; --lit; }
          const uint32_t mcode = Li - LZ4_MINMATCH;
          uint32_t size = 0;
          if (chosen) size = 1u + lit + (lit >= 15u ? (lit - 15u) / 255u + 1u : 0u) + 2u + (mcode >= 15u ? (mcode - 15u) / 255u + 1u : 0u);
          uint32_t incl = size;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1)
            {
            const uint32_t up = __shfl_up_sync(FULL, incl, o);
            if (lane >= (unsigned)o) incl += up;
            }
          // the sequences that fit into the staging buffer (a prefix of the chosen ones; the first always does)
          const unsigned kept = sel & __ballot_sync(FULL, incl <= LZ4_ENC_STAGE);
          const int lastk = 31 - __clz((int)kept);
          const uint32_t total = __shfl_sync(FULL, incl, lastk);
          const uint32_t end_rel = __shfl_sync(FULL, e_rel, lastk);
          if ((kept >> lane) & 1u)
            {
            uint8_t* o = stage + (incl - size);
            uint32_t k = 0;
            o[k++] = (uint8_t)(((lit >= 15u ? 15u : lit) << 4) | (mcode >= 15u ? 15u : mcode));
            if (lit >= 15u) { uint32_t r = lit - 15u; while (r >= 255u) { o[k++] = 255; r -= 255u; } o[k++] = (uint8_t)r; }
            for (uint32_t j = 0; j < lit; ++j) o[k + j] = src[mqi - lit + j];
            k += lit;
            const uint32_t offv = mqi - mci;
            o[k++] = (uint8_t)offv; o[k++] = (uint8_t)(offv >> 8);
            if (mcode >= 15u) { uint32_t r = mcode - 15u; while (r >= 255u) { o[k++] = 255; r -= 255u; } o[k++] = (uint8_t)r; }
            }
          __syncwarp();
          for (uint32_t i = lane; i < total; i += 32) dst[op + i] = stage[i];
          op += total;
          // next scan position: behind the last sequence, or behind the window if nothing else in it matched
          const unsigned rest = end_rel < 32u ? (mask & (0xffffffffu << end_rel)) : 0u;
          const uint32_t pnew = (end_rel < 32u && rest == 0) ? p + 32u : p + end_rel;
          // inserts: scanned positions that are not inside a match
          const unsigned atbelow = kept & (lt | (1u << lane));
          const int js = atbelow ? 31 - __clz((int)atbelow) : 0;
          const uint32_t endjs = __shfl_sync(FULL, e_rel, js);
          const bool inside = atbelow != 0 && lane > (unsigned)js && lane < endjs;
          insert(valid && q < pnew && !inside, h, q);
          // like lz4.c:1118, remember one position inside the tail of a match that left the window
          if (end_rel > 32u && lane == 0 && pnew - 2u <= mflimit) table[(smem_read32(src, pnew - 2u) * 2654435761u) >> (32 - HLOG)] = (uint16_t)(pnew - 2u);
          __syncwarp();
          anchor = p + end_rel;
          p = pnew;
          nseq += (uint32_t)__popc(kept);
          if (handover && anchor >= LZ4_HANDOVER_MIN && nseq * LZ4_HANDOVER_SEQ > anchor) return 0xffffffffu;
          dense = TB200_LZ4_DENSE_SEQ != 0u && anchor >= (handover ? LZ4_HANDOVER_MIN : TB200_LZ4_DENSE_MIN) && nseq * TB200_LZ4_DENSE_SEQ > anchor && 10u * op > TB200_LZ4_DENSE_OUT * anchor;
          TB200_EPH(1);
          continue;
          }
        // a long first match
        insert(valid && (int)lane <= longf, h, q);
        mq = p + (uint32_t)longf;
        mc = __shfl_sync(FULL, cand, longf);
        }
long_match:
      attempts = 0;
      TB200_EPH(1);
      // backward extension over bytes not yet emitted (lz4.c:947-950 does the same serially)
        {
        const uint32_t room = min(min(mq - anchor, mc), 32u);
        const bool eqb = lane < room && src[mq - 1 - lane] == src[mc - 1 - lane];
        const unsigned neb = ~__ballot_sync(FULL, eqb);
        const uint32_t back = neb ? (uint32_t)__ffs((int)neb) - 1u : 32u;
        mq -= back; mc -= back;
        }
      // forward extension: a first step of 4 bytes per lane (most matches end inside 128 bytes),
      // then 16 bytes per lane = 512 bytes per step; bytes past the block end are zero padding and
      // the result is clamped to the last position a match may cover
      const uint32_t maxlen = matchlimit - mq;
      uint32_t len = LZ4_MINMATCH;
        {
        const uint32_t x = smem_read32(src, mq + len + 4 * lane) ^ smem_read32(src, mc + len + 4 * lane);
        const unsigned ne = __ballot_sync(FULL, x != 0);
        if (ne)
          {
          const int fl = __ffs((int)ne) - 1;
          const uint32_t xf = __shfl_sync(FULL, x, fl);
          len += 4u * (uint32_t)fl + (((uint32_t)__ffs((int)xf) - 1u) >> 3);
          }
        else
          {
          len += 128;
          len -= (mc + len) & 15u;      // back up (over bytes known to be equal) until the candidate side is 16-byte aligned: its vectors are single reads
          while (len < maxlen)
            {
            const uint4 va = smem_read128u(src, mq + len + 16 * lane), vb = *reinterpret_cast<const uint4*>(src + mc + len + 16 * lane);
            const uint32_t x0 = va.x ^ vb.x, x1 = va.y ^ vb.y, x2 = va.z ^ vb.z, x3 = va.w ^ vb.w;
            const unsigned nw = __ballot_sync(FULL, (x0 | x1 | x2 | x3) != 0);
            if (nw == 0) { len += 512; continue; }
            const int fl = __ffs((int)nw) - 1;
            const uint32_t xw = x0 ? x0 : x1 ? x1 : x2 ? x2 : x3;
            const uint32_t wi = x0 ? 0u : x1 ? 1u : x2 ? 2u : 3u;
            const uint32_t mine = 4u * wi + (((uint32_t)__ffs((int)xw) - 1u) >> 3);
            len += 16u * (uint32_t)fl + __shfl_sync(FULL, mine, fl);
            break;
            }
          }
        }
      if (len > maxlen) len = maxlen;
      TB200_EPH(2);
      op = lz4_emit(dst, op, src, anchor, mq - anchor, mq - mc, len);
      p = anchor = mq + len;
      ++nseq;
      if (handover && p >= LZ4_HANDOVER_MIN && nseq * LZ4_HANDOVER_SEQ > p) return 0xffffffffu;
      dense = TB200_LZ4_DENSE_SEQ != 0u && p >= (handover ? LZ4_HANDOVER_MIN : TB200_LZ4_DENSE_MIN) && nseq * TB200_LZ4_DENSE_SEQ > p && 10u * op > TB200_LZ4_DENSE_OUT * p;
      // like lz4.c:1118, remember one position inside the match tail
      if (lane == 0 && p - 2 <= mflimit) table[(smem_read32(src, p - 2) * 2654435761u) >> (32 - HLOG)] = (uint16_t)(p - 2);
      __syncwarp();
      TB200_EPH(3);
      }
    }
  if (give_up || lz4_not_worth(op + lz4_literal_run_bytes(n - anchor), n)) { op = 0; anchor = 0; }     // stored: the block becomes one literal run
  op = lz4_emit(dst, op, src, anchor, n - anchor, 0, 0);
  TB200_EPH(4);
  if (dbg && lane == 0) for (int i = 0; i < 5; ++i) atomicAdd(dbg + i, (unsigned long long)acc_ph[i]);
  return op;
#undef TB200_EPH
  }

// copy `n` bytes inside the output buffer from distance `dist` >= n behind (non-overlapping)
template <bool DST_GLOBAL>
__device__ __forceinline__ void lz4_copy_back(uint8_t* dst, uint32_t to, uint32_t dist, uint32_t n)
  {
  const unsigned lane = lane_id();
  const uint8_t* ms = dst + to - dist;
  if (DST_GLOBAL)
    {
    for (uint32_t i = lane; i < n; i += 32) dst[to + i] = __ldcg(ms + i);
    return;
    }
  if (n <= 64)
    {
    for (uint32_t i = lane; i < n; i += 32) dst[to + i] = ms[i];
    return;
    }
  const uint32_t head = (4u - (to & 3u)) & 3u;         // shared-memory dst buffers are 16-byte aligned
  if (lane < head) dst[to + lane] = ms[lane];
  const uint32_t nw = (n - head) >> 2;
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + to + head);
  const uint32_t sfrom = to - dist + head;
  for (uint32_t i = lane; i < nw; i += 32) dw[i] = smem_read32(dst, sfrom + 4 * i);
  const uint32_t done = head + (nw << 2);
  if (done + lane < n) dst[to + done + lane] = ms[done + lane];
  }

// Warp-cooperative LZ4 block decoder.  src: compressed bytes (global).  dst: output, 4-byte
// aligned; with DST_GLOBAL the match source is re-read from global memory behind a warp fence
// (reference-format whole-plane blocks), otherwise dst is shared memory.  Returns the number of
// bytes produced, or 0xffffffff on a malformed block (out-of-range offset / overrun).
template <bool DST_GLOBAL>
__device__ __forceinline__ uint32_t lz4_decompress_warp(const uint8_t* __restrict__ src, uint32_t src_len,
                                                        uint8_t* dst, uint32_t dst_cap)
  {
  const unsigned lane = lane_id();
  uint32_t ip = 0, op = 0;
  if (src_len == 0) return 0xffffffffu;
  const uint8_t* src_end = src + src_len;
  for (;;)
    {
    if (ip >= src_len) return 0xffffffffu;
    const uint32_t token = src[ip++];
    uint32_t lit = token >> 4;
    if (lit == 15)
      {
      uint32_t b;
      do { if (ip >= src_len) return 0xffffffffu; b = src[ip++]; lit += b; } while (b == 255);
      }
    if (ip + lit > src_len || op + lit > dst_cap) return 0xffffffffu;
    if (lit <= 64)
      {
      for (uint32_t i = lane; i < lit; i += 32) dst[op + i] = src[ip + i];
      }
    else
      { // long literal run: 16-byte stores to the aligned destination, source words funnel-shifted;
        // four vectors per lane are loaded before any is stored so the global latency overlaps
      const uint32_t head = (16u - ((uint32_t)reinterpret_cast<uintptr_t>(dst + op) & 15u)) & 15u;
      if (lane < head) dst[op + lane] = src[ip + lane];
      const uint32_t nv = (lit - head) >> 4;
      const uint8_t* sp = src + ip + head;
      const uint32_t* sa = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(sp) & ~(uintptr_t)3);
      const unsigned sh = ((unsigned)reinterpret_cast<uintptr_t>(sp) & 3u) * 8u;
      const uint32_t* s_last = reinterpret_cast<const uint32_t*>((reinterpret_cast<uintptr_t>(src_end) - 1) & ~(uintptr_t)3);   // last word holding a source byte
      uint4* dv = reinterpret_cast<uint4*>(dst + op + head);
      constexpr int UN = 4;
      for (uint32_t i0 = 0; i0 < nv; i0 += 32 * UN)
        {
        uint32_t w[UN][5];
#pragma unroll
        for (int u = 0; u < UN; ++u)
          {
          const uint32_t i = i0 + lane + 32 * u;
          const uint32_t* s = sa + 4 * (i < nv ? i : 0);
#pragma unroll
          for (int j = 0; j < 5; ++j) { const uint32_t* q = s + j; w[u][j] = *(q <= s_last ? q : s_last); }
          }
#pragma unroll
        for (int u = 0; u < UN; ++u)
          {
          const uint32_t i = i0 + lane + 32 * u;
          if (i < nv)
            dv[i] = make_uint4(__funnelshift_r(w[u][0], w[u][1], sh), __funnelshift_r(w[u][1], w[u][2], sh),
                               __funnelshift_r(w[u][2], w[u][3], sh), __funnelshift_r(w[u][3], w[u][4], sh));
          }
        }
      const uint32_t done = head + (nv << 4);
      if (done + lane < lit) dst[op + done + lane] = src[ip + done + lane];
      }
    ip += lit; op += lit;
    if (ip >= src_len) break;                       // last sequence has no match part
    if (ip + 2 > src_len) return 0xffffffffu;
    const uint32_t offset = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8);
    ip += 2;
    uint32_t mlen = token & 15u;
    if (mlen == 15)
      {
      uint32_t b;
      do { if (ip >= src_len) return 0xffffffffu; b = src[ip++]; mlen += b; } while (b == 255);
      }
    mlen += LZ4_MINMATCH;
    if (offset == 0 || offset > op || op + mlen > dst_cap) return 0xffffffffu;
    if (DST_GLOBAL) __threadfence_block();
    __syncwarp();                                   // literals of this sequence are visible
    // Overlapping copy.  The match is periodic with period `offset`, so bytes can be taken from
    // any multiple of `offset` behind.  A short period is first expanded to 128 bytes in one
    // step (lane l writes bytes 4l..4l+3 of the pattern); after that every pass copies as much
    // as is already written - the distance doubles, every pass is a plain non-overlapping copy.
    uint32_t copied = 0, dist = offset;
    if (!DST_GLOBAL && offset < 32u && mlen > offset)
      {
      const uint32_t n0 = min(mlen, 128u);
      const uint32_t inv = (65536u + offset - 1u) / offset;
      const uint8_t* ms = dst + op - offset;
      uint32_t r = 4u * lane;
      r -= offset * ((r * inv) >> 16);                        // (4*lane) mod offset
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j)
        {
        if (4u * lane + j < n0) dst[op + 4u * lane + j] = ms[r];
        r = (r + 1u == offset) ? 0u : r + 1u;
        }
      copied = n0;
      dist = ((offset + copied) / offset) * offset;           // largest multiple of the period now available
      __syncwarp();
      }
    while (copied < mlen)
      {
      const uint32_t chunk = min(dist, mlen - copied);
      lz4_copy_back<DST_GLOBAL>(dst, op + copied, dist, chunk);
      copied += chunk;
      if (dist < 2048u && dist < mlen) dist <<= 1;            // bytes [op-offset, op+copied) are periodic: 2*dist <= offset+copied
      if (DST_GLOBAL) __threadfence_block();
      __syncwarp();
      }
    op += mlen;
    }
  return op;
  }

// ---------------------------------------------------------------------------------------------
// K2 + K5: byte-plane split (trico_transpose_uint{16,32,64}_aos_to_soa,
// transpose_aos_to_soa.c:84-147) fused with per-plane-block LZ4 compression and the assembly.
//
// Two kernels.  lz4_encode_kernel: every WARP is independent (no barrier, no ordering): persistent
// warps pull ranges k of B elements from an atomic ticket and compress their planes p one after
// the other (chunk g = k*WB + p):
//   1. the warp reads the range's AoS elements with coalesced 16-byte loads and keeps byte p of
//      every element (the range comes from DRAM once; the passes for the other planes hit L2);
//      planes that the first pass found to be one repeated byte are not extracted at all
//   2. it compresses the plane block into the chunk's own scratch slot and records the size.
// lz4_assemble_kernel: block scan of the sizes + decoupled look-back over tiles of 64 chunks,
// then every block is copied to its final offset (only compressed bytes move twice).
// ---------------------------------------------------------------------------------------------
struct Lz4EncodeArgs
  {
  const void* in;          // device, n elements of WB bytes
  uint64_t n;
  uint32_t nranges;
  int log2B;
  uint8_t* sizes;          // u16 LE [nranges * WB]
  uint8_t* payload;
  uint8_t* total_field;
  uint64_t* total;
  uint8_t* scratch;        // one slot per chunk
  uint32_t slot;           // bytes per scratch slot (>= lz4_block_bound(B), multiple of 16)
  uint64_t* desc;          // look-back descriptors of the assemble kernel, zeroed
  uint32_t* ticket;        // chunk ticket, zeroed
  uint32_t* dense_list;    // chunk ids handed over to lz4_encode_dense_kernel (nullptr: never hand over)
  uint32_t* dense_count;   // zeroed
  unsigned long long* dbg; // phase-cycle counters by plane (experiments; nullptr in production)
  };

// whole v1 streams: the assembly's last tile also writes the fixed header fields (type, count, codec
// info, chunk size; the payload byte count is total_field) and makes *total the STREAM's byte count -
// a launch less per stream, which is what a batch of small streams is made of.  (Its own kernel
// parameter: growing Lz4EncodeArgs cost lz4_encode_kernel 4 % on C2.)
struct Lz4StreamHeader
  {
  uint8_t* hdr = nullptr;
  uint32_t hdr_count = 0;
  uint8_t hdr_type = 0, hdr_info = 0, hdr_log2 = 0;
  uint64_t hdr_fixed_plus_table = 0;
  uint32_t stages = 0;             // assembly: shared-memory stages of `slot` bytes for the bulk-copy ring (0: register path)
  };

constexpr int LZ4_HLOG = 10;     // default: 1024 u16 entries = 2 KiB per warp (12 resident warps per SM with 16 KiB blocks)
template <int WB> struct Lz4Cta { static constexpr int WARPS = WB > 4 ? WB : 4; };

// keeps byte `p` of every WB-byte element of one 16-byte vector; returns them packed LSB first
template <int WB>
__device__ __forceinline__ uint32_t plane_bytes(const uint4 v, uint32_t p)
  {
  if (WB == 4)
    {
    const uint32_t sel = p | ((4u + p) << 4);                       // byte p of the first, byte p of the second operand
    const uint32_t lo = __byte_perm(v.x, v.y, sel), hi = __byte_perm(v.z, v.w, sel);
    return __byte_perm(lo, hi, 0x5410);
    }
  if (WB == 2)
    { // 8 elements -> this returns the first four; see plane_bytes_hi for the rest
    const uint32_t sel = p | ((2u + p) << 4) | ((4u + p) << 8) | ((6u + p) << 12);
    return __byte_perm(v.x, v.y, sel);
    }
  // WB == 8: two elements
  const uint32_t a = p < 4 ? v.x : v.y, b = p < 4 ? v.z : v.w;
  return ((a >> (8 * (p & 3))) & 0xffu) | (((b >> (8 * (p & 3))) & 0xffu) << 8);
  }

#ifndef TB200_LZ4_ENC_LOAD_UN
#define TB200_LZ4_ENC_LOAD_UN 16     // 16-byte loads in flight per lane while a range is split into planes
#endif
template <int WB, int HLOG>
__global__ void __launch_bounds__(Lz4Cta<WB>::WARPS * 32, Lz4Cta<WB>::WARPS == 4 ? 3 : 2)      // shared memory admits 12 (16) resident warps per SM
lz4_encode_kernel(const Lz4EncodeArgs a)
  {
  constexpr int WARPS = Lz4Cta<WB>::WARPS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t B = 1u << a.log2B;
  const uint32_t pstride = B + LZ4_SRC_PAD;                          // block buffer padded for read-ahead
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  uint8_t* buf = smem_raw + (size_t)warp * pstride;
  uint16_t* table = reinterpret_cast<uint16_t*>(smem_raw + (size_t)WARPS * pstride) + ((size_t)warp << HLOG);
  uint8_t* stage = smem_raw + (size_t)WARPS * pstride + ((size_t)WARPS << HLOG) * sizeof(uint16_t) + (size_t)warp * LZ4_ENC_STAGE;

  // A ticket is one HALF of the planes of a range of B elements - the even or the odd ones - which
  // the warp compresses one after the other.  (The two halves are taken by two warps at about the
  // same time, so the range crosses from DRAM once; a warp that keeps a range to itself for all
  // planes holds it for ~60 us, and 1776 such ranges do not fit in L2.)  The pass that extracts the
  // first plane also notes in which bits the elements of the range differ at all: a plane none of
  // whose bits ever changes (the upper planes of index data, mostly) is one repeated byte and gets
  // its encoding written without being extracted or searched.
  // Tickets are taken one ahead so the next range can be pulled towards L2 meanwhile.
  constexpr uint32_t HALVES = WB > 1 ? 2 : 1;
  const uint64_t ntickets = (uint64_t)a.nranges * HALVES;
  uint32_t t32 = 0;
  if (lane == 0) t32 = atomicAdd(a.ticket, 1u);
  uint64_t tk = __shfl_sync(FULL, t32, 0);
  while (tk < ntickets)
    {
    if (lane == 0) t32 = atomicAdd(a.ticket, 1u);
    const uint64_t tknext = __shfl_sync(FULL, t32, 0);
    const uint64_t k = tk / HALVES, knext = tknext / HALVES;
    const uint32_t half = (uint32_t)(tk % HALVES);
    const uint64_t lo = k << a.log2B;
    const uint32_t cnt = (uint32_t)((a.n - lo < B) ? (a.n - lo) : B);
    const uint8_t* gin = reinterpret_cast<const uint8_t*>(a.in) + lo * WB;
    const bool aligned = (reinterpret_cast<uintptr_t>(gin) & 15u) == 0;

    uint32_t diff_lo = 0xffffffffu, diff_hi = 0xffffffffu;            // bits that differ somewhere in the range (bytes = planes); all set = unknown
    uint32_t e0_lo = 0, e0_hi = 0;                                     // the range's first element
#pragma unroll 1
    for (uint32_t p = half; p < (uint32_t)WB; p += HALVES)
      {
      const uint64_t g = k * WB + p;
      if (p + HALVES >= (uint32_t)WB && knext != k && knext < a.nranges)
        { // the next range -> L2, a few microseconds before its first pass (earlier and it is evicted again)
        const uint8_t* nx = reinterpret_cast<const uint8_t*>(a.in) + (knext << a.log2B) * WB;
        const uint64_t lim = a.n * WB;
        for (uint32_t o = lane * 128u; o < B * WB; o += 32u * 128u)
          if ((knext << a.log2B) * WB + o < lim) asm volatile("prefetch.global.L2 [%0];" :: "l"(nx + o));
        }
      long long t_ph = a.dbg ? clock64() : 0;
      uint32_t nbytes;
      const uint32_t pdiff = ((p < 4 ? diff_lo : diff_hi) >> (8 * (p & 3))) & 0xffu;
      if (pdiff == 0)
        nbytes = lz4_emit_run(a.scratch + g * a.slot, cnt, ((p < 4 ? e0_lo : e0_hi) >> (8 * (p & 3))) & 0xffu);
      else
        {
        // 1. plane p of the range -> buf
        if (aligned)
          {
          constexpr int EPV = 16 / WB;                                   // elements per 16-byte vector
          constexpr int UN = TB200_LZ4_ENC_LOAD_UN;
          const uint32_t nvec = cnt / EPV;
          const uint4* g4 = reinterpret_cast<const uint4*>(gin);
          const uint32_t nfull = nvec / (32 * UN);                       // whole batches of UN vectors per lane
          const uint32_t sel2 = p | ((2u + p) << 4) | ((4u + p) << 8) | ((6u + p) << 12);
          // first pass over a complete range: collect the differing bits of all planes
          const bool survey = p == half && WB > 1 && cnt == B && nfull * 32 * UN == nvec;
          const bool SURVEY = survey;
            {
            uint32_t r_lo = 0, r_hi = 0, d_lo = 0, d_hi = 0;               // reference element (this lane's first), differences to it
            auto put = [&](const uint4 v, uint32_t i)
              { // byte p of every element of vector i -> plane buffer
              if (WB == 1) reinterpret_cast<uint4*>(buf)[i] = v;
              else if (WB == 4) reinterpret_cast<uint32_t*>(buf)[i] = plane_bytes<4>(v, p);
              else if (WB == 2) reinterpret_cast<uint2*>(buf)[i] = make_uint2(__byte_perm(v.x, v.y, sel2), __byte_perm(v.z, v.w, sel2));
              else reinterpret_cast<uint16_t*>(buf)[i] = (uint16_t)plane_bytes<8>(v, p);
              if (SURVEY)
                {
                if (WB == 8) { d_lo |= (v.x ^ r_lo) | (v.z ^ r_lo); d_hi |= (v.y ^ r_hi) | (v.w ^ r_hi); }
                else d_lo |= (v.x ^ r_lo) | (v.y ^ r_lo) | (v.z ^ r_lo) | (v.w ^ r_lo);
                }
              };
            // rolling window of UN vectors per lane: as soon as a vector has been split its register takes the
            // load of the same position in the next batch, so UN loads per lane stay in flight throughout
            uint4 r[UN];
            if (nfull)
              {
#pragma unroll
              for (int u = 0; u < UN; ++u) r[u] = __ldg(g4 + lane + 32 * u);
              if (SURVEY)
                { // 2-byte elements: compare whole words against the first element in both halves
                r_lo = WB == 2 ? __byte_perm(r[0].x, 0u, 0x1010) : r[0].x;
                r_hi = r[0].y;
                }
              }
            for (uint32_t bidx = 0; bidx < nfull; ++bidx)
              {
              const bool more = bidx + 1 < nfull;
#pragma unroll
              for (int u = 0; u < UN; ++u)
                {
                put(r[u], bidx * 32 * UN + lane + 32 * u);
                if (more) r[u] = __ldg(g4 + (bidx + 1) * 32 * UN + lane + 32 * u);
                }
              }
            for (uint32_t i = nfull * 32 * UN + lane; i < nvec; i += 32) put(__ldg(g4 + i), i);
            for (uint32_t i = nvec * EPV + lane; i < cnt; i += 32) buf[i] = gin[(size_t)i * WB + p];
            if (SURVEY)
              { // differences inside the lanes, plus between the lanes' reference elements
              e0_lo = __shfl_sync(FULL, r_lo, 0); e0_hi = __shfl_sync(FULL, r_hi, 0);
              d_lo |= r_lo ^ e0_lo; d_hi |= r_hi ^ e0_hi;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) { d_lo |= __shfl_xor_sync(FULL, d_lo, o); d_hi |= __shfl_xor_sync(FULL, d_hi, o); }
              if (WB == 2) d_lo |= d_lo >> 16;                             // both halves of a word hold an element
              diff_lo = d_lo; diff_hi = d_hi;
              }
            }
          }
        else
          for (uint32_t i = lane; i < cnt; i += 32) buf[i] = gin[(size_t)i * WB + p];
        // zero the read-ahead pad: the compressor compares up to LZ4_SRC_PAD bytes past cnt
        for (uint32_t i = lane; i < LZ4_SRC_PAD; i += 32) buf[cnt + i] = 0;
        __syncwarp();
        if (a.dbg && lane == 0) { const long long t2 = clock64(); atomicAdd(a.dbg + (p & 7), (unsigned long long)(t2 - t_ph)); t_ph = t2; }

        // 2. compress into this chunk's slot
        nbytes = lz4_compress_warp<HLOG>(buf, cnt, a.scratch + g * a.slot, table, stage, a.dense_list != nullptr, a.dbg ? a.dbg + 16 + 8 * (p & 7) : nullptr);
        if (nbytes == 0xffffffffu)
          { // many short sequences: the lane-parallel parser of the second pass takes this block
          if (lane == 0) a.dense_list[atomicAdd(a.dense_count, 1u)] = (uint32_t)g;
          nbytes = 0;
          }
        }
      if (a.dbg && lane == 0) { const long long t2 = clock64(); atomicAdd(a.dbg + 8 + (p & 7), (unsigned long long)(t2 - t_ph)); }
      if (lane == 0)
        {
        uint8_t* sz = a.sizes + 2 * g;
        sz[0] = (uint8_t)nbytes; sz[1] = (uint8_t)(nbytes >> 8);
        }
      __syncwarp();
      }
    tk = tknext;
    }
  }

// Assembly: chunk g's block sits at scratch + g*slot with its size in sizes[g]; tiles of
// LZ4_ASM_TILE chunks take their base from a look-back over tile totals (all known up front, so
// no waiting), then the CTA copies the tile's blocks to their final offsets.
constexpr int LZ4_ASM_THREADS = 256;
constexpr int LZ4_ASM_TILE = 128;           // 64 measured the same, 256 slower

#ifndef TB200_HOST_EMU
#ifndef TB200_LZ4_ASM_BIG
#define TB200_LZ4_ASM_BIG 4096
#endif
constexpr uint32_t LZ4_ASM_BIG = TB200_LZ4_ASM_BIG;   // blocks from this size on are copied by the whole CTA

// nbytes from src (16-byte aligned, one readable spare vector behind the block) to dst (any alignment) by NT
// threads, tid = 0..NT-1
// destination vector = 16 source bytes starting `4 ws + sh / 8` bytes into the aligned pair (va, vb)
__device__ __forceinline__ uint4 lz4_asm_shift(const uint4 va, const uint4 vb, uint32_t ws, unsigned sh)
  {
  const uint32_t w[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
  switch (ws)
    {
    case 0: return make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh));
    case 1: return make_uint4(__funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh), __funnelshift_r(w[4], w[5], sh));
    case 2: return make_uint4(__funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh), __funnelshift_r(w[4], w[5], sh), __funnelshift_r(w[5], w[6], sh));
    default: return make_uint4(__funnelshift_r(w[3], w[4], sh), __funnelshift_r(w[4], w[5], sh), __funnelshift_r(w[5], w[6], sh), __funnelshift_r(w[6], w[7], sh));
    }
  }

// the same copy out of a shared-memory stage (16-byte aligned, filled by a bulk copy), whole CTA
__device__ __forceinline__ void lz4_asm_copy_staged(uint8_t* dst, const uint8_t* stage, uint32_t nbytes, uint32_t tid, uint32_t nthreads)
  {
  uint32_t head = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
  if (head > nbytes) head = nbytes;
  const uint32_t nvec = (nbytes - head) >> 4;
  const uint32_t done = head + (nvec << 4);
  if (tid < head) dst[tid] = stage[tid];
  if (done + tid < nbytes) dst[done + tid] = stage[done + tid];
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
  const uint4* sv = reinterpret_cast<const uint4*>(stage);
  const unsigned sh = (head & 3u) * 8u;
  const uint32_t ws = head >> 2;
  if (head)
    for (uint32_t i = tid; i < nvec; i += nthreads) __stcs(dv + i, lz4_asm_shift(sv[i], sv[i + 1], ws, sh));
  else
    for (uint32_t i = tid; i < nvec; i += nthreads) __stcs(dv + i, sv[i]);
  }

template <int NT>
__device__ __forceinline__ void lz4_asm_copy(uint8_t* dst, const uint8_t* src, uint32_t nbytes, uint32_t tid)
  {
  uint32_t head = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
  if (head > nbytes) head = nbytes;
  const uint32_t nvec = (nbytes - head) >> 4;
  const uint32_t done = head + (nvec << 4);
  uint8_t hb = 0, tb = 0;
  if (tid < head) hb = __ldcs(src + tid);
  if (done + tid < nbytes) tb = __ldcs(src + done + tid);
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
  const uint4* sv = reinterpret_cast<const uint4*>(src);                      // vector i of the body = source bytes [head + 16 i, head + 16 i + 16)
  const unsigned sh = (head & 3u) * 8u;
  const uint32_t ws = head >> 2;                                              // 0..3, the same for the whole block
  constexpr int UN = 4;
  for (uint32_t i0 = tid; i0 < nvec; i0 += NT * UN)
    {
    uint4 va[UN], vb[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u)
      {
      const uint32_t i = i0 + (uint32_t)u * NT;
      if (i < nvec) { va[u] = __ldcs(sv + i); if (head) vb[u] = __ldcs(sv + i + 1); }
      }
    if (i0 == tid)
      { // (the first batch's loads are under way)
      if (tid < head) dst[tid] = hb;
      if (done + tid < nbytes) dst[done + tid] = tb;
      }
#pragma unroll
    for (int u = 0; u < UN; ++u)
      {
      const uint32_t i = i0 + (uint32_t)u * NT;
      if (i >= nvec) continue;
      dv[i] = head ? lz4_asm_shift(va[u], vb[u], ws, sh) : va[u];
      }
    }
  if (tid >= nvec)
    { // threads that had no vector of the first batch
    if (tid < head) dst[tid] = hb;
    if (done + tid < nbytes) dst[done + tid] = tb;
    }
  }

__global__ void __launch_bounds__(LZ4_ASM_THREADS)
lz4_assemble_kernel(const Lz4EncodeArgs a, uint64_t nchunks, const Lz4StreamHeader h)
  {
  __shared__ uint32_t sh_off[LZ4_ASM_TILE];
  __shared__ uint32_t sh_sz[LZ4_ASM_TILE];
  __shared__ uint32_t sh_wsum[LZ4_ASM_TILE / 32];
  __shared__ uint64_t sh_base;
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t tile = blockIdx.x;
  const uint64_t g0 = (uint64_t)tile * LZ4_ASM_TILE;
  uint32_t mysz = 0;
  if (threadIdx.x < LZ4_ASM_TILE && g0 + threadIdx.x < nchunks)
    {
    const uint8_t* sz = a.sizes + 2 * (g0 + threadIdx.x);
    mysz = (uint32_t)sz[0] | ((uint32_t)sz[1] << 8);
    }
  uint32_t incl = mysz;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    {
    const uint32_t up = __shfl_up_sync(FULL, incl, o);
    if (lane >= (unsigned)o) incl += up;
    }
  if (lane == 31 && warp < LZ4_ASM_TILE / 32) sh_wsum[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0, tsum = 0;
#pragma unroll
  for (int w = 0; w < LZ4_ASM_TILE / 32; ++w) { const uint32_t t = sh_wsum[w]; if (w < (int)warp) wbase += t; tsum += t; }
  if (threadIdx.x < LZ4_ASM_TILE)
    {
    sh_off[threadIdx.x] = wbase + incl - mysz;
    sh_sz[threadIdx.x] = mysz;
    }
  if (warp == 0)
    {
    const uint64_t excl = lookback_exclusive(a.desc, tile, tsum);
    if (lane == 0)
      {
      sh_base = excl;
      if (tile == gridDim.x - 1)
        {
        store_u64_bytes(a.total_field, excl + tsum);
        if (h.hdr)
          {
          h.hdr[0] = h.hdr_type;
          h.hdr[1] = (uint8_t)h.hdr_count; h.hdr[2] = (uint8_t)(h.hdr_count >> 8); h.hdr[3] = (uint8_t)(h.hdr_count >> 16); h.hdr[4] = (uint8_t)(h.hdr_count >> 24);
          h.hdr[5] = h.hdr_info; h.hdr[6] = h.hdr_log2;
          *a.total = h.hdr_fixed_plus_table + excl + tsum;
          }
        else *a.total = excl + tsum;
        }
      }
    }
  __syncthreads();
  // Large blocks (an incompressible plane is 200 times larger than a constant one) are copied by the
  // whole CTA, one after the other; every smaller block takes one warp, eight blocks in flight per CTA.
  // Either way (lz4_asm_copy): 16-byte stores to the aligned body of the destination, each built from
  // the two aligned source vectors around it (the shift is the same for the whole block), and EVERY
  // load of a batch - head bytes, tail bytes, vectors - is issued before the first store: a block of a
  // few hundred bytes costs one memory latency, not one per 32 bytes.
  const uint32_t nloc = (uint32_t)((nchunks - g0 < (uint64_t)LZ4_ASM_TILE) ? (nchunks - g0) : (uint64_t)LZ4_ASM_TILE);
  uint8_t* D = a.payload + sh_base;
  const uint8_t* sbase = a.scratch + g0 * a.slot;
  // With a stage ring (h.stages > 0) the large blocks arrive by bulk copy (TMA, one instruction of one
  // thread per block), `stages` blocks ahead of the CTA: the bytes in flight per SM are bounded by
  // shared memory (3 CTAs x 4 x 16.5 KB) instead of by registers (4 CTAs x 256 threads x 8 vectors), and
  // the small blocks' latencies pass while the first stages fill.
  // (16, like every other kernel's dynamic shared memory: all `extern __shared__` arrays of a translation unit
  // are ONE symbol, and a larger alignment here moves the dynamic base of every other kernel - measured: it
  // pads lz4_decode_multi_kernel's static shared memory from 1360 to 1408 bytes and that kernel then faults)
  extern __shared__ __align__(16) uint8_t asm_stage[];
  __shared__ uint8_t sh_big[LZ4_ASM_TILE];
  __shared__ uint32_t sh_nbig;
  __shared__ __align__(8) uint64_t sh_full[4];
  const uint32_t stages = h.stages;
  uint64_t pol_first = 0;
  if (stages)
    {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    if (warp == 0)
      {
      uint32_t n = 0;
      for (uint32_t c0 = 0; c0 < nloc; c0 += 32)
        {
        const uint32_t c = c0 + lane;
        const bool big = c < nloc && sh_sz[c] >= LZ4_ASM_BIG;
        const unsigned m = __ballot_sync(FULL, big);
        if (big) sh_big[n + __popc(m & lanemask_lt())] = (uint8_t)c;
        n += __popc(m);
        }
      if (lane == 0)
        {
        sh_nbig = n;
        for (uint32_t st = 0; st < stages; ++st) mbar_init((uint32_t)__cvta_generic_to_shared(&sh_full[st]), 1);
        }
      }
    __syncthreads();
    }
  const uint32_t nbig = stages ? sh_nbig : 0;
  const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(asm_stage);
  auto issue = [&](uint32_t j)
    { // one thread: block j of the list -> its stage
    const uint32_t c = sh_big[j], st = j % stages;
    const uint32_t nb16 = (sh_sz[c] + 15u) & ~15u;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sh_full[st]);
    mbar_expect_tx(bar, nb16);
    bulk_g2s(stage_s + st * a.slot, sbase + (size_t)c * a.slot, nb16, bar, pol_first);
    };
  if (stages && threadIdx.x == 0)
    for (uint32_t j = 0; j < stages && j < nbig; ++j) issue(j);
  if (!stages)
    for (uint32_t c = 0; c < nloc; ++c)
      {
      const uint32_t nbytes = sh_sz[c];
      if (nbytes < LZ4_ASM_BIG) continue;
      lz4_asm_copy<LZ4_ASM_THREADS>(D + sh_off[c], sbase + (size_t)c * a.slot, nbytes, threadIdx.x);
      }
  for (uint32_t c = warp; c < nloc; c += LZ4_ASM_THREADS / 32)
    {
    const uint32_t nbytes = sh_sz[c];
    if (nbytes >= LZ4_ASM_BIG || nbytes == 0) continue;
    lz4_asm_copy<32>(D + sh_off[c], sbase + (size_t)c * a.slot, nbytes, lane);
    }
  for (uint32_t j = 0; j < nbig; ++j)
    {
    const uint32_t c = sh_big[j], st = j % stages;
    mbar_wait((uint32_t)__cvta_generic_to_shared(&sh_full[st]), (j / stages) & 1u);
    lz4_asm_copy_staged(D + sh_off[c], asm_stage + (size_t)st * a.slot, sh_sz[c], threadIdx.x, LZ4_ASM_THREADS);
    __syncthreads();                                        // the stage has been read by everybody
    if (threadIdx.x == 0 && j + stages < nbig) issue(j + stages);
    }
  }

#endif // TB200_HOST_EMU

// ---------------------------------------------------------------------------------------------
// In-place block decoder on shared memory (K6).  The compressed block sits at the END of the
// plane's buffer, the output grows from the start: with LZ4's in-place margin
// ((csize >> 8) + 32 bytes, lz4.h "LZ4_DECOMPRESS_INPLACE_MARGIN") the write position never
// passes the read position, so one buffer serves both and every token / length byte is a
// shared-memory read.  All copies are warp-wide: literals and far matches move 512 bytes per
// step (16 per lane); near matches (offset <= 32: runs, interleaved index patterns) are periodic
// and are generated from their first 64 bytes with independent 16-byte reads and stores.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t lz4_inplace_stride(uint32_t B)
  { // output + in-place margin of the largest block + 16 (vector-store overshoot) + 15 (alignment of
    // the staged block) + 32 readable bytes behind the block + slack
  return ((B + (lz4_block_bound(B) >> 8) + 32u + 16u + 15u + 32u + 17u) + 15u) & ~15u;
  }

// reads a length continuation (bytes of 255 terminated by a byte < 255, lz4.c:1629-1649) at ip;
// 32 bytes are inspected per step.  Returns the sum and advances ip past the terminator.
__device__ __forceinline__ uint32_t lz4_read_ext(const uint8_t* buf, uint32_t& ip, uint32_t iend)
  {
  const unsigned lane = lane_id();
  uint32_t add = 0;
  for (;;)
    {
    const uint32_t b = (ip + lane < iend) ? buf[ip + lane] : 0u;       // past the end: terminates, caller's bound checks fail
    const unsigned m = __ballot_sync(FULL, b != 255u);
    if (m == 0) { add += 255u * 32u; ip += 32; continue; }
    const int e = __ffs((int)m) - 1;
    add += 255u * (uint32_t)e + __shfl_sync(FULL, b, e);
    ip += (uint32_t)e + 1u;
    return add;
    }
  }

// Warp copy inside one shared-memory buffer, dst and src arbitrary.  FORWARD_OVERLAP: dst < src
// and the regions may overlap (literals of the in-place decoder) - every batch is read completely
// before it is written.  Otherwise the regions must not overlap at all (src + n <= dst).
template <bool FORWARD_OVERLAP>
__device__ __forceinline__ void lz4_smem_move(uint8_t* buf, uint32_t dst, uint32_t src, uint32_t n)
  {
  const unsigned lane = lane_id();
  if (n <= 32)
    {
    uint32_t t = 0;
    if (lane < n) t = buf[src + lane];
    if (FORWARD_OVERLAP) __syncwarp();
    if (lane < n) buf[dst + lane] = (uint8_t)t;
    return;
    }
  uint32_t head = (16u - (dst & 15u)) & 15u;
    {
    uint32_t t = 0;
    if (lane < head) t = buf[src + lane];
    if (FORWARD_OVERLAP) __syncwarp();
    if (lane < head) buf[dst + lane] = (uint8_t)t;
    }
  const uint32_t nv = (n - head) >> 4;
  const uint32_t s0 = src + head, d0 = dst + head;
  constexpr int UN = 4;
  for (uint32_t i0 = 0; i0 < nv; i0 += 32 * UN)
    {
    uint4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u)
      {
      const uint32_t i = i0 + lane + 32 * u;
      if (i < nv) v[u] = smem_read128u(buf, s0 + 16u * i);
      }
    if (FORWARD_OVERLAP) __syncwarp();
#pragma unroll
    for (int u = 0; u < UN; ++u)
      {
      const uint32_t i = i0 + lane + 32 * u;
      if (i < nv) *reinterpret_cast<uint4*>(buf + d0 + 16u * i) = v[u];
      }
    }
  const uint32_t done = head + (nv << 4);
    {
    uint32_t t = 0;
    if (done + lane < n) t = buf[src + done + lane];
    if (FORWARD_OVERLAP) __syncwarp();
    if (done + lane < n) buf[dst + done + lane] = (uint8_t)t;
    }
  }

// ceil(2^32 / o) for o = 2..32: t mod o = t - o * umulhi(t, inv) is exact for t < 2^16.  Offset 1
// would need 2^32; its entry is 0 and the result is masked to 0 (every index of a run is 0).
__constant__ uint32_t c_lz4_inv[33] = { 0x00000000u, 0x00000000u, 0x80000000u, 0x55555556u, 0x40000000u, 0x33333334u, 0x2aaaaaabu, 0x24924925u, 0x20000000u, 0x1c71c71du, 0x1999999au, 0x1745d175u, 0x15555556u, 0x13b13b14u, 0x12492493u, 0x11111112u, 0x10000000u, 0x0f0f0f10u, 0x0e38e38fu, 0x0d79435fu, 0x0ccccccdu, 0x0c30c30du, 0x0ba2e8bbu, 0x0b21642du, 0x0aaaaaabu, 0x0a3d70a4u, 0x09d89d8au, 0x097b425fu, 0x0924924au, 0x08d3dcb1u, 0x08888889u, 0x08421085u, 0x08000000u };

// S(offset) = 32 - 32 mod Q, Q = offset / gcd(offset, 16): see lz4_match_warp
__constant__ uint8_t c_lz4_stride[33] = { 32, 32, 32, 30, 32, 30, 30, 28, 32, 27, 30, 22, 30, 26, 28, 30, 32, 17, 27, 19, 30, 21, 22, 23, 30, 25, 26, 27, 28, 29, 30, 31, 32 };

// Rest of a short-period match (offset <= 32 < mlen) once its first 64 bytes are in place: they
// serve as a look-up table - the 16 bytes at any later position x are the 16 bytes at
// opm + (x - opm) mod offset - so the remainder is produced with independent unaligned 16-byte
// reads and aligned stores.  The aligned vectors repeat every Q = offset / gcd(offset, 16)
// vectors, so a lane that strides by S = the largest multiple of Q not above 32 stores the SAME
// vector every time: one table read per lane, whatever the match length.  Exact: nothing is written
// at or past opm + mlen.
template <typename Mod>
__device__ __forceinline__ void lz4_period_rest(uint8_t* buf, uint32_t opm, uint32_t mlen, uint32_t S, Mod modo)
  {
  const unsigned lane = lane_id();
  const uint32_t end = opm + mlen;
  const uint32_t xa = (opm + 64u) & ~15u;                                // aligned, inside the table: rewriting [xa, opm+64) is harmless
  const uint32_t nv = (end - xa) >> 4;
  const uint32_t t0 = xa - opm;
  if (lane < S)
    {
    const uint4 v = smem_read128(buf, opm + modo(t0 + 16u * lane));
    for (uint32_t i = lane; i < nv; i += S) *reinterpret_cast<uint4*>(buf + xa + 16u * i) = v;
    }
  const uint32_t done = xa + (nv << 4);
  if (done + lane < end) buf[done + lane] = buf[opm + modo(done - opm + lane)];
  }

// One match, produced by the whole warp.  opm = output position of the match.  Exact: nothing is
// written at or past opm + mlen.
__device__ __forceinline__ void lz4_match_warp(uint8_t* buf, uint32_t opm, uint32_t offset, uint32_t mlen)
  {
  const unsigned lane = lane_id();
  if (offset <= 32u && mlen > offset)
    { // Short period (runs, interleaved index patterns).  The first 64 output bytes are written
      // byte-wise (lane l: bytes l and l+32 of the pattern), the rest by lz4_period_rest.
    const uint8_t* ms = buf + opm - offset;
    const uint32_t inv = c_lz4_inv[offset];
    const uint32_t keep = offset == 1u ? 0u : 0xffffffffu;
    auto modo = [&](uint32_t t) { return (t - offset * __umulhi(t, inv)) & keep; };
    const uint32_t q0 = ms[modo(lane)], q1 = ms[modo(lane + 32u)];
    if (lane < mlen) buf[opm + lane] = (uint8_t)q0;
    if (lane + 32u < mlen) buf[opm + lane + 32u] = (uint8_t)q1;
    if (mlen > 64u)
      {
      __syncwarp();
      lz4_period_rest(buf, opm, mlen, c_lz4_stride[offset], modo);
      }
    }
  else if (offset >= mlen)
    lz4_smem_move<false>(buf, opm, opm - offset, mlen);
  else
    { // periodic with a long period: grow by copying everything available, distance doubling
    uint32_t copied = 0, dist = offset;
    while (copied < mlen)
      {
      const uint32_t chunk = min(dist, mlen - copied);
      lz4_smem_move<false>(buf, opm + copied, opm + copied - dist, chunk);
      copied += chunk;
      if (dist < 4096u) dist <<= 1;               // [opm-offset, opm+copied) is periodic: 2*dist <= offset + copied
      __syncwarp();
      }
    }
  }

// Up to 32 consecutive SHORT sequences at once (literal run <= LZ4_BATCH_MAXLIT, match <= LZ4_BATCH_MAXMATCH):
// the regime of real index planes, colour planes and attribute lists, where a sequence is a few
// bytes and the one-sequence-per-iteration loop below pays ~1000 cycles for each.
//   1. the token walk, speculatively: every lane computes the length of the sequence that WOULD start
//      at each of its four bytes of a 128-byte window (independent reads), then the chain through the
//      window is followed with one shuffle per sequence; lane k keeps sequence k.  (A serial walk -
//      one dependent shared-memory read and ~25 instructions per sequence - was 70 % of the batch.)
//   2. every lane fetches its own literals and offset, then writes the literals - all input is read
//      before anything is written: the output of a later sequence may cover the (consumed) input of
//      an earlier one, in-place margin or not
//   3. the matches, every lane its own, in rounds: a match is copied once everything below the
//      end of its source is final, i.e. lies below the first match that is still open.  Far matches
//      (the common case) all go in the first round.
// Returns the number of sequences done (0: the next one is not of this kind), -1 on a malformed one.
constexpr uint32_t LZ4_BATCH_MAXMATCH = 64;
constexpr uint32_t LZ4_BATCH_MAXLIT = 32;
// ib: compressed bytes, ob: output (the same buffer for the in-place decoder); `floor`: index of the
// first real output byte in ob (offsets must not reach below it).
__device__ __forceinline__ int lz4_decode_batch(const uint8_t* ib, uint8_t* ob, uint32_t& ip, uint32_t iend, uint32_t& op, uint32_t cap, uint32_t floor = 0)
  {
  const unsigned lane = lane_id();
  // 1a. speculation: every lane works out, for each of ITS four bytes of the 128-byte window behind
  //     ip, how long a batchable sequence starting at that byte would be (token, literal length
  //     byte, match length byte: the same tests as a serial walk makes) - 0 if none could start there.
  //     The reads are independent of each other; 32 readable bytes follow every block.
  const uint32_t base = ip + 4u * lane;
  uint32_t packed = 0;
  if (base + 8u <= iend + 32u)
    {
    const uint32_t w0 = smem_read32(ib, base), w1 = smem_read32(ib, base + 4u);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      {
      const uint32_t q = base + (uint32_t)j;
      const uint32_t t = (w0 >> (8 * j)) & 0xffu;
      const uint32_t x = j < 3 ? (w0 >> (8 * (j + 1))) & 0xffu : w1 & 0xffu;
      uint32_t lit = t >> 4;
      uint32_t adv = 3u;
      bool ok = q < iend;
      if (lit == 15u) { ok = ok && x <= LZ4_BATCH_MAXLIT - 15u; lit += x; ++adv; }      // else: a long literal run
      adv += lit;
      ok = ok && q + adv <= iend;                                 // else: the block's last sequence (no match part)
      if (ok && (t & 15u) == 15u) { ok = (uint32_t)ib[q + adv] <= LZ4_BATCH_MAXMATCH - 19u; ++adv; }   // else: a long match
      packed |= (ok ? adv : 0u) << (8 * j);
      }
    }
  // 1b. the chain through the window: one shuffle per sequence (the serial walk paid a dependent
  //     shared-memory read and ~25 instructions for each); lane k keeps sequence k
  uint32_t cur = 0, my_q = 0, my_adv = 0;
  int nseq = 0;
#pragma unroll 1
  for (int k = 0; k < 32 && cur < 128u; ++k)
    {
    const uint32_t adv = (__shfl_sync(FULL, packed, (int)(cur >> 2)) >> (8u * (cur & 3u))) & 0xffu;
    if (adv == 0u) break;
    if ((int)lane == k) { my_q = cur; my_adv = adv; }
    cur += adv; ++nseq;
    }
  if (nseq == 0) return 0;
  // 1c. every lane reads the lengths of its own sequence; output positions by a scan
  const uint32_t my_sp = ip + my_q;
  uint32_t my_lit = 0, my_ml = 0;
  if ((int)lane < nseq)
    {
    const uint32_t t = ib[my_sp];
    my_lit = t >> 4;
    if (my_lit == 15u) my_lit += ib[my_sp + 1u];
    my_ml = LZ4_MINMATCH + (t & 15u);
    if ((t & 15u) == 15u) my_ml += ib[my_sp + my_adv - 1u];
    }
  uint32_t incl = my_lit + my_ml;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
    {
    const uint32_t up = __shfl_up_sync(FULL, incl, d);
    if (lane >= (unsigned)d) incl += up;
    }
  const uint32_t my_op = op + incl - (my_lit + my_ml);
  // a sequence that does not fit ends the batch (a malformed block, or the end of a segment: the caller decides)
  const unsigned over = __ballot_sync(FULL, (int)lane < nseq && op + incl > cap);
  if (over) nseq = __ffs((int)over) - 1;
  if (nseq == 0) return 0;
  const uint32_t sp = ip + __shfl_sync(FULL, my_q + my_adv, nseq - 1);
  const uint32_t o = op + __shfl_sync(FULL, incl, nseq - 1);
  const bool act = (int)lane < nseq;
  uint64_t lw[4] = {0, 0, 0, 0};                            // this lane's literals (<= 32 bytes)
  uint32_t off = 0;
  if (act)
    {
    const uint32_t l0 = my_sp + (my_lit >= 15u ? 2u : 1u);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      {
      if (my_lit > 8u * u) lw[u] = (uint64_t)smem_read32(ib, l0 + 8u * u);
      if (my_lit > 8u * u + 4u) lw[u] |= (uint64_t)smem_read32(ib, l0 + 8u * u + 4u) << 32;
      }
    const uint32_t e = l0 + my_lit;
    off = (uint32_t)ib[e] | ((uint32_t)ib[e + 1u] << 8);
    }
  const uint32_t dst = my_op + my_lit;
  const bool fine = !act || (off != 0u && off <= dst - floor && dst + my_ml <= cap);
  if (!__all_sync(FULL, fine)) return -1;
  __syncwarp();
  if (act)
    {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      for (uint32_t j = 8u * u; j < my_lit && j < 8u * u + 8u; ++j)
        ob[my_op + j] = (uint8_t)(lw[u] >> (8u * (j - 8u * u)));
    }
  __syncwarp();
  // the matches of earlier sequences this lane's source overlaps: [s, e) = the part of the source
  // that this lane does not produce itself, against every earlier match [dst_j, dst_j + ml_j)
  const uint32_t s_lo = dst - off, s_hi = act ? min(s_lo + my_ml, dst) : 0u;
  // the earlier matches are sorted by position, so those that overlap [s_lo, s_hi) are a range of
  // lanes [jlo, jhi): jlo = the first match that ends behind s_lo, jhi = the first one that starts at
  // or behind s_hi - two binary searches over the lanes' registers
  const uint32_t m_end = dst + my_ml;
  uint32_t jlo = 0, jhi = 0;
    {
    uint32_t lo = 0, hi = (uint32_t)nseq, lo2 = 0, hi2 = (uint32_t)nseq;
#pragma unroll
    for (int step = 0; step < 6; ++step)
      {
      const uint32_t mid = (lo + hi) >> 1, mid2 = (lo2 + hi2) >> 1;
      const uint32_t be = __shfl_sync(FULL, m_end, mid & 31u), as = __shfl_sync(FULL, dst, mid2 & 31u);
      if (lo < hi) { if (be <= s_lo) lo = mid + 1u; else hi = mid; }
      if (lo2 < hi2) { if (as < s_hi) lo2 = mid2 + 1u; else hi2 = mid2; }
      }
    jlo = lo; jhi = lo2;
    }
  if (jhi > lane) jhi = lane;                                // only earlier sequences
  const unsigned waits = (act && jhi > jlo) ? (((jhi >= 32u ? 0u : (1u << jhi)) - 1u) & ~((1u << jlo) - 1u)) : 0u;
  bool open = act;
  for (;;)
    {
    const unsigned und = __ballot_sync(FULL, open);
    if (und == 0) break;
    if (open && (und & waits) == 0)
      {
      const uint8_t* ms = ob + dst - off;
      uint8_t* md = ob + dst;
      if (off >= 4u)
        { // four bytes per step: the loads of a step lie before its first store
        for (uint32_t j = 0; j < my_ml; j += 4u)
          {
          const uint32_t b0 = ms[j], b1 = ms[j + 1u], b2 = ms[j + 2u], b3 = ms[j + 3u];
          md[j] = (uint8_t)b0;
          if (j + 1u < my_ml) md[j + 1u] = (uint8_t)b1;
          if (j + 2u < my_ml) md[j + 2u] = (uint8_t)b2;
          if (j + 3u < my_ml) md[j + 3u] = (uint8_t)b3;
          }
        }
      else
        { // period 1..3: the pattern comes from a register
        const uint32_t pat = (uint32_t)ms[0] | ((uint32_t)ms[off > 1u ? 1 : 0] << 8) | ((uint32_t)ms[off > 2u ? 2 : 0] << 16);
        uint32_t idx = 0;
        for (uint32_t j = 0; j < my_ml; ++j)
          {
          md[j] = (uint8_t)(pat >> (8u * idx));
          idx = (idx + 1u == off) ? 0u : idx + 1u;
          }
        }
      open = false;
      }
    __syncwarp();
    }
  ip = sp; op = o;
  return nseq;
  }

// Decodes the block buf[ip, iend) into buf[0, cap).  Returns the bytes produced or 0xffffffff.
// One sequence per iteration, every step warp-wide.  A lone warp issues one dependent instruction
// every ~6 cycles, so the loop is written for instruction count: the token and the 31 bytes behind
// it arrive in one read, a short sequence takes its literals, offset and length continuation out
// of that register, and the match generator does no work it does not need.  (A variant that parsed
// 32 sequences ahead and ran literals / independent matches in lane groups executed more
// instructions on the critical warp than it saved and was slower on index planes.)
// The checks are the memory-safety ones (every access stays inside the buffer for any input); a
// malformed block is reported through the byte count it produces.
__device__ __forceinline__ uint32_t lz4_decode_inplace(uint8_t* buf, uint32_t ip, uint32_t iend, uint32_t cap)
  {
  const unsigned lane = lane_id();
  uint32_t op = 0;
  if (ip >= iend) return 0xffffffffu;
  // constants of the short-period generator, one offset per lane (lane l: offset l + 1); a shuffle
  // is quicker than an indexed constant-memory read inside the loop
  const uint32_t lane_inv = c_lz4_inv[lane + 1u], lane_stride = c_lz4_stride[lane + 1u];
  // the token and the 31 bytes behind it in one read (32 readable bytes follow every block); the
  // window of the NEXT sequence is requested as soon as its position is known, before the match of
  // the current one is produced
  uint32_t b = buf[ip + lane];
  bool batching = false;                        // the last sequence was a short one: try the batch decoder
  for (;;)
    {
    if (batching)
      {
      const int nb = lz4_decode_batch(buf, buf, ip, iend, op, cap);
      if (nb < 0) return 0xffffffffu;
      if (nb < 4) batching = false;
      if (nb > 0) { __syncwarp(); b = buf[ip + lane]; continue; }
      }
    const uint32_t token = __shfl_sync(FULL, b, 0);
    uint32_t lit = token >> 4, ml = token & 15u, offset;
    const bool short_lit = lit < 15u;
    if (short_lit)
      {
      const uint32_t e = lit + 1u;                                  // index of the offset's low byte inside b
      if (ip + e > iend || op + lit > cap) return 0xffffffffu;
      __syncwarp();                                                 // every lane holds its byte of b
      if (lane - 1u < lit) buf[op + lane - 1u] = (uint8_t)b;        // lanes 1..lit
      op += lit; ip += e;
      if (ip >= iend) break;                                        // last sequence has no match part
      offset = __shfl_sync(FULL, b, e) | (__shfl_sync(FULL, b, e + 1u) << 8);
      ip += 2;
      if (ml == 15u)
        { // continuation bytes start at index e+2 of b
        const unsigned m = __ballot_sync(FULL, b != 255u) >> (e + 2u);
        if (m)
          {
          const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
          ml += 255u * k + __shfl_sync(FULL, b, e + 2u + k);
          ip += k + 1u;
          }
        else
          {
          ml += 255u * (30u - e); ip += 30u - e;
          ml += lz4_read_ext(buf, ip, iend);
          }
        }
      }
    else
      {
      ip += 1;
      lit += lz4_read_ext(buf, ip, iend);
      if (lit > iend || ip + lit > iend || op + lit > cap || op > ip) return 0xffffffffu;
      lz4_smem_move<true>(buf, op, ip, lit);
      ip += lit; op += lit;
      if (ip >= iend) break;
      offset = (uint32_t)buf[ip] | ((uint32_t)buf[ip + 1] << 8);
      ip += 2;
      if (ml == 15u) ml += lz4_read_ext(buf, ip, iend);
      }
    const uint32_t mlen = ml + LZ4_MINMATCH;
    if (ip > iend || mlen > cap || offset == 0 || offset > op || op + mlen > cap) return 0xffffffffu;
    // next window: unread input, which this sequence's output never reaches (in-place margin)
    const uint32_t bnext = buf[ip + lane];
    if (short_lit && offset <= 32u && mlen > offset)
      { // Short-period match behind a short literal run (runs, interleaved index patterns).  The
        // pattern is the `offset` bytes before the match: its last `lit` bytes are the literals, which
        // this warp still holds in b (lanes 1..lit) - only older bytes are read from the buffer, so
        // nothing here waits for the literal stores above.
      const uint32_t inv = __shfl_sync(FULL, lane_inv, offset - 1u), S = __shfl_sync(FULL, lane_stride, offset - 1u);
      const uint32_t keep = offset == 1u ? 0u : 0xffffffffu;
      auto modo = [&](uint32_t t) { return (t - offset * __umulhi(t, inv)) & keep; };
      const uint32_t r0 = modo(lane), r1 = modo(lane + 32u);
      const int32_t first_lit = (int32_t)offset - (int32_t)lit;      // pattern index of the first literal (<= 0: literals only)
      uint32_t q0 = __shfl_sync(FULL, b, (1u + r0 - (uint32_t)first_lit) & 31u);
      uint32_t q1 = __shfl_sync(FULL, b, (1u + r1 - (uint32_t)first_lit) & 31u);
      const uint8_t* ms = buf + op - offset;
      if ((int32_t)r0 < first_lit) q0 = ms[r0];
      if ((int32_t)r1 < first_lit) q1 = ms[r1];
      if (lane < mlen) buf[op + lane] = (uint8_t)q0;
      if (lane + 32u < mlen) buf[op + lane + 32u] = (uint8_t)q1;
      if (mlen > 64u)
        {
        __syncwarp();
        lz4_period_rest(buf, op, mlen, S, modo);
        }
      }
    else
      {
      __syncwarp();                                 // literals of this sequence are visible
      lz4_match_warp(buf, op, offset, mlen);
      }
    op += mlen;
    b = bnext;
    batching = lit <= LZ4_BATCH_MAXLIT && mlen <= 32u;
    __syncwarp();
    }
  return op;
  }

// ---------------------------------------------------------------------------------------------
// K6: per-block LZ4 decode fused with the plane merge (trico_transpose_uint*_soa_to_aos,
// transpose_aos_to_soa.c:94-147).  CTA = WB warps with one in-place plane buffer each; a tile = one
// range of B elements = WB blocks, contiguous in the payload.
//   0. warp 0 runs one step ahead: ticket, block sizes, base offset (look-back), and a look at the
//      blocks themselves - a block that is exactly the run encoding of lz4_emit_run (one repeated
//      byte: the upper planes of index data, mostly) needs neither a buffer nor a decoder.  When two
//      consecutive tiles have at most WB/2 other planes each, the CTA takes them TOGETHER: the tile
//      time is set by the plane with the most sequences while the other warps wait at the merge
//      barrier, so two tiles in flight per CTA is twice the throughput.
//   1. every working warp stages its block into the tail of its buffer (16-byte cp.async copies)
//      and waits for nothing else
//   2. it decodes the plane in place (lz4_decode_inplace)
//   3. the CTA writes the merged elements with 16-byte stores, repeated bytes from registers.
// ---------------------------------------------------------------------------------------------
struct Lz4DecodeArgs
  {
  const uint8_t* sizes;
  const uint8_t* payload;
  uint64_t payload_bytes;
  uint64_t n;
  uint32_t nranges;
  int log2B;
  void* out;
  uint64_t* desc;
  uint32_t* ticket;
  uint32_t* status;        // set to 1 if any block was malformed
  };

struct Lz4TileInfo
  {
  uint64_t base;           // payload offset of the tile's first block
  uint32_t tile;
  uint32_t ok;             // sizes plausible and inside the payload
  uint32_t cmask;          // bit q: plane q is one repeated byte
  uint32_t cval[2];        // those bytes: byte q & 3 of word q >> 2
  uint32_t sz[8];
  };

template <int WB>
__global__ void __launch_bounds__(WB * 32, WB == 8 ? 2 : 3)      // what the shared memory of the plane buffers admits
lz4_decode_kernel(const Lz4DecodeArgs a)
  {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t B = 1u << a.log2B;
  const uint32_t pstride = lz4_inplace_stride(B);
  uint8_t* planes = smem_raw;
  constexpr uint32_t ALLP = (1u << WB) - 1u;
  constexpr uint32_t HALF = WB / 2;
  __shared__ Lz4TileInfo sh_info[2][2];            // [step parity][tile of the pair]
  __shared__ uint32_t sh_ntiles[2];
  __shared__ Lz4TileInfo sh_carry;                 // a tile that was looked at but could not be paired (warp 0 only)
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  bool has_carry = false;                          // warp 0, uniform

  // warp 0, all lanes: everything about tile t that does not need a plane buffer
  auto analyze = [&](uint32_t t, Lz4TileInfo* info)
    {
    uint32_t mysz = 0;
    if (lane < WB)
      {
      const uint8_t* sz = a.sizes + 2 * ((uint64_t)t * WB + lane);
      mysz = (uint32_t)sz[0] | ((uint32_t)sz[1] << 8);
      }
    uint32_t agg = mysz, pre = mysz;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
      {
      agg += __shfl_xor_sync(FULL, agg, o);                                   // lanes >= WB hold 0
      const uint32_t up = __shfl_up_sync(FULL, pre, o);
      if (lane >= (unsigned)o) pre += up;
      }
    agg = __shfl_sync(FULL, agg, 0);
    pre -= mysz;                                                               // bytes of the tile's blocks before this lane's
    const bool sizes_ok = __all_sync(FULL, lane >= WB || (mysz != 0 && mysz <= lz4_block_bound(B)));
    const uint64_t excl = lookback_exclusive(a.desc, t, agg);
    const bool ok = sizes_ok && excl + agg <= a.payload_bytes;
    const uint64_t lo = (uint64_t)t << a.log2B;
    const uint32_t cnt = (uint32_t)((a.n - lo < B) ? (a.n - lo) : B);
    uint32_t cmask = 0, cv0 = 0, cv1 = 0;
    constexpr int NB = 3;                                    // 32-byte rows of a run block that are looked at: blocks up to 20 KiB
    const uint32_t m = cnt >= 25u ? cnt - 10u : 15u, mext = (m - 15u) / 255u + 1u, rs = 10u + mext, last = (m - 15u) % 255u;
    if (ok && cnt >= 25u && rs <= 32u * NB)
      { // which blocks are the run encoding of lz4_emit_run?  Size first, then every byte; the bytes
        // of all candidate blocks are requested before any is looked at (one memory round trip)
      uint32_t bv[WB][NB], v[WB];
      bool cand[WB];
#pragma unroll
      for (int q = 0; q < WB; ++q)
        {
        cand[q] = __shfl_sync(FULL, mysz, q) == rs;
        const uint8_t* blk = a.payload + excl + __shfl_sync(FULL, pre, q);
        v[q] = 0;
        if (cand[q]) v[q] = blk[1];
#pragma unroll
        for (int r = 0; r < NB; ++r)
          {
          const uint32_t i = lane + 32u * r;
          bv[q][r] = 0;
          if (cand[q] && i < rs) bv[q][r] = blk[i];
          }
        }
#pragma unroll
      for (int q = 0; q < WB; ++q)
        {
        bool same = cand[q];
#pragma unroll
        for (int r = 0; r < NB; ++r)
          {
          const uint32_t i = lane + 32u * r;
          uint32_t e;
          if (i == 0) e = 0x1fu; else if (i == 1) e = v[q]; else if (i == 2) e = 1u; else if (i == 3) e = 0u;
          else if (i < 4u + mext) e = (i + 1u == 4u + mext) ? last : 255u;
          else if (i == 4u + mext) e = 0x50u; else e = v[q];
          if (i < rs && bv[q][r] != e) same = false;
          }
        if (__all_sync(FULL, same))
          {
          cmask |= 1u << q;
          if (q < 4) cv0 |= v[q] << (8 * q); else cv1 |= v[q] << (8 * (q - 4));
          }
        }
      }
    if (ok)
      { // the other blocks -> L2
      const uint8_t* p0 = a.payload + excl;
      for (uint32_t o = lane * 128u; o < agg; o += 32u * 128u) asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + o));
      }
    if (lane < WB) info->sz[lane] = mysz;
    if (lane == 0) { info->base = excl; info->tile = t; info->ok = ok; info->cmask = cmask; info->cval[0] = cv0; info->cval[1] = cv1; }
    __syncwarp();
    };
  auto take_ticket = [&]()
    {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1u);
    return __shfl_sync(FULL, t, 0);
    };
  // warp 0: the tile(s) of the next step
  auto fetch = [&](int par)
    {
    uint32_t n = 0;
    if (has_carry)
      {
      if (lane == 0) sh_info[par][0] = sh_carry;
      __syncwarp();
      has_carry = false; n = 1;
      }
    else
      {
      const uint32_t t = take_ticket();
      if (t < a.nranges) { analyze(t, &sh_info[par][0]); n = 1; }
      }
    if (n == 1 && HALF >= 1 && (uint32_t)__popc(~sh_info[par][0].cmask & ALLP) <= HALF)
      { // room for a second tile
      const uint32_t t = take_ticket();
      if (t < a.nranges)
        {
        analyze(t, &sh_carry);
        if ((uint32_t)__popc(~sh_carry.cmask & ALLP) <= HALF)
          {
          if (lane == 0) sh_info[par][1] = sh_carry;
          n = 2;
          }
        else has_carry = true;
        __syncwarp();
        }
      }
    if (lane == 0) sh_ntiles[par] = n;
    };

  if (warp == 0) fetch(0);
  int cur = 0;
  for (;;)
    {
    __syncthreads();
    const uint32_t ntiles = sh_ntiles[cur];
    if (ntiles == 0) break;

    // this warp's plane: alone, tile 0's plane `warp`; in a pair, the first HALF warps take tile 0's
    // planes that need decoding (in order), the others tile 1's
    const uint32_t j = (ntiles == 2 && warp >= HALF) ? 1u : 0u;
    const Lz4TileInfo& ti = sh_info[cur][j];
    const uint32_t work = ~ti.cmask & ALLP;
    uint32_t p = warp;
    bool active = (work >> warp) & 1u;
    if (ntiles == 2)
      {
      const uint32_t rank = warp - j * HALF;
      active = rank < (uint32_t)__popc(work);
      p = active ? __fns(work, 0, (int)rank + 1) : 0u;
      }
    const uint64_t lo_t = (uint64_t)ti.tile << a.log2B;
    const uint32_t cnt_t = (uint32_t)((a.n - lo_t < B) ? (a.n - lo_t) : B);
    uint32_t my_ip = 0, my_end = 0;
    if (active && ti.ok)
      { // 1. stage: the block goes to the tail of the buffer, at an offset congruent to its global
        //    address modulo 16 so that the body moves as 16-byte cp.async copies; whole vectors from
        //    the boundary below the block to the boundary above it (the few bytes copied in front of /
        //    behind the block land on free buffer space; nothing at or past the end of the payload
        //    is read: src-size operand)
      uint64_t off = ti.base;
      for (uint32_t q = 0; q < p; ++q) off += ti.sz[q];
      const uint32_t sz = ti.sz[p];
      const uint8_t* src = a.payload + off;
      const uint32_t al = (uint32_t)reinterpret_cast<uintptr_t>(src) & 15u;
      const uint32_t d0 = ((pstride - 32u - sz - al) & ~15u) + al;           // block occupies [d0, d0 + sz)
      const uint8_t* src16 = src - al;
      const uint32_t dst_s = (uint32_t)__cvta_generic_to_shared(planes + (size_t)warp * pstride + (d0 - al));
      const uint32_t nv = (al + sz + 15u) >> 4;
      const uint64_t left = a.payload_bytes - off + al;                       // bytes from src16 to the end of the payload
      for (uint32_t i = lane; i < nv; i += 32)
        {
        const uint64_t rem = left - 16ull * i;
        const uint32_t ssz = rem >= 16 ? 16u : (uint32_t)rem;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst_s + 16u * i), "l"(src16 + 16u * i), "r"(ssz) : "memory");
        }
      my_ip = d0; my_end = d0 + sz;
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // 2. decode in place
    if (active)
      {
      uint32_t got = 0xffffffffu;
      if (ti.ok) got = lz4_decode_inplace(planes + (size_t)warp * pstride, my_ip, my_end, cnt_t);
      if (got != cnt_t && lane == 0) *a.status = 1;
      }
    else if (!ti.ok && lane == 0) *a.status = 1;
    if (warp == 0) fetch(cur ^ 1);
    __syncthreads();

    // 3. merge: element i = bytes plane[q][i], q = 0..WB-1 (LSB first)
    for (uint32_t jj = 0; jj < ntiles; ++jj)
      {
      const Lz4TileInfo& tm = sh_info[cur][jj];
      const uint32_t cm = tm.cmask, wk = ~cm & ALLP;
      const uint64_t lo = (uint64_t)tm.tile << a.log2B;
      const uint32_t cnt = (uint32_t)((a.n - lo < B) ? (a.n - lo) : B);
      const uint8_t* src[WB];                                  // plane q's buffer (unused when it is a repeated byte)
      uint32_t cw[WB];                                         // the repeated byte in every byte of a word
#pragma unroll
      for (int q = 0; q < WB; ++q)
        {
        const uint32_t bi = ntiles == 2 ? jj * HALF + (uint32_t)__popc(wk & ((1u << q) - 1u)) : (uint32_t)q;
        src[q] = planes + (size_t)(bi < (uint32_t)WB ? bi : 0u) * pstride;
        cw[q] = ((tm.cval[q >> 2] >> (8 * (q & 3))) & 0xffu) * 0x01010101u;
        }
      auto word = [&](int q, uint32_t i) { return (cm >> q) & 1u ? cw[q] : reinterpret_cast<const uint32_t*>(src[q])[i]; };
      auto half = [&](int q, uint32_t i) { return (cm >> q) & 1u ? (cw[q] & 0xffffu) : (uint32_t)reinterpret_cast<const uint16_t*>(src[q])[i]; };
      auto byte_of = [&](int q, uint32_t i) { return (cm >> q) & 1u ? (uint8_t)cw[q] : src[q][i]; };
      uint8_t* gout = reinterpret_cast<uint8_t*>(a.out) + lo * WB;
      if ((reinterpret_cast<uintptr_t>(gout) & 15u) == 0)
        {
        constexpr int EPV = 16 / WB;
        const uint32_t nvec = cnt / EPV;
        uint4* g4 = reinterpret_cast<uint4*>(gout);
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x)
          {
          uint32_t w[4] = {0, 0, 0, 0};
          if constexpr (WB == 1)
            {
            w[0] = word(0, 4 * i); w[1] = word(0, 4 * i + 1); w[2] = word(0, 4 * i + 2); w[3] = word(0, 4 * i + 3);
            }
          else if constexpr (WB == 4)
            { // 4x4 byte transpose: one word of each plane -> four elements
            const uint32_t p0 = word(0, i), p1 = word(1, i), p2 = word(2, i), p3 = word(3, i);
            const uint32_t a01l = __byte_perm(p0, p1, 0x5140), a01h = __byte_perm(p0, p1, 0x7362);   // (p0.b0 p1.b0 p0.b1 p1.b1), (b2.. b3..)
            const uint32_t a23l = __byte_perm(p2, p3, 0x5140), a23h = __byte_perm(p2, p3, 0x7362);
            w[0] = __byte_perm(a01l, a23l, 0x5410); w[1] = __byte_perm(a01l, a23l, 0x7632);
            w[2] = __byte_perm(a01h, a23h, 0x5410); w[3] = __byte_perm(a01h, a23h, 0x7632);
            }
          else if constexpr (WB == 2)
            { // eight elements: one 8-byte group of each plane
            const uint32_t q0x = word(0, 2 * i), q0y = word(0, 2 * i + 1), q1x = word(1, 2 * i), q1y = word(1, 2 * i + 1);
            w[0] = __byte_perm(q0x, q1x, 0x5140); w[1] = __byte_perm(q0x, q1x, 0x7362);
            w[2] = __byte_perm(q0y, q1y, 0x5140); w[3] = __byte_perm(q0y, q1y, 0x7362);
            }
          else
            { // two elements: one 2-byte group of each of the 8 planes
            uint32_t h[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = half(q, i);
            const uint32_t a01 = __byte_perm(h[0], h[1], 0x5140), a23 = __byte_perm(h[2], h[3], 0x5140);   // e0.b0 e0.b1 e1.b0 e1.b1
            const uint32_t a45 = __byte_perm(h[4], h[5], 0x5140), a67 = __byte_perm(h[6], h[7], 0x5140);
            w[0] = __byte_perm(a01, a23, 0x5410); w[1] = __byte_perm(a45, a67, 0x5410);
            w[2] = __byte_perm(a01, a23, 0x7632); w[3] = __byte_perm(a45, a67, 0x7632);
            }
          __stcs(g4 + i, make_uint4(w[0], w[1], w[2], w[3]));
          }
        for (uint32_t i = nvec * EPV + threadIdx.x; i < cnt; i += blockDim.x)
          for (int q = 0; q < WB; ++q) gout[(size_t)i * WB + q] = byte_of(q, i);
        }
      else
        {
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x)
          for (int q = 0; q < WB; ++q) gout[(size_t)i * WB + q] = byte_of(q, i);
        }
      }
    cur ^= 1;
    }
  }

// ---------------------------------------------------------------------------------------------
// K6L: reference-format planes - ONE LZ4 block per whole plane (trico.c:346, :1101), up to 2 GB.
// A block is one serial chain of sequences, so a plane is decoded by one warp (block p = plane p);
// what the warp can do is keep the chain in shared memory:
//   * the compressed bytes pass through an 8 KiB window (16-byte cp.async fills);
//   * the output is produced in a flat buffer of 64 KiB of history + a 16 KiB segment: every
//     match source (offsets reach 65535 bytes back) is a shared-memory read, all copies are the
//     warp-wide ones of the chunked decoder (512 bytes per step, periodic fill for short offsets),
//     runs of short sequences go through the batch decoder (lz4_decode_batch);
//   * a full segment leaves as coalesced 16-byte stores and the history moves down by a segment.
// Sequences are cut at segment / window boundaries and continued after the flush / refill.
// The planes are merged by a second kernel (planes.cuh).
// ---------------------------------------------------------------------------------------------
struct Lz4LegacyDecodeArgs
  {
  const uint8_t* src[8];
  uint32_t src_len[8];
  uint8_t* planes;         // nplanes buffers of plane_stride bytes (16-byte aligned)
  uint64_t plane_stride;
  uint32_t raw_len;
  uint32_t* status;
  };

constexpr uint32_t LZ4_LEG_HIST = 65536, LZ4_LEG_SEG = 16384, LZ4_LEG_IN = 8192;
__host__ __device__ constexpr size_t lz4_legacy_smem() { return (size_t)LZ4_LEG_HIST + LZ4_LEG_SEG + 64 + LZ4_LEG_IN + 64; }

// warp copy between two shared-memory buffers, any alignment, non-overlapping
__device__ __forceinline__ void lz4_smem_copy(uint8_t* dst, const uint8_t* src, uint32_t n)
  {
  const unsigned lane = lane_id();
  if (n <= 64u)
    {
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    return;
    }
  const uint32_t head = (4u - ((uint32_t)__cvta_generic_to_shared(dst) & 3u)) & 3u;
  if (lane < head) dst[lane] = src[lane];
  const uint32_t nw = (n - head) >> 2;
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
  const uint8_t* sb = src + head;
  const uint32_t sal = (uint32_t)__cvta_generic_to_shared(sb) & 3u;
  const uint8_t* sbase = sb - sal;                         // 4-byte aligned
  for (uint32_t i = lane; i < nw; i += 32) dw[i] = smem_read32(sbase, sal + 4u * i);
  const uint32_t done = head + (nw << 2);
  if (done + lane < n) dst[done + lane] = src[done + lane];
  }

#ifndef TB200_HOST_EMU
__global__ void __launch_bounds__(32)
lz4_decode_legacy_kernel(const Lz4LegacyDecodeArgs a)
  {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* ob = smem_raw;                                            // [HIST | SEG | 64 spare]
  uint8_t* ib = smem_raw + LZ4_LEG_HIST + LZ4_LEG_SEG + 64;          // [IN | 64 spare]
  const unsigned lane = lane_id();
  const unsigned pl = blockIdx.x;
  const uint8_t* src = a.src[pl];
  const uint32_t src_len = a.src_len[pl];
  uint8_t* gout = a.planes + (size_t)pl * a.plane_stride;
  const uint32_t raw_len = a.raw_len;
  constexpr uint32_t CAP = LZ4_LEG_HIST + LZ4_LEG_SEG;

  uint32_t consumed = 0;          // compressed bytes before ib[0]
  uint32_t in_len = 0;            // valid bytes in ib
  uint32_t ip = 0;                // next unread byte in ib
  uint32_t seg_base = 0;          // output position of ob[HIST]
  uint32_t op = LZ4_LEG_HIST;     // next output byte in ob
  bool bad = false;

  // moves the unread input to the front of the window and fills the rest from global memory
  auto refill = [&]()
    {
    const uint32_t keep = in_len - ip;
    if (ip != 0u)
      {
      for (uint32_t i0 = 0; i0 < keep; i0 += 32)
        {
        uint32_t t = 0;
        if (i0 + lane < keep) t = ib[ip + i0 + lane];
        __syncwarp();
        if (i0 + lane < keep) ib[i0 + lane] = (uint8_t)t;
        __syncwarp();
        }
      consumed += ip; ip = 0; in_len = keep;
      }
    const uint32_t left = src_len - (consumed + in_len);
    const uint32_t take = min(left, LZ4_LEG_IN - in_len);
    const uint8_t* g = src + consumed + in_len;
    for (uint32_t i = lane; i < take; i += 32) ib[in_len + i] = __ldg(g + i);
    in_len += take;
    for (uint32_t i = lane; i < 64u; i += 32) ib[in_len + i] = 0;     // readable slack behind the window
    __syncwarp();
    };
  // the segment is full (or the block is finished): write it out, move the history down
  auto flush = [&](bool last)
    {
    uint32_t nout = op - LZ4_LEG_HIST;
    if (seg_base + nout > raw_len) { bad = true; nout = raw_len - seg_base; }     // never write past the plane
    const uint32_t nv = (reinterpret_cast<uintptr_t>(gout) & 15u) == 0 ? nout >> 4 : 0u;
    uint4* gv = reinterpret_cast<uint4*>(gout + seg_base);
    const uint4* sv = reinterpret_cast<const uint4*>(ob + LZ4_LEG_HIST);
    for (uint32_t i = lane; i < nv; i += 32) gv[i] = sv[i];
    for (uint32_t i = (nv << 4) + lane; i < nout; i += 32) gout[seg_base + i] = ob[LZ4_LEG_HIST + i];
    __syncwarp();
    if (!last)
      { // history: ob[SEG .. SEG + HIST) -> ob[0 .. HIST); every batch is read completely before it is written
      constexpr int UN = 4;
      uint4* v = reinterpret_cast<uint4*>(ob);
      for (uint32_t i0 = 0; i0 < LZ4_LEG_HIST / 16u; i0 += 32 * UN)
        {
        uint4 t[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) t[u] = v[LZ4_LEG_SEG / 16u + i0 + lane + 32 * u];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < UN; ++u) v[i0 + lane + 32 * u] = t[u];
        __syncwarp();
        }
      seg_base += nout; op = LZ4_LEG_HIST;
      }
    };
  // reads a length continuation at ip (lz4.c:1629-1649), refilling the window as often as it takes
  auto read_ext = [&]() -> uint32_t
    {
    uint32_t add = 0;
    for (;;)
      {
      if (in_len - ip < 64u && consumed + in_len < src_len) refill();
      const uint32_t avail = in_len - ip;
      if (avail == 0u) { bad = true; return add; }
      const uint32_t b = lane < avail ? ib[ip + lane] : 0u;
      const unsigned m = __ballot_sync(FULL, b != 255u);
      if (m == 0u) { add += 255u * 32u; ip += 32; continue; }         // (avail >= 32 here: a shorter window ends with a zero from the slack)
      const int e = __ffs((int)m) - 1;
      add += 255u * (uint32_t)e + __shfl_sync(FULL, b, e);
      ip += (uint32_t)e + 1u;
      return add;
      }
    };

  if (src_len == 0u) bad = true;
  else refill();
  bool batching = false;
  while (!bad)
    {
    if (in_len - ip < 1024u && consumed + in_len < src_len) refill();
    if (consumed + ip >= src_len) { bad = true; break; }              // a block ends with a literal-only sequence, never here
    const uint32_t floor = LZ4_LEG_HIST - min(seg_base, LZ4_LEG_HIST);
    if (batching)
      {
      const int nb = lz4_decode_batch(ib, ob, ip, in_len, op, CAP, floor);
      if (nb < 0) { bad = true; break; }
      if (nb < 4) batching = false;
      if (nb > 0) { __syncwarp(); if (op == CAP) flush(false); if (seg_base + (op - LZ4_LEG_HIST) > raw_len) bad = true; continue; }
      }
    const uint32_t token = ib[ip++];
    uint32_t lit = token >> 4, ml = token & 15u;
    if (lit == 15u) { lit += read_ext(); if (bad) break; }
    // literals, in pieces: what the window holds, what the segment takes
    uint32_t rem = lit;
    while (rem != 0u && !bad)
      {
      if (ip == in_len) { if (consumed + in_len >= src_len) { bad = true; break; } refill(); }
      if (op == CAP) flush(false);
      const uint32_t piece = min(rem, min(in_len - ip, CAP - op));
      if (seg_base + (op - LZ4_LEG_HIST) + piece > raw_len) { bad = true; break; }
      lz4_smem_copy(ob + op, ib + ip, piece);
      __syncwarp();
      ip += piece; op += piece; rem -= piece;
      }
    if (bad) break;
    if (consumed + ip >= src_len) break;                               // the last sequence has no match part
    if (in_len - ip < 2u) refill();
    if (in_len - ip < 2u) { bad = true; break; }
    const uint32_t offset = (uint32_t)ib[ip] | ((uint32_t)ib[ip + 1] << 8);
    ip += 2;
    if (ml == 15u) { ml += read_ext(); if (bad) break; }
    const uint32_t mlen = ml + LZ4_MINMATCH;
    const uint32_t outpos = seg_base + (op - LZ4_LEG_HIST);
    if (offset == 0u || offset > outpos || outpos + mlen > raw_len || outpos + mlen < outpos) { bad = true; break; }
    rem = mlen;
    while (rem != 0u)
      {
      if (op == CAP) flush(false);
      const uint32_t piece = min(rem, CAP - op);
      lz4_match_warp(ob, op, offset, piece);
      __syncwarp();
      op += piece; rem -= piece;
      }
    batching = lit <= LZ4_BATCH_MAXLIT && mlen <= 32u;
    }
  const uint32_t produced = seg_base + (op - LZ4_LEG_HIST);
  if (!bad) flush(true);
  if ((bad || produced != raw_len) && lane == 0) *a.status = 1;
  }

#endif // TB200_HOST_EMU

} // namespace tb200
