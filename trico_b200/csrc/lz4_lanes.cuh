// lz4_lanes.cuh - lane-parallel LZ4 block parser for planes made of MANY SHORT sequences (index
// planes of real meshes, attribute lists): 32 parsers per warp instead of one.
//
// The warp-cooperative matcher of lz4.cuh tests 32 positions per step but takes a few hundred
// instructions for every sequence it emits; on data with a sequence every 6..16 bytes that is two
// orders of magnitude below the memory system.  Here every LANE runs the serial greedy parse of
// LZ4_compress_generic (lz4.c:793-1181: hash, candidate, verify, extend, emit) on its own 64-byte
// piece of a 2 KiB "wave", all 32 in lock step, so an iteration of the loop advances 32 parses.
//   * candidates, nearest first: the wave's repeat offset (the longest match of the previous wave),
//     the lane's own piece (a 32-entry private table: exact sequential order, which is what runs and
//     short periods need), the shared table of the current wave (pieces of lower lanes are
//     earlier positions; entries of higher lanes are later ones and fail the `< position` test) and
//     the shared table of the waves before.  The two shared tables swap roles from wave to wave, so
//     the inserts of a wave never destroy what the previous wave left.  The longest candidate wins.
//   * one position of look-ahead (a longer match starting one byte later is preferred, as in lz4hc)
//   * matches end at the end of their piece: the parses are independent, and their sequences are
//     stitched in position order at the end of the wave: the first sequence of a piece takes the
//     literals the pieces before it left over, offsets come from a prefix sum over the pieces'
//     sizes, and the wave's bytes leave through a staging buffer as one coalesced copy.
// The result is a valid LZ4 block (lz4.c:189-196 end-of-block rules), decodable by
// LZ4_decompress_safe; tools/sim/lz4_lanes_model.py is the executable model of this parse.
// Blocks come here from lz4_encode_kernel, which hands over (does not finish) the blocks whose
// first kilobyte averages less than 16 bytes per sequence.
#pragma once

#include "lz4.cuh"

namespace tb200 {

#if defined(TB200_LZ4L_DEBUG) || defined(TB200_LZ4L_WATCHONLY)
#define TB200_LZ4L_PROBE 1
__device__ volatile unsigned int* g_lz4l_host;      // mapped page-locked host memory: survives a faulting kernel
__device__ unsigned int g_lz4l_dbg[8];
__device__ unsigned int g_lz4l_cur[1 << 16];        // chunk id a CTA is working on
#define LZ4L_REPORT(code, val) do { if (atomicCAS(&g_lz4l_dbg[0], 0u, (unsigned)(code)) == 0u) { g_lz4l_dbg[1] = (unsigned)(val); g_lz4l_dbg[2] = lane_id(); g_lz4l_dbg[3] = g_lz4l_cur[blockIdx.x & 0xffffu]; \
  if (g_lz4l_host) { g_lz4l_host[1] = (unsigned)(val); g_lz4l_host[2] = lane_id(); g_lz4l_host[3] = g_lz4l_cur[blockIdx.x & 0xffffu]; g_lz4l_host[0] = (unsigned)(code); __threadfence_system(); } } } while (0)
#define LZ4L_WATCH(counter, limit, code, val) do { if (++(counter) > (limit)) { LZ4L_REPORT(code, val); return 0; } } while (0)
__device__ unsigned int g_lz4l_target[2] = {0xffffffffu, 0u};      // chunk id and lane whose parse is traced into the host page
#define LZ4L_TRACE(slot, v) do { if (g_lz4l_host && g_lz4l_cur[blockIdx.x & 0xffffu] == g_lz4l_target[0] && lane_id() == g_lz4l_target[1]) { g_lz4l_host[16 + (slot)] = (unsigned)(v); } } while (0)
#define LZ4L_TRACE_FENCE() __threadfence_system()
#define LZ4L_CONVERGED(code) do { const unsigned am_ = __activemask(); if (am_ != FULL) LZ4L_REPORT(code, am_); } while (0)
#else
#define LZ4L_WATCH(counter, limit, code, val) do { } while (0)
#define LZ4L_CONVERGED(code) do { } while (0)
#define LZ4L_TRACE(slot, v) do { } while (0)
#define LZ4L_TRACE_FENCE() do { } while (0)
#endif
#ifdef TB200_LZ4L_DEBUG
#define LZ4L_CHECK(cond, code, val) do { if (!(cond)) LZ4L_REPORT(code, val); } while (0)
#define LZ4L_BAD(cond) (!(cond))
#else
#define LZ4L_CHECK(cond, code, val) do { } while (0)
#define LZ4L_BAD(cond) false
#endif

#ifndef LZ4L_EMIT
#define LZ4L_EMIT lz4_emit_bytes
#endif
#ifndef LZ4L_LAZY
#define LZ4L_LAZY 1
#endif
#ifndef LZ4L_REP
#define LZ4L_REP 1
#endif
#ifndef LZ4L_OWN
#define LZ4L_OWN 1
#endif
constexpr uint32_t LZ4L_S = 64;                 // bytes per lane and wave
constexpr uint32_t LZ4L_WAVE = 32 * LZ4L_S;
constexpr int LZ4L_HLOG = 11;                   // two shared tables of 2^11 u16 entries
constexpr int LZ4L_OWNBITS = 5;                 // private table: 32 entries per lane
constexpr uint32_t LZ4L_REGION = 68;            // bytes of shared memory per lane for its sequences (odd number of words: no bank conflicts)
constexpr uint32_t LZ4L_SHORTLIT = 32;          // first-sequence literal runs up to this go through the staging buffer
constexpr uint32_t LZ4L_STAGE = 32 * (LZ4L_S + 6 + LZ4L_SHORTLIT + 2);      // 3328 bytes: what a wave can stage
constexpr uint32_t LZ4L_PAD = 64;               // zeroed bytes readable past the block

__host__ __device__ constexpr size_t lz4_lanes_smem(uint32_t B)
  {
  return (size_t)B + LZ4L_PAD + 2 * sizeof(uint16_t) * ((size_t)1 << LZ4L_HLOG) + 32 * ((size_t)1 << LZ4L_OWNBITS) + 32 * LZ4L_REGION + LZ4L_STAGE;
  }

// number of bytes of one sequence
__device__ __forceinline__ uint32_t lz4_seq_bytes(uint32_t lit, uint32_t ml)
  {
  const uint32_t m = ml - LZ4_MINMATCH;
  return 1u + lit + (lit >= 15u ? (lit - 15u) / 255u + 1u : 0u) + 2u + (m >= 15u ? (m - 15u) / 255u + 1u : 0u);
  }

// writes one sequence (token, literal length bytes, literals from src[lit_start ..), offset, match
// length bytes) at o; returns the bytes written.  One lane, byte stores.
__device__ __forceinline__ uint32_t lz4_put_seq(uint8_t* o, const uint8_t* src, uint32_t lit_start, uint32_t lit, uint32_t off, uint32_t ml)
  {
  const uint32_t m = ml - LZ4_MINMATCH;
  uint32_t k = 0;
  o[k++] = (uint8_t)(((lit >= 15u ? 15u : lit) << 4) | (m >= 15u ? 15u : m));
  if (lit >= 15u) { uint32_t r = lit - 15u; while (r >= 255u) { o[k++] = 255; r -= 255u; } o[k++] = (uint8_t)r; }
  for (uint32_t j = 0; j < lit; ++j) o[k + j] = src[lit_start + j];
  k += lit;
  o[k++] = (uint8_t)off; o[k++] = (uint8_t)(off >> 8);
  if (m >= 15u) { uint32_t r = m - 15u; while (r >= 255u) { o[k++] = 255; r -= 255u; } o[k++] = (uint8_t)r; }
  return k;
  }

// one sequence written by the whole warp, byte-wise (long literal runs are rare in the blocks that come here)
template <typename DstPtr>
__device__ __forceinline__ uint32_t lz4_emit_bytes(DstPtr dst, uint32_t op, const uint8_t* src, uint32_t lit_start, uint32_t nlit, uint32_t offset, uint32_t mlen)
  {
  const unsigned lane = lane_id();
  const uint32_t m = mlen ? mlen - LZ4_MINMATCH : 0u;
  const uint32_t next = nlit >= 15u ? (nlit - 15u) / 255u + 1u : 0u;
  if (lane == 0) dst[op] = (uint8_t)(((nlit >= 15u ? 15u : nlit) << 4) | (m >= 15u ? 15u : m));
  for (uint32_t i = lane; i < next; i += 32) dst[op + 1u + i] = (i + 1u == next) ? (uint8_t)((nlit - 15u) % 255u) : (uint8_t)255;
  op += 1u + next;
  for (uint32_t i = lane; i < nlit; i += 32) dst[op + i] = src[lit_start + i];
  op += nlit;
  if (mlen)
    {
    const uint32_t mext = m >= 15u ? (m - 15u) / 255u + 1u : 0u;
    if (lane == 0) { dst[op] = (uint8_t)offset; dst[op + 1u] = (uint8_t)(offset >> 8); }
    for (uint32_t i = lane; i < mext; i += 32) dst[op + 2u + i] = (i + 1u == mext) ? (uint8_t)((m - 15u) % 255u) : (uint8_t)255;
    op += 2u + mext;
    }
  __syncwarp();
  return op;
  }

// Compresses src[0..n) (shared memory, LZ4L_PAD zero bytes readable past n, n <= 65535) into dst.
// T: 2 << LZ4L_HLOG u16; own: 32 << LZ4L_OWNBITS bytes; regions: 32 * LZ4L_REGION bytes; stage: LZ4L_STAGE bytes.
#ifdef LZ4L_NOINLINE
#define LZ4L_INLINE __noinline__
#else
#define LZ4L_INLINE __forceinline__
#endif
template <typename DstPtr>
__device__ LZ4L_INLINE uint32_t lz4_compress_lanes(const uint8_t* src, uint32_t n, DstPtr dst, uint16_t* T, uint8_t* own, uint8_t* regions, uint8_t* stage)
  {
  constexpr int HLOG = LZ4L_HLOG;
  const unsigned lane = lane_id();
  const unsigned lt = lanemask_lt();
  uint32_t op = 0, lastend = 0;                      // lastend: input position up to which sequences have been emitted
  if (n >= LZ4_MFLIMIT + 1)
    {
    for (uint32_t i = lane; i < (2u << HLOG); i += 32) T[i] = 0;
    for (uint32_t i = lane; i < (32u << LZ4L_OWNBITS) / 4u; i += 32) reinterpret_cast<uint32_t*>(own)[i] = 0;   // (what was here before must not steer the parse)
    __syncwarp();
    const uint32_t mflimit = n - LZ4_MFLIMIT, matchlimit = n - LZ4_LASTLITERALS;
    uint8_t* const reg = regions + lane * LZ4L_REGION;
    uint8_t* const myown = own + lane;               // entry e of this lane: myown[32 * e]
    uint32_t rep = 0, misses = 0;
    uint32_t wi = 0;
    for (uint32_t w0 = 0; w0 < n; w0 += LZ4L_WAVE, ++wi)
      {
      uint16_t* const Tc = T + ((wi & 1u) << HLOG);
      const uint16_t* const To = T + (((wi & 1u) ^ 1u) << HLOG);
      const uint32_t sub = w0 + lane * LZ4L_S;
      const bool mine = sub < n;
      const uint32_t end = mine ? min(sub + LZ4L_S, n) : sub;
      const uint32_t lim = min(end, matchlimit);     // a match of this lane ends here at the latest
      uint32_t pos = sub, anchor = sub;
      uint32_t nseq = 0, rbytes = 0;
      uint32_t f_q = 0, f_ml = 0, f_off = 0;         // the lane's first match: its literal run is completed when the wave is stitched
      bool pend = false;
      uint32_t p_q = 0, p_c = 0, p_ml = 0;           // look-ahead: a match found at pos - 1, not yet committed
      uint32_t best_ml = 15, best_off = rep;         // the wave's longest match (the next wave tries its offset first)
      const uint32_t wave_rep = rep;

      auto commit = [&](uint32_t q, uint32_t c, uint32_t ml)
        {
        const uint32_t off = q - c;
        LZ4L_TRACE(18, q); LZ4L_TRACE(19, (c << 16) | ml); LZ4L_TRACE(20, 0xdead0003u); LZ4L_TRACE_FENCE();
        LZ4L_CHECK(q >= anchor && q + ml <= end && c < q && ml >= 4u && q - anchor <= LZ4L_S && anchor >= sub, 1, (q << 16) | ml);
        if (nseq == 0) { f_q = q; f_ml = ml; f_off = off; }
        else
          {
          LZ4L_CHECK(rbytes + lz4_seq_bytes(q - anchor, ml) <= LZ4L_REGION, 2, rbytes);
          rbytes += lz4_put_seq(reg + rbytes, src, anchor, q - anchor, off, ml);
          }
        ++nseq;
        LZ4L_TRACE(20, 0xdead0004u); LZ4L_TRACE_FENCE();
        anchor = q + ml;
        if (ml > best_ml) { best_ml = ml; best_off = off; }
        };

      [[maybe_unused]] uint32_t watch_main = 0, watch_lost = 0, watch_stitch = 0;
      for (;;)
        {
        LZ4L_WATCH(watch_main, 8192u, 9, pos);
        __syncwarp();
        LZ4L_CONVERGED(20);
        // (No lane-divergent code between the end of an iteration and these votes: a pending match
        // that has nothing left to be compared with is committed in the parse section below, and
        // its lane probes again in the next iteration.)
        const bool can = mine && pos <= mflimit && pos + LZ4_MINMATCH <= end;
        if (__ballot_sync(FULL, can || pend) == 0) break;
        const unsigned probing = __ballot_sync(FULL, can);
        const uint32_t stride = 1u + (misses >> 6);
        // ---- candidates ----
        uint32_t seq = 0, h = 0, ho = 0;
        uint32_t c[4] = {0, 0, 0, 0};
        bool run[4] = {false, false, false, false};
        if (can)
          {
          seq = smem_read32(src, pos);
          h = (seq * 2654435761u) >> (32 - HLOG);
          ho = h >> (HLOG - LZ4L_OWNBITS);
          c[0] = pos - wave_rep;                                         // the wave's repeat offset
          run[0] = LZ4L_REP && wave_rep != 0u && pos >= wave_rep && smem_read32(src, c[0]) == seq;
          c[1] = sub + ((uint32_t)myown[32u * ho] & (LZ4L_S - 1u));     // own piece (entries of earlier waves are checked by content)
          run[1] = LZ4L_OWN && c[1] < pos && smem_read32(src, c[1]) == seq;
          c[2] = Tc[h];
          run[2] = c[2] < pos && smem_read32(src, c[2]) == seq;
          c[3] = To[h];
          run[3] = c[3] < pos && smem_read32(src, c[3]) == seq;
          LZ4L_CHECK(pos + 4u <= n && h < (1u << HLOG) && ho < 32u && c[1] < n && c[2] < n && c[3] < n, 5, pos);
          }
        // ---- inserts: after every look-up of the step; the highest position of a bucket wins ----
        LZ4L_CONVERGED(23);
        __syncwarp();
        if (can) { myown[32u * ho] = (uint8_t)(pos - sub); Tc[h] = (uint16_t)pos; }
        __syncwarp();
        for (;;)
          {
          const bool lost = can && Tc[h] < (uint16_t)pos;
          if (!__any_sync(FULL, lost)) break;
          LZ4L_WATCH(watch_lost, 100000u, 10, pos);
          if (lost) Tc[h] = (uint16_t)pos;
          __syncwarp();
          }
        LZ4L_TRACE(0, w0); LZ4L_TRACE(1, pos); LZ4L_TRACE(2, anchor); LZ4L_TRACE(3, (unsigned)can | (pend << 1) | (run[0] << 4) | (run[1] << 5) | (run[2] << 6) | (run[3] << 7));
        LZ4L_TRACE(4, c[0]); LZ4L_TRACE(5, c[1]); LZ4L_TRACE(6, c[2]); LZ4L_TRACE(7, c[3]); LZ4L_TRACE(8, lim); LZ4L_TRACE(9, end); LZ4L_TRACE(10, nseq); LZ4L_TRACE(11, rbytes);
        LZ4L_TRACE(12, p_q); LZ4L_TRACE(13, p_c); LZ4L_TRACE(14, p_ml); LZ4L_TRACE(15, watch_main); LZ4L_TRACE(20, 0xdead0001u); LZ4L_TRACE_FENCE();
        // ---- match lengths, four bytes per step, all candidates side by side; the longest wins, the nearest on a tie ----
        uint32_t ml = 0, mc = 0;
        if (can && (run[0] || run[1] || run[2] || run[3]))
          {
          const uint32_t room = lim - pos;                               // >= 4
          uint32_t len[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) len[k] = run[k] ? LZ4_MINMATCH : 0u;
          for (uint32_t i = LZ4_MINMATCH; i < room && (run[0] || run[1] || run[2] || run[3]); i += 4u)
            {
            const uint32_t a = smem_read32(src, pos + i);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (run[k])
                {
                const uint32_t x = a ^ smem_read32(src, c[k] + i);
                if (x) { len[k] += ((uint32_t)__ffs((int)x) - 1u) >> 3; run[k] = false; } else len[k] += 4u;
                }
            }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            {
            const uint32_t l = min(len[k], room);
            if (l > ml || (l == ml && l != 0u && c[k] > mc)) { ml = l; mc = c[k]; }
            }
          }
        // ---- the parse ----
        LZ4L_TRACE(16, ml); LZ4L_TRACE(17, mc); LZ4L_TRACE(20, 0xdead0002u); LZ4L_TRACE_FENCE();
        LZ4L_CONVERGED(24);
        bool hit = false, fin = false;                   // fin: the match in p_q/p_c/p_ml is final (ONE commit site: see DESIGN.md)
        if (can)
          {
          if (pend)
            {
            hit = true;
            if (ml > p_ml) { p_q = pos; p_c = mc; p_ml = ml; pos += 1u; }              // the later match is longer: the pending one's first byte becomes a literal
            else fin = true;
            }
          else if (ml >= LZ4_MINMATCH)
            {
            hit = true;
            uint32_t q = pos, cc = mc;
            while (q > anchor && cc > 0u && src[q - 1u] == src[cc - 1u]) { --q; --cc; ++ml; }   // backward extension (lz4.c:947-950)
            p_q = q; p_c = cc; p_ml = ml;
            if (LZ4L_LAZY && q == pos && pos + 1u <= mflimit && pos + 1u + LZ4_MINMATCH <= end) { pend = true; pos += 1u; }
            else fin = true;
            }
          else pos += stride;
          }
        else if (pend) { hit = true; fin = true; }       // nothing left to compare the pending match with
        if (fin) { commit(p_q, p_c, p_ml); pend = false; pos = anchor; }
#ifdef LZ4L_SYNC_EVERY
        __syncwarp();
#endif
        __syncwarp();
        LZ4L_CONVERGED(25);
        misses = __any_sync(FULL, hit) ? 0u : misses + (uint32_t)__popc(probing);
        }

      // ---- stitch the wave ----
      __syncwarp();
      LZ4L_CONVERGED(21);
      const bool has = nseq != 0u;
      const unsigned N = __ballot_sync(FULL, has);
      if (N != 0u)
        {
        const unsigned below = N & lt;
        uint32_t prev_end = __shfl_sync(FULL, anchor, below ? 31 - __clz((int)below) : 0);
        if (!below) prev_end = lastend;
        const uint32_t flit = has ? f_q - prev_end : 0u;                                // literals of the lane's first sequence
        LZ4L_CHECK(!has || (f_q >= prev_end && f_q < n && prev_end <= n && f_ml >= 4u && f_ml <= 2u * LZ4L_S && f_off >= 1u && f_off <= f_q), 8, (f_q << 16) | prev_end);
        const uint32_t fsz = has ? lz4_seq_bytes(flit, f_ml) : 0u;
        const unsigned big = __ballot_sync(FULL, has && flit > LZ4L_SHORTLIT);
        unsigned todo = N;
        while (todo != 0u)
          {
          LZ4L_WATCH(watch_stitch, 256u, 11, todo);
          LZ4L_CONVERGED(22);
          // the lanes up to the first one with a long literal run go through the staging buffer
          const unsigned b = big & todo;
          const unsigned seg = b ? (todo & ((1u << (__ffs((int)b) - 1)) - 1u)) : todo;
          if (seg != 0u)
            {
            const bool in = (seg >> lane) & 1u;
            const uint32_t sz = in ? fsz + rbytes : 0u;
            uint32_t incl = sz;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
              {
              const uint32_t up = __shfl_up_sync(FULL, incl, o);
              if (lane >= (unsigned)o) incl += up;
              }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            LZ4L_CHECK(total <= LZ4L_STAGE, 3, total);
            LZ4L_CHECK(op + total <= n + n / 255u + 16u, 4, op + total);
            LZ4L_CHECK(!in || (flit <= LZ4L_SHORTLIT && rbytes <= LZ4L_REGION && prev_end + flit <= n), 6, (flit << 16) | rbytes);
            if (LZ4L_BAD(total <= LZ4L_STAGE && op + total <= n + n / 255u + 16u)) return 0;
            if (in)
              {
              uint8_t* o = stage + (incl - sz);
              o += lz4_put_seq(o, src, prev_end, flit, f_off, f_ml);
              for (uint32_t j = 0; j < rbytes; ++j) o[j] = reg[j];
              }
            __syncwarp();
#ifndef LZ4L_NOEMIT
            for (uint32_t i = lane; i < total; i += 32) dst[op + i] = stage[i];
#endif
            __syncwarp();
            op += total;
            todo &= ~seg;
            }
          if (b)
            { // a first sequence behind a long literal run: the whole warp writes it
            const int L = __ffs((int)b) - 1;
            LZ4L_CHECK(__shfl_sync(FULL, prev_end, L) + __shfl_sync(FULL, flit, L) <= n && op + __shfl_sync(FULL, flit, L) + 70u <= n + n / 255u + 16u, 7, __shfl_sync(FULL, flit, L));
            if (LZ4L_BAD(__shfl_sync(FULL, prev_end, L) + __shfl_sync(FULL, flit, L) <= n && op + __shfl_sync(FULL, flit, L) + 70u <= n + n / 255u + 16u)) return 0;
#ifndef LZ4L_NOBIG
            op = LZ4L_EMIT(dst, op, src, __shfl_sync(FULL, prev_end, L), __shfl_sync(FULL, flit, L), __shfl_sync(FULL, f_off, L), __shfl_sync(FULL, f_ml, L));
#endif
            const uint32_t rb = __shfl_sync(FULL, rbytes, L);
            const uint8_t* rL = regions + (uint32_t)L * LZ4L_REGION;
            __syncwarp();
#ifndef LZ4L_NOEMIT
            for (uint32_t i = lane; i < rb; i += 32) dst[op + i] = rL[i];
#endif
            __syncwarp();
            op += rb;
            todo &= ~(1u << L);
            }
          }
        lastend = __shfl_sync(FULL, anchor, 31 - __clz((int)N));
        }
      // the next wave's repeat offset: that of this wave's longest match
        {
        uint32_t bm = best_ml, bo = best_off;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          {
          const uint32_t om = __shfl_xor_sync(FULL, bm, o), oo = __shfl_xor_sync(FULL, bo, o);
          if (om > bm || (om == bm && oo < bo)) { bm = om; bo = oo; }
          }
        rep = bo;
        }
      __syncwarp();
      }
    }
#ifdef LZ4L_NOFINAL
  return op;
#else
  if (lz4_not_worth(op + lz4_literal_run_bytes(n - lastend), n)) { op = 0; lastend = 0; }      // stored: the block becomes one literal run
  return lz4_emit(dst, op, src, lastend, n - lastend, 0, 0);
#endif
  }

// ---------------------------------------------------------------------------------------------
// Second pass of K5: the plane blocks lz4_encode_kernel handed over.  One warp per CTA (its shared
// memory is a block, two tables, the private tables, the lanes' regions and the staging buffer:
// ~31 KB, seven CTAs per SM); a warp pulls chunk ids from the list, extracts the plane again (the
// range is read a second time: these are a minority of the chunks) and parses it in lane mode.
// ---------------------------------------------------------------------------------------------
struct Lz4DenseArgs
  {
  const void* in;
  uint64_t n;
  int log2B;
  uint8_t* sizes;           // u16 LE per chunk
  uint8_t* scratch;         // one slot per chunk (as lz4_encode_kernel)
  uint32_t slot;
  const uint32_t* list;     // chunk ids handed over
  const uint32_t* count;    // how many
  uint32_t* ticket;         // zeroed
  };

template <int WB>
__global__ void __launch_bounds__(32)
lz4_encode_dense_kernel(const Lz4DenseArgs a)
  {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t B = 1u << a.log2B;
  const unsigned lane = lane_id();
  uint8_t* buf = smem_raw;
  uint16_t* T = reinterpret_cast<uint16_t*>(smem_raw + B + LZ4L_PAD);
  uint8_t* own = reinterpret_cast<uint8_t*>(T + (2u << LZ4L_HLOG));
  uint8_t* regions = own + (32u << LZ4L_OWNBITS);
  uint8_t* stage = regions + 32u * LZ4L_REGION;
  const uint32_t count = *a.count;
#ifdef LZ4L_ONE_PER_CTA
  for (uint32_t t = blockIdx.x; t < count; t += 0x7fffffffu)
    {
#else
  for (;;)
    {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1u);
    t = __shfl_sync(FULL, t, 0);
    if (t >= count) break;
#endif
    const uint64_t g = a.list[t];
#ifdef TB200_LZ4L_PROBE
    if (lane == 0) g_lz4l_cur[blockIdx.x & 0xffffu] = (unsigned)g;
    __syncwarp();
#endif
    const uint64_t k = g / WB;
    const uint32_t p = (uint32_t)(g % WB);
    const uint64_t lo = k << a.log2B;
    const uint32_t cnt = (uint32_t)((a.n - lo < B) ? (a.n - lo) : B);
    const uint8_t* gin = reinterpret_cast<const uint8_t*>(a.in) + lo * WB;
    // plane p of the range -> buf
    if ((reinterpret_cast<uintptr_t>(gin) & 15u) == 0)
      {
      constexpr int EPV = 16 / WB;
      const uint32_t nvec = cnt / EPV;
      const uint4* g4 = reinterpret_cast<const uint4*>(gin);
      const uint32_t sel2 = p | ((2u + p) << 4) | ((4u + p) << 8) | ((6u + p) << 12);
      constexpr int UN = 8;
      for (uint32_t i0 = lane; i0 < nvec; i0 += 32 * UN)
        {
        uint4 v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) if (i0 + 32u * u < nvec) v[u] = __ldg(g4 + i0 + 32u * u);
#pragma unroll
        for (int u = 0; u < UN; ++u)
          {
          const uint32_t i = i0 + 32u * u;
          if (i >= nvec) continue;
          if (WB == 1) reinterpret_cast<uint4*>(buf)[i] = v[u];
          else if (WB == 4) reinterpret_cast<uint32_t*>(buf)[i] = plane_bytes<4>(v[u], p);
          else if (WB == 2) reinterpret_cast<uint2*>(buf)[i] = make_uint2(__byte_perm(v[u].x, v[u].y, sel2), __byte_perm(v[u].z, v[u].w, sel2));
          else reinterpret_cast<uint16_t*>(buf)[i] = (uint16_t)plane_bytes<8>(v[u], p);
          }
        }
      for (uint32_t i = nvec * EPV + lane; i < cnt; i += 32) buf[i] = gin[(size_t)i * WB + p];
      }
    else
      for (uint32_t i = lane; i < cnt; i += 32) buf[i] = gin[(size_t)i * WB + p];
    for (uint32_t i = lane; i < LZ4L_PAD; i += 32) buf[cnt + i] = 0;
    __syncwarp();
    const uint32_t nbytes = lz4_compress_lanes(buf, cnt, a.scratch + g * a.slot, T, own, regions, stage);
    if (lane == 0)
      {
      uint8_t* sz = a.sizes + 2 * g;
      sz[0] = (uint8_t)nbytes; sz[1] = (uint8_t)(nbytes >> 8);
      }
    __syncwarp();
    }
  }

} // namespace tb200
