// lz4_multi.cuh - K6M: K6 (lz4_decode_kernel, lz4.cuh) for streams whose tiles are HEAVY: planes made
// of very many short sequences (real index planes, colour planes, attribute lists).  There one plane
// of a tile keeps its warp busy for 10^5 cycles while the tile's other planes are constant or stored
// and their warps wait at the merge barrier.  This variant
//   * packs up to four tiles into one CTA step (as long as their decoded planes fit the WB plane
//     buffers), so every warp of the CTA has a dense block to decode;
//   * does not give a buffer to a STORED plane (one literal run) of a heavy tile: the merge reads its
//     bytes straight from the payload.
// The launcher picks it for streams that compress by less than 3x (device_api_lz4.inc); streams of
// long matches (synthetic grids: C2..C4 indices) stay with K6, whose per-tile bookkeeping is lighter.
// Same wire format, same checks, same Lz4DecodeArgs.
#pragma once

#include "lz4.cuh"

namespace tb200 {

struct Lz4MultiTileInfo
  {
  uint64_t base;           // payload offset of the tile's first block
  uint32_t tile;
  uint32_t ok;             // sizes plausible and inside the payload
  uint32_t cmask;          // bit q: plane q is one repeated byte
  uint32_t cval[2];        // those bytes: byte q & 3 of word q >> 2
  uint32_t smask;          // bit q: plane q is stored (one literal run) AND the merge reads it straight from the payload
  uint32_t sz[8];
  };
constexpr int LZ4_DEC_MAXTILES = 4;          // tiles a CTA decodes together at most
constexpr uint32_t LZ4_DEC_HEAVY = 2048;     // a block this large (bytes) has many sequences: its tile gives up the buffers of its stored planes

template <int WB>
__global__ void __launch_bounds__(WB * 32, WB == 8 ? 2 : (WB == 4 ? 4 : 8))      // 8 KiB plane blocks (colours, 2- and 8-byte lists) leave shared memory for this many CTAs
lz4_decode_multi_kernel(const Lz4DecodeArgs a)
  {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t B = 1u << a.log2B;
  const uint32_t pstride = lz4_inplace_stride(B);
  uint8_t* planes = smem_raw;
  constexpr uint32_t ALLP = (1u << WB) - 1u;
  constexpr int MAXT = WB < LZ4_DEC_MAXTILES ? WB : LZ4_DEC_MAXTILES;
  __shared__ Lz4MultiTileInfo sh_info[2][MAXT];         // [step parity][tile of the step]
  __shared__ uint32_t sh_ntiles[2];
  __shared__ Lz4MultiTileInfo sh_carry;                 // a tile that was looked at but could not be paired (warp 0 only)
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  bool has_carry = false;                          // warp 0, uniform

  // warp 0, all lanes: everything about tile t that does not need a plane buffer
  auto analyze = [&](uint32_t t, Lz4MultiTileInfo* info)
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
    // Stored planes: a block that is one literal run (token 0xf0, length bytes, the bytes: what the
    // encoder leaves of incompressible data) needs no decoder, and no buffer either if the merge
    // reads it from the payload.  That costs the merge unaligned global reads, so it is only done
    // for tiles with a HEAVY plane (many sequences: the warp that decodes it sets the tile's time,
    // and the freed warps and buffers take the heavy planes of the next tiles meanwhile).
    uint32_t smask = 0;
    if (ok && cnt >= 15u)
      {
      const uint32_t H = 2u + (cnt - 15u) / 255u;                              // token + length bytes
      const bool heavy = __any_sync(FULL, lane < WB && !((cmask >> lane) & 1u) && mysz >= LZ4_DEC_HEAVY && mysz != H + cnt);
      if (heavy && H <= 32u * NB)
        {
#pragma unroll
        for (int q = 0; q < WB; ++q)
          {
          if (__shfl_sync(FULL, mysz, q) != H + cnt) continue;
          const uint8_t* blk = a.payload + excl + __shfl_sync(FULL, pre, q);
          bool same = true;
#pragma unroll
          for (int r = 0; r < NB; ++r)
            {
            const uint32_t i = lane + 32u * r;
            if (i < H)
              {
              const uint32_t e = i == 0 ? 0xf0u : (i + 1u == H ? (cnt - 15u) % 255u : 255u);
              if (blk[i] != e) same = false;
              }
            }
          if (__all_sync(FULL, same)) smask |= 1u << q;
          }
        }
      }
    if (ok)
      { // the other blocks -> L2
      const uint8_t* p0 = a.payload + excl;
      for (uint32_t o = lane * 128u; o < agg; o += 32u * 128u) asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + o));
      }
    if (lane < WB) info->sz[lane] = mysz;
    if (lane == 0) { info->base = excl; info->tile = t; info->ok = ok; info->cmask = cmask; info->cval[0] = cv0; info->cval[1] = cv1; info->smask = smask; }
    __syncwarp();
    };
  auto take_ticket = [&]()
    {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1u);
    return __shfl_sync(FULL, t, 0);
    };
  // warp 0: the tiles of the next step - as many consecutive tiles as the CTA has warps (= plane
  // buffers) for: a tile needs one per plane that is neither a repeated byte nor read from the payload
  auto fetch = [&](int par)
    {
    uint32_t n = 0, used = 0;
    for (;;)
      {
      if (!has_carry)
        {
        const uint32_t t = take_ticket();
        if (t >= a.nranges) break;
        analyze(t, &sh_carry);
        has_carry = true;
        }
      const uint32_t need = (uint32_t)__popc(~(sh_carry.cmask | sh_carry.smask) & ALLP);
      if (n != 0 && (used + need > (uint32_t)WB || n >= (uint32_t)MAXT)) break;      // stays in sh_carry for the next step
      if (lane == 0) sh_info[par][n] = sh_carry;
      __syncwarp();
      has_carry = false;
      ++n; used += need;
      if (used >= (uint32_t)WB || n >= (uint32_t)MAXT) break;
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

    // this warp's plane: the planes that need decoding, tile by tile and in plane order, go to the warps in order
    uint32_t j = 0, before = 0;
    for (; j + 1u < ntiles; ++j)
      {
      const uint32_t nw = (uint32_t)__popc(~(sh_info[cur][j].cmask | sh_info[cur][j].smask) & ALLP);
      if (warp < before + nw) break;
      before += nw;
      }
    const Lz4MultiTileInfo& ti = sh_info[cur][j];
    const uint32_t work = ~(ti.cmask | ti.smask) & ALLP;
    const uint32_t rank = warp - before;
    const bool active = rank < (uint32_t)__popc(work);
    const uint32_t p = active ? __fns(work, 0, (int)rank + 1) : 0u;
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
      const Lz4MultiTileInfo& tm = sh_info[cur][jj];
      const uint32_t cm = tm.cmask, sm = tm.smask, wk = ~(cm | sm) & ALLP;
      uint32_t bufs_before = 0;
      for (uint32_t i = 0; i < jj; ++i) bufs_before += (uint32_t)__popc(~(sh_info[cur][i].cmask | sh_info[cur][i].smask) & ALLP);
      const uint64_t lo = (uint64_t)tm.tile << a.log2B;
      const uint32_t cnt = (uint32_t)((a.n - lo < B) ? (a.n - lo) : B);
      const uint8_t* src[WB];                                  // plane q's buffer (unused when it is a repeated byte)
      uint32_t cw[WB];                                         // the repeated byte in every byte of a word
#pragma unroll
      for (int q = 0; q < WB; ++q)
        {
        const uint32_t bi = bufs_before + (uint32_t)__popc(wk & ((1u << q) - 1u));
        src[q] = planes + (size_t)(bi < (uint32_t)WB ? bi : 0u) * pstride;
        cw[q] = ((tm.cval[q >> 2] >> (8 * (q & 3))) & 0xffu) * 0x01010101u;
        if ((sm >> q) & 1u)
          { // stored: the literal bytes in the payload, behind the token and the length bytes
          uint64_t off = tm.base;
          for (int i = 0; i < q; ++i) off += tm.sz[i];
          src[q] = a.payload + off + (2u + (cnt - 15u) / 255u);
          }
        }
      // byte i of a stored plane: any alignment, global memory (read-only path)
      auto gbyte = [&](const uint8_t* s8, uint32_t i) { return (uint32_t)__ldg(s8 + i); };
      auto gword = [&](const uint8_t* s8, uint32_t i)
        { // bytes [4i, 4i+4) of a stored plane: two aligned words around them
        const uintptr_t at = reinterpret_cast<uintptr_t>(s8) + 4u * (uintptr_t)i;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(at & ~(uintptr_t)3);
        const uint32_t sh = ((uint32_t)at & 3u) * 8u;
        const uint32_t w0 = __ldg(w);
        return sh ? __funnelshift_r(w0, __ldg(w + 1), sh) : w0;
        };
      auto word = [&](int q, uint32_t i) { return (cm >> q) & 1u ? cw[q] : (sm >> q) & 1u ? gword(src[q], i) : reinterpret_cast<const uint32_t*>(src[q])[i]; };
      auto half = [&](int q, uint32_t i) { return (cm >> q) & 1u ? (cw[q] & 0xffffu) : (sm >> q) & 1u ? (gbyte(src[q], 2u * i) | (gbyte(src[q], 2u * i + 1u) << 8)) : (uint32_t)reinterpret_cast<const uint16_t*>(src[q])[i]; };
      auto byte_of = [&](int q, uint32_t i) { return (cm >> q) & 1u ? (uint8_t)cw[q] : (sm >> q) & 1u ? (uint8_t)gbyte(src[q], i) : src[q][i]; };
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

} // namespace tb200
