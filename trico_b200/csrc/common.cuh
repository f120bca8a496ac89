// common.cuh - device helpers shared by the trico_b200 kernels (sm_100a).
#pragma once

#ifdef TB200_HOST_EMU
#include "warp_emu.hpp"            // tools/sim: the device source compiled for the CPU, one std::thread per lane
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace tb200 {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
#ifdef TB200_HOST_EMU
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << lane_id()) - 1u; }
__device__ __forceinline__ unsigned lanemask_gt() { return lane_id() == 31u ? 0u : ~((2u << lane_id()) - 1u); }
#else
__device__ __forceinline__ unsigned lanemask_lt()
  {
  unsigned m; asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
  }
__device__ __forceinline__ unsigned lanemask_gt()
  {
  unsigned m; asm volatile("mov.u32 %0, %%lanemask_gt;" : "=r"(m)); return m;
  }
#endif

// ---------------------------------------------------------------------------------------------
// Single-pass chained scan ("decoupled look-back") over tile aggregates.
// One 64-bit descriptor per tile: bits 63..62 = state (0 empty, 1 aggregate, 2 inclusive prefix),
// bits 61..0 = byte count.  Value and state travel in one word, so relaxed accesses suffice.
// Tiles take their index from an atomic ticket so a tile can only wait on tiles that started.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t LB_AGG = 1ull << 62;
constexpr uint64_t LB_INC = 2ull << 62;
constexpr uint64_t LB_VAL = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t lb_load(const uint64_t* p)
  {
  uint64_t v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
  }
__device__ __forceinline__ void lb_store(uint64_t* p, uint64_t v)
  {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
  }

// Called by ONE full warp of the tile: publishes `aggregate` for `tile` (tile 0: as the inclusive
// prefix it is).
__device__ __forceinline__ void lookback_publish(uint64_t* desc, uint32_t tile, uint64_t aggregate)
  {
  if (lane_id() == 0) lb_store(desc + tile, (tile == 0 ? LB_INC : LB_AGG) | aggregate);
  }

// Called by ONE full warp of the tile, after lookback_publish: returns the exclusive prefix (sum
// of all earlier tiles' aggregates) to every lane and publishes the tile's inclusive prefix.
__device__ __forceinline__ uint64_t lookback_walk(uint64_t* desc, uint32_t tile, uint64_t aggregate)
  {
  const unsigned lane = lane_id();
  if (tile == 0) return 0;
  uint64_t excl = 0;
  int64_t look = (int64_t)tile - 1;           // lane 0 inspects `look`, lane i inspects look - i
  for (;;)
    {
    const int64_t idx = look - (int64_t)lane;
    uint64_t d = LB_INC;                      // tiles before 0 behave as an inclusive prefix of 0
    if (idx >= 0)
      {
      d = lb_load(desc + idx);
      while ((d >> 62) == 0) { __nanosleep(20); d = lb_load(desc + idx); }
      }
    const unsigned inc = __ballot_sync(FULL, (d >> 62) == 2);
    // nearest inclusive prefix (smallest lane index); everything nearer contributes its aggregate
    const unsigned stop = inc ? (unsigned)__ffs((int)inc) - 1u : 32u;
    uint64_t contrib = (lane <= stop) ? (d & LB_VAL) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(FULL, contrib, o);
    excl += contrib;
    if (inc) break;
    look -= 32;
    }
  if (lane == 0) lb_store(desc + tile, LB_INC | (excl + aggregate));
  return excl;
  }

// Publishes `aggregate` for `tile` and returns the exclusive prefix to every lane.
__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t* desc, uint32_t tile, uint64_t aggregate)
  {
  lookback_publish(desc, tile, aggregate);
  return lookback_walk(desc, tile, aggregate);
  }

// ---------------------------------------------------------------------------------------------
// Warp copy of `n` bytes from shared memory (src 4-byte aligned) to an arbitrarily aligned global
// destination using 16-byte stores for the aligned body.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_copy_smem_to_global(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n)
  {
  const unsigned lane = lane_id();
  uint32_t head = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
  if (head > n) head = n;
  if (lane < head) dst[lane] = src[lane];
  const uint32_t nvec = (n - head) >> 4;
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(src + (head & ~3u));
  const unsigned sh = (head & 3u) * 8u;
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
  for (uint32_t i = lane; i < nvec; i += 32)
    {
    const uint32_t* s = sw + 4 * i;
    const uint32_t w0 = s[0], w1 = s[1], w2 = s[2], w3 = s[3], w4 = s[4];
    uint4 o;
    o.x = __funnelshift_r(w0, w1, sh);
    o.y = __funnelshift_r(w1, w2, sh);
    o.z = __funnelshift_r(w2, w3, sh);
    o.w = __funnelshift_r(w3, w4, sh);
    dv[i] = o;
    }
  const uint32_t done = head + (nvec << 4);
  if (done + lane < n) dst[done + lane] = src[done + lane];
  }

__device__ __forceinline__ void store_u64_bytes(uint8_t* p, uint64_t v)
  {
#pragma unroll
  for (int b = 0; b < 8; ++b) p[b] = (uint8_t)(v >> (8 * b));
  }

} // namespace tb200
