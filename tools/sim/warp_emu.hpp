// warp_emu.hpp - a 32-thread host emulation of the warp-level CUDA constructs the LZ4 lane parser
// uses, so that the DEVICE SOURCE of trico_b200/csrc/lz4_lanes.cuh can be compiled with g++ and
// run on the CPU (tools/sim/lanes_emu.cpp): every lane is a std::thread, every *_sync intrinsic is
// a barrier with an exchange, and every call site checks that all 32 lanes arrived at the SAME
// site - a lane that takes a different path around a synchronisation point is reported at once.
// Test infrastructure only.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __global__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__
#define __constant__ static const
#define __align__(n_) __attribute__((aligned(n_)))

struct emu_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
using std::min;
using std::max;

static inline int __ffs(int x) { return x ? __builtin_ctz((unsigned)x) + 1 : 0; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
  {
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
  }

struct emu_warp
  {
  std::barrier<> bar{32};
  uint64_t val[32];
  int site[32];
  std::atomic<int> mismatches{0};
  };
static emu_warp* g_emu_warp = nullptr;

static inline void emu_arrive(int site)
  {
  emu_warp* w = g_emu_warp;
  const int l = threadIdx.x & 31;
  w->site[l] = site;
  w->bar.arrive_and_wait();
  for (int i = 0; i < 32; ++i)
    if (w->site[i] != site && l == 0 && w->mismatches.fetch_add(1) < 8)
      fprintf(stderr, "warp_emu: lane %d is at line %d while lane 0 is at line %d\n", i, w->site[i], site);
  }
static inline void emu_syncwarp(int site) { emu_arrive(site); g_emu_warp->bar.arrive_and_wait(); }
static inline unsigned emu_ballot(int site, bool p)
  {
  emu_warp* w = g_emu_warp;
  w->val[threadIdx.x & 31] = p;
  emu_arrive(site);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(w->val[i] != 0) << i;
  w->bar.arrive_and_wait();
  return m;
  }
template <typename T> static inline T emu_shfl(int site, T v, int src)
  {
  emu_warp* w = g_emu_warp;
  w->val[threadIdx.x & 31] = (uint64_t)v;
  emu_arrive(site);
  const T r = (T)w->val[src & 31];
  w->bar.arrive_and_wait();
  return r;
  }
static inline int emu_lane() { return (int)(threadIdx.x & 31); }
#define __syncwarp(...) emu_syncwarp(__LINE__)
#define __ballot_sync(mask_, pred_) emu_ballot(__LINE__, (pred_))
#define __any_sync(mask_, pred_) (emu_ballot(__LINE__, (pred_)) != 0)
#define __all_sync(mask_, pred_) (emu_ballot(__LINE__, (pred_)) == 0xffffffffu)
#define __shfl_sync(mask_, val_, src_) emu_shfl(__LINE__, (val_), (int)(src_))
#define __shfl_up_sync(mask_, val_, delta_) emu_shfl(__LINE__, (val_), (emu_lane() >= (int)(delta_) ? emu_lane() - (int)(delta_) : emu_lane()))
#define __shfl_xor_sync(mask_, val_, xor_) emu_shfl(__LINE__, (val_), (emu_lane() ^ (int)(xor_)))
static inline void __syncthreads() {}
static inline unsigned emu_match_any(int site, uint32_t v)
  {
  emu_warp* w = g_emu_warp;
  w->val[threadIdx.x & 31] = v;
  emu_arrive(site);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(w->val[i] == v) << i;
  w->bar.arrive_and_wait();
  return m;
  }
#define __match_any_sync(mask_, val_) emu_match_any(__LINE__, (val_))
static inline uint32_t __cvta_generic_to_shared(const void* p) { return (uint32_t)(uintptr_t)p; }
static inline long long clock64() { return 0; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) {}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline int __fns(unsigned mask, unsigned base, int offset)
  { int n = 0; for (unsigned i = base; i < 32; ++i) if ((mask >> i) & 1u) { if (++n == offset) return (int)i; } return -1; }
