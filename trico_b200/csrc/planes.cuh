// planes.cuh - standalone AoS <-> SoA and byte-plane kernels (sm_100a).
//
// The fused encode/decode kernels (fpc.cuh, lz4.cuh) do these transposes on the tile they already
// hold in shared memory; the standalone versions exist for the 14 exported trico_transpose_*
// symbols (/root/reference/trico/transpose_aos_to_soa.h:12-38) and for reference-format (v0)
// archives, whose planes are decoded whole before they can be merged.
#pragma once

#include "common.cuh"

namespace tb200 {

constexpr int TR_TILE = 1024;      // elements per CTA tile

struct TransposeArgs
  {
  void* aos;
  void* comp[8];           // component / plane arrays
  uint64_t n;
  };

// K1: AoS (n x NCOMP words) <-> NCOMP arrays of n words, staged through shared memory so both
// sides move as 16-byte coalesced vectors.  TO_SOA = trico_transpose_{xyz,uv}_aos_to_soa
// (transpose_aos_to_soa.c:8,:28,:48,:66), else the _soa_to_aos twins (:18,:38,:57,:75).
template <typename W, int NCOMP, bool TO_SOA>
__global__ void __launch_bounds__(256)
transpose_kernel(const TransposeArgs a)
  {
  __shared__ __align__(16) W tile[TR_TILE * NCOMP];
  constexpr int PER = 16 / sizeof(W);
  for (uint64_t t0 = (uint64_t)blockIdx.x * TR_TILE; t0 < a.n; t0 += (uint64_t)gridDim.x * TR_TILE)
    {
    const uint32_t cnt = (uint32_t)((a.n - t0 < TR_TILE) ? (a.n - t0) : TR_TILE);
    W* g = reinterpret_cast<W*>(a.aos) + t0 * NCOMP;
    const bool aos_vec = (reinterpret_cast<uintptr_t>(g) & 15u) == 0;
    if (TO_SOA)
      {
      const uint32_t nw = cnt * NCOMP;
      if (aos_vec)
        {
        const uint32_t nvec = nw / PER;
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<uint4*>(tile)[i] = reinterpret_cast<const uint4*>(g)[i];
        for (uint32_t i = nvec * PER + threadIdx.x; i < nw; i += blockDim.x) tile[i] = g[i];
        }
      else
        for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) tile[i] = g[i];
      __syncthreads();
#pragma unroll
      for (int c = 0; c < NCOMP; ++c)
        {
        W* o = reinterpret_cast<W*>(a.comp[c]) + t0;
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) o[i] = tile[i * NCOMP + c];
        }
      __syncthreads();
      }
    else
      {
#pragma unroll
      for (int c = 0; c < NCOMP; ++c)
        {
        const W* s = reinterpret_cast<const W*>(a.comp[c]) + t0;
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) tile[i * NCOMP + c] = s[i];
        }
      __syncthreads();
      const uint32_t nw = cnt * NCOMP;
      if (aos_vec)
        {
        const uint32_t nvec = nw / PER;
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(tile)[i];
        for (uint32_t i = nvec * PER + threadIdx.x; i < nw; i += blockDim.x) g[i] = tile[i];
        }
      else
        for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) g[i] = tile[i];
      __syncthreads();
      }
    }
  }

// byte planes: element i (WB bytes, little-endian) <-> plane p holds byte p of every element.
// trico_transpose_uint{16,32,64}_aos_to_soa / _soa_to_aos (transpose_aos_to_soa.c:84-147).
template <int WB, bool SPLIT>
__global__ void __launch_bounds__(256)
planes_kernel(const TransposeArgs a)
  {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * blockDim.x)
    {
    uint8_t* e = reinterpret_cast<uint8_t*>(a.aos) + i * WB;
    if (SPLIT)
      {
#pragma unroll
      for (int p = 0; p < WB; ++p) reinterpret_cast<uint8_t*>(a.comp[p])[i] = e[p];
      }
    else
      {
      uint8_t b[WB];
#pragma unroll
      for (int p = 0; p < WB; ++p) b[p] = reinterpret_cast<const uint8_t*>(a.comp[p])[i];
#pragma unroll
      for (int p = 0; p < WB; ++p) e[p] = b[p];
      }
    }
  }

} // namespace tb200
