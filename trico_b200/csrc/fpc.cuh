// fpc.cuh - FPC-style (FCM + DFCM predictor, XOR residual, leading-zero-byte code) float/double
// stream codec for sm_100a.
//
// Replaces trico_compress / trico_decompress (+ _double_precision) of the reference:
//   /root/reference/trico/floating_point_stream_compression.c:86-417 (float), :576-1164 (double).
//
// ENCODE is data-parallel inside a chunk.  Both predictors of the reference are finite-context:
//   FCM  context of value j  = top e1 bits of v[j-1]                       (fpc.c:76-79, :135)
//   DFCM context of value j  = ((t[j-2] & (2^(e2/2)-1)) << e2/2) ^ t[j-1], t = stride >> (bits-e2)
//                                                                          (fpc.c:81-84, :142)
// and the prediction is "what followed the most recent earlier element with the same context"
// (table store at the old hash, then lookup at the new one, fpc.c:134-136, :141-143).  A warp holds
// 32 CONSECUTIVE values, finds that earlier element with match.any + shfl inside the window and
// with a small table (per warp, shared memory) across windows, so the emitted bytes are identical
// to the serial reference while every instruction works on 32 values.
//
// DECODE is inherently serial inside a chunk (the context of value j needs the decoded v[j-1]), so
// the parallelism is across chunks: one lane per chunk, lane-private tables interleaved in shared
// memory (bank = lane), compressed bytes staged per warp by coalesced loads, output transposed
// through shared memory so global stores are contiguous.
#pragma once

#include "common.cuh"

#include <type_traits>

namespace tb200 {

template <typename W> struct FpcTraits;
template <> struct FpcTraits<uint32_t>
  {
  static constexpr int BITS = 32;     // value width
  static constexpr int GROUP = 8;     // values per code word            (fpc.c:12-18)
  static constexpr int CBITS = 3;     // bits per code
  static constexpr int HDR = 3;       // code word bytes, big-endian
  static constexpr int BASE2 = 4;     // codes > BASE2 select the DFCM residual (fpc.c:310)
  static constexpr int WBYTES = 4;
  };
template <> struct FpcTraits<uint64_t>
  {
  static constexpr int BITS = 64;
  static constexpr int GROUP = 2;     // fpc.c:421-425
  static constexpr int CBITS = 4;
  static constexpr int HDR = 1;
  static constexpr int BASE2 = 8;     // fpc.c:979
  static constexpr int WBYTES = 8;
  };

// worst-case bytes of one chunk payload (groups only, no 5-byte stream header)
__host__ __device__ constexpr uint32_t fpc_chunk_bound(uint32_t values, int wbytes)
  {
  return wbytes == 4 ? values * 4u + 3u * ((values + 7u) / 8u) + 8u
                     : values * 8u + ((values + 1u) / 2u) + 2u;
  }

// words between consecutive rows of the staging tile (row = chunk range = lane).  Odd, so that the
// 32 lanes of a warp, each on the same word of its own row, hit 32 different banks.
__host__ __device__ constexpr int fpc_stage_row_words(int sb, int ncomp, int wbytes) { return sb * ncomp * (wbytes / 4) + 1; }
// the encoder fills its tile with 16-byte cp.async copies, so its rows stay 16-byte aligned (and
// its per-value reads see a 4-way bank conflict; 4-byte copies into odd rows cost more than that)
__host__ __device__ constexpr int fpc_stage_row_words_enc(int sb, int ncomp, int wbytes) { return sb * ncomp * (wbytes / 4) + 4; }

// significant bytes of x (0 for x = 0): bfind gives the index of the leading one, 0xffffffff for zero
__device__ __forceinline__ int sig_bytes(uint32_t x)
  {
  uint32_t b; asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(x));
  return (int)((b + 8u) >> 3);
  }
__device__ __forceinline__ int sig_bytes(uint64_t x) { return (71 - __clzll((long long)x)) >> 3; }

// byte store to shared memory through its 32-bit window address (one STS, no generic-address
// arithmetic per store) or to a generic address
template <bool OUT_SHARED>
__device__ __forceinline__ void fpc_st8(uint8_t* gp, uint32_t sp, uint32_t off, uint32_t v)
  {
  if (OUT_SHARED) asm volatile("st.shared.u8 [%0], %1;" :: "r"(sp + off), "r"(v) : "memory");
  else gp[off] = (uint8_t)v;
  }

// ---------------------------------------------------------------------------------------------
// Warp-cooperative encoder of one chunk (or, with cnt = whole stream, of a reference v0 stream).
//   src      values of this component: element j lives at src[j * stride]
//   out      destination for the groups (code words + residual bytes); shared memory when
//            OUT_SHARED, else any address space
//   T1, T2   per-warp predictor tables with (1<<e1) / (1<<e2) entries, zeroed here
// Returns the number of bytes written.  All 32 lanes must call it.
//
// Byte offsets inside a window need no scan: the code words sit at fixed lanes (one per GROUP
// values), so a lane's offset is HDR * (groups started so far) plus the residual bytes of the
// lanes before it, and that sum is three (four) ballots of the bits of nb and population counts.
// ---------------------------------------------------------------------------------------------
// CONTINUE: the chain does not start here - the tables hold the predictor state in front of
// src[0] and (v_before, ta_before, tb_before) are the previous value and the stride classes of
// the two previous elements (the tile-parallel v0 encoder below).
template <typename W, bool OUT_SHARED, typename SrcPtr, bool CONTINUE = false>
__device__ __forceinline__ uint32_t fpc_encode_warp(SrcPtr src, uint32_t stride, uint32_t cnt,
                                                    uint8_t* out, W* T1, W* T2, int e1, int e2,
                                                    W v_before = 0, uint32_t ta_before = 0, uint32_t tb_before = 0)
  {
  using TR = FpcTraits<W>;
  const unsigned lane = lane_id();
  const unsigned lt = lanemask_lt(), gt = lanemask_gt();
  if (!CONTINUE)
    {
    for (uint32_t i = lane; i < (1u << e1); i += 32) T1[i] = 0;
    for (uint32_t i = lane; i < (1u << e2); i += 32) T2[i] = 0;
    }
  __syncwarp();

  const int h = e2 >> 1;
  const uint32_t lowmask = (1u << h) - 1u;
  const uint32_t padlimit = (cnt + TR::GROUP - 1) / TR::GROUP * TR::GROUP;
  const uint32_t out_s = OUT_SHARED ? (uint32_t)__cvta_generic_to_shared(out) : 0u;
  const uint32_t gl = lane & (TR::GROUP - 1);               // position inside the group
  const uint32_t hdr_before = TR::HDR * (lane / TR::GROUP + 1);   // code-word bytes up to and including this lane's group
  W carry_v = CONTINUE ? v_before : (W)0;
  uint32_t carry_ta = CONTINUE ? ta_before : 0u, carry_tb = CONTINUE ? tb_before : 0u;      // t[j-1], t[j-2] entering the window
  uint32_t obase = 0;

  W vnext = lane < cnt ? (W)src[(size_t)lane * stride] : (W)0;      // loads run two windows ahead
  W vnext2 = lane + 32u < cnt ? (W)src[(size_t)(lane + 32u) * stride] : (W)0;
  for (uint32_t i0 = 0; i0 < cnt; i0 += 32)
    {
    const uint32_t j = i0 + lane;
    const bool full = i0 + 32 <= cnt;       // warp-uniform: every lane holds a value
    const bool act = full || j < cnt;
    const W v = vnext;
    vnext = vnext2;
    vnext2 = j + 64u < cnt ? (W)src[(size_t)(j + 64u) * stride] : (W)0;
    W vprev = __shfl_up_sync(FULL, v, 1);
    if (lane == 0) vprev = carry_v;

    // FCM: context = top e1 bits of the previous value (initial hash 0 == context of a zero value)
    const uint32_t c1 = (uint32_t)(vprev >> (TR::BITS - e1));
    const unsigned m1 = __match_any_sync(FULL, c1);
    const unsigned early1 = m1 & lt;
    const int src1 = 31 - __clz((int)early1);                     // -1 when there is none (shfl result unused)
    const W p1s = __shfl_sync(FULL, v, src1);
    const W p1t = T1[c1];
    const W x1 = v ^ (early1 ? p1s : p1t);

    // DFCM: context from the two previous strides
    const W s = v - vprev;
    const uint32_t t = (uint32_t)(s >> (TR::BITS - e2));
    uint32_t ta = __shfl_up_sync(FULL, t, 1);
    uint32_t tb = __shfl_up_sync(FULL, t, 2);
    if (lane == 0) { ta = carry_ta; tb = carry_tb; }
    if (lane == 1) { tb = carry_ta; }
    const uint32_t c2 = ((tb & lowmask) << h) ^ ta;
    const unsigned m2 = __match_any_sync(FULL, c2);
    const unsigned early2 = m2 & lt;
    const int src2 = 31 - __clz((int)early2);
    const W p2s = __shfl_sync(FULL, s, src2);
    const W p2t = T2[c2];
    const W x2 = v ^ (vprev + (early2 ? p2s : p2t));

    // code selection, fpc.c:146-189 / :635-782
    const int n1 = sig_bytes(x1);
    int n2 = sig_bytes(x2); if (n2 == 0) n2 = 1;
    const bool use2 = (n1 >= 2) && (n2 < n1);
    uint32_t code = use2 ? TR::BASE2 + n2 : n1;
    uint32_t nb = use2 ? n2 : n1;
    W x = use2 ? x2 : x1;
    unsigned actmask = FULL;
    if (!full)
      { // last window: pad slots of the last group get code 1 + one zero byte (fpc.c:196-204, :789-794)
      if (!act)
        {
        const bool pad = j < padlimit;
        code = pad ? 1 : 0; nb = pad ? 1 : 0; x = 0;
        }
      actmask = __ballot_sync(FULL, act);
      }

    // table update: the last active element of every context wins (what a serial pass leaves)
    if (act && ((m1 & gt & actmask) == 0)) T1[c1] = v;
    if (act && ((m2 & gt & actmask) == 0)) T2[c2] = s;

    // code word of this lane's group (every lane of the group ends up with it)
    uint32_t bc = code << (TR::CBITS * gl);
#pragma unroll
    for (int o = 1; o < TR::GROUP; o <<= 1) bc |= __shfl_xor_sync(FULL, bc, o);

    // byte offsets from the bit planes of nb
    const unsigned b0 = __ballot_sync(FULL, nb & 1u), b1 = __ballot_sync(FULL, nb & 2u), b2 = __ballot_sync(FULL, nb & 4u);
    uint32_t pre = __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
    uint32_t sum = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
    if (TR::WBYTES == 8)
      {
      const unsigned b3 = __ballot_sync(FULL, nb & 8u);
      pre += 8u * __popc(b3 & lt); sum += 8u * __popc(b3);
      }
    const uint32_t ngroups = full ? 32u / TR::GROUP : (min(cnt - i0, 32u) + TR::GROUP - 1) / TR::GROUP;
    // code word: byte k of group g is written by lane g*GROUP + k at (bytes before the group) + k
    const uint32_t gpre = __shfl_sync(FULL, pre, lane & ~(TR::GROUP - 1));
    if (gl < (uint32_t)TR::HDR && j < padlimit)
      fpc_st8<OUT_SHARED>(out, out_s, obase + hdr_before - TR::HDR + gpre + gl, (bc >> (8 * (TR::HDR - 1 - gl))) & 0xffu);
    // residual bytes, most significant first
    if (nb)
      {
      const uint32_t at = obase + hdr_before + pre;
      const W xs = x << (8 * (TR::WBYTES - nb));                  // left-justified: byte b is at a fixed position
#pragma unroll
      for (int b = 0; b < TR::WBYTES; ++b)
        if ((uint32_t)b < nb) fpc_st8<OUT_SHARED>(out, out_s, at + b, (uint32_t)(xs >> (8 * (TR::WBYTES - 1 - b))) & 0xffu);
      }

    obase += TR::HDR * ngroups + sum;
    carry_v = __shfl_sync(FULL, v, 31);
    carry_tb = __shfl_sync(FULL, t, 30);
    carry_ta = __shfl_sync(FULL, t, 31);
    __syncwarp();
    }
  return obase;
  }

// arguments of the chunked encoder (K3, fpc_encode_lanes_kernel below)
struct FpcEncodeArgs
  {
  const void* in;          // device, AoS: n * ncomp words
  uint64_t n;              // values per component
  uint32_t nranges;        // ceil(n / S)
  uint32_t ntiles;
  int log2S, e1, e2;
  uint8_t* sizes;          // u16 LE [nranges * ncomp] (may be unaligned inside an archive)
  uint8_t* payload;        // chunk payloads, packed
  uint8_t* total_field;    // 8 bytes (unaligned): payload byte count, little-endian (stream header)
  uint64_t* total;         // aligned device scalar with the same value
  uint64_t* desc;          // look-back descriptors [ntiles], zeroed
  uint32_t* ticket;        // zeroed
  uint8_t* scratch;        // lane-per-chunk kernel: one slot per chunk in flight (grid * chunks per tile)
  uint32_t slot;           // bytes per scratch slot (multiple of 16, >= chunk bound + 16)
  };

// ---------------------------------------------------------------------------------------------
// Legacy (reference v0) stream encoder: one warp per component stream, whole stream as one chain,
// same warp routine with cnt = n.  Tables live in global memory when they do not fit in shared
// memory ((20,20) doubles: 2 x 8 MiB, fpc.c:588-594).  Output bytes go straight to global memory.
// ---------------------------------------------------------------------------------------------
struct FpcLegacyEncodeArgs
  {
  const void* in;          // device: component c, element j at in[(j * stride + c)]
  uint32_t n;
  uint32_t stride;
  int nstreams;            // components; stream c is encoded by block c
  int e1, e2;
  uint8_t* out;            // nstreams slots of out_stride bytes: 5-byte header + groups
  uint64_t out_stride;
  uint32_t* nbytes;        // [nstreams]
  void* gtables;           // nullptr, or nstreams * ((1<<e1)+(1<<e2)) words of global scratch
  };

template <typename W>
__global__ void __launch_bounds__(32)
fpc_encode_legacy_kernel(const FpcLegacyEncodeArgs a)
  {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const unsigned c = blockIdx.x, lane = lane_id();
  const size_t tw = ((size_t)1 << a.e1) + ((size_t)1 << a.e2);
  W* T1 = a.gtables ? reinterpret_cast<W*>(a.gtables) + (size_t)c * tw : reinterpret_cast<W*>(smem_raw);
  W* T2 = T1 + ((size_t)1 << a.e1);
  uint8_t* out = a.out + (size_t)c * a.out_stride;
  if (lane == 0)
    {
    out[0] = (uint8_t)(((a.e1 >> 1) << 4) | (a.e2 >> 1));                 // fpc.c:120
    out[1] = (uint8_t)(a.n >> 24); out[2] = (uint8_t)(a.n >> 16);          // fpc.c:123-126
    out[3] = (uint8_t)(a.n >> 8);  out[4] = (uint8_t)a.n;
    }
  const W* src = reinterpret_cast<const W*>(a.in) + c;
  uint32_t nb;
  if (a.n == 0)
    { // the reference emits one group of pad slots here (slot 0 from uninitialised stack, fpc.c:196-204)
    using TR = FpcTraits<W>;
    if (lane == 0)
      {
      uint32_t bc = 0;
      for (int g = 0; g < TR::GROUP; ++g) bc |= 1u << (TR::CBITS * g);
      uint8_t* p = out + 5;
      if (TR::HDR == 3) { p[0] = (uint8_t)(bc >> 16); p[1] = (uint8_t)(bc >> 8); p[2] = (uint8_t)bc; } else p[0] = (uint8_t)bc;
      for (int g = 0; g < TR::GROUP; ++g) p[TR::HDR + g] = 0;
      }
    nb = TR::HDR + TR::GROUP;
    }
  else
    nb = fpc_encode_warp<W, false>(src, a.stride, a.n, out + 5, T1, T2, a.e1, a.e2);
  if (lane == 0) a.nbytes[c] = nb + 5;
  }

// ---------------------------------------------------------------------------------------------
// K3S: chunked encode of SMALL streams, one WARP per chunk.  K3 (below) gives every lane a whole
// chunk: 512 serial values per lane, ~150 us however few chunks the stream has - fine when a
// stream fills the machine with lanes, all latency when it is the 117 chunks of a 20,000-vertex
// mesh (C5: thousands of such streams).  Here a warp runs the warp-cooperative encoder
// (fpc_encode_warp: 32 values per step, the same bytes) on its chunk: 16 steps instead of 512
// values, the blocks go to scratch slots and lz4_assemble_kernel (the LZ4 path's assembly: sizes ->
// offsets -> copy) lays them out.  Picked by the launcher for streams of at most 8192 chunks.
// ---------------------------------------------------------------------------------------------
struct FpcChunksArgs
  {
  const void* in;          // device, AoS: n * ncomp words
  uint64_t n;              // values per component
  uint32_t nranges;
  int log2S, e1, e2, ncomp;
  uint8_t* sizes;          // u16 LE [nranges * ncomp]
  uint8_t* scratch;        // one slot per chunk
  uint32_t slot;
  };

constexpr int FPC_CHUNKS_WARPS = 4;

template <typename W>
__global__ void __launch_bounds__(FPC_CHUNKS_WARPS * 32)
fpc_encode_chunks_kernel(const FpcChunksArgs a)
  {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t tw = (1u << a.e1) + (1u << a.e2);
  const size_t per_warp = (((size_t)tw * sizeof(W) + 15) & ~(size_t)15) + a.slot + 16;
  uint8_t* my = smem_raw + (size_t)warp * per_warp;
  W* T1 = reinterpret_cast<W*>(my);
  uint8_t* ob = my + (((size_t)tw * sizeof(W) + 15) & ~(size_t)15);
  const uint64_t g = (uint64_t)blockIdx.x * FPC_CHUNKS_WARPS + warp;
  if (g >= (uint64_t)a.nranges * a.ncomp) return;
  const uint64_t k = g / a.ncomp;
  const uint32_t c = (uint32_t)(g % a.ncomp);
  const uint64_t lo = k << a.log2S;
  const uint32_t S = 1u << a.log2S;
  const uint32_t cnt = (uint32_t)(a.n - lo < S ? a.n - lo : S);
  const W* src = reinterpret_cast<const W*>(a.in) + lo * a.ncomp + c;
  const uint32_t nb = fpc_encode_warp<W, true>(src, (uint32_t)a.ncomp, cnt, ob, T1, T1 + (1u << a.e1), a.e1, a.e2);
  __syncwarp();
  uint4* sv = reinterpret_cast<uint4*>(a.scratch + g * a.slot);
  for (uint32_t v4 = lane; v4 < (nb + 15u) >> 4; v4 += 32) sv[v4] = *reinterpret_cast<const uint4*>(ob + 16u * v4);
  if (lane == 0) { a.sizes[2 * g] = (uint8_t)nb; a.sizes[2 * g + 1] = (uint8_t)(nb >> 8); }
  }

// ---------------------------------------------------------------------------------------------
// K3L: TILE-PARALLEL encoder of reference-format (v0) streams - byte-identical to trico_compress
// (fpc.c:86-210), many warps on ONE chain.
//
// The chain of a v0 stream runs through the predictor tables, and both predictors are finite-context
// (see DESIGN.md): the prediction for an element is the value (stride) that followed the MOST
// RECENT EARLIER element with the same context.  Inside a tile of FPC_V0_TILE values
// fpc_encode_warp finds that element itself; what a tile needs from the past is the table as it
// stands in front of it - for every context the last writer among all earlier tiles.  "Last writer
// wins" is associative, so that state is a scan over tiles:
// A warp takes a RUN of consecutive tiles (the unit of the table scan; 1..8 tiles, as many as still
// leave every warp of the grid a few runs):
//   1. it computes the run's LOCAL table (last writer per context inside the run) and the mask of
//      contexts it wrote, and publishes both;
//   2. look-back: walking the earlier runs of the stream it fills the contexts it has not seen yet
//      from their local tables until it meets a run whose INCLUSIVE table (the full state behind
//      that run) is published, which completes the state; it then publishes its own inclusive table;
//   3. it encodes its tiles one after the other with fpc_encode_warp - tables initialised to the
//      state from 2. and simply carried from tile to tile, the chain's registers (previous value,
//      stride classes of the two previous elements) read from the input;
//   4. per tile, a second look-back over byte counts gives the offset in the stream.
// Runs take their number from a ticket, so a run only ever waits for runs that have started.
// (With one tile per run a tile visited ~25 predecessors on average, 4.3 KB each: the scan, not the
// encoder, set the speed.)
// Tables of (e1, e2) = (4, 10) floats: 1040 words; the kernel takes any table that fits shared
// memory twice (the (20, 20) tables of v0 doubles do not: those streams keep the one-warp kernel).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_copy_global(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n);      // below

#ifndef TB200_FPC_V0_TILE
#define TB200_FPC_V0_TILE 512      // 2048: 193 GB/s, 1024: 230, 512: 250 on C2 (shared memory per warp sets the occupancy), 4096: 143
#endif
constexpr uint32_t FPC_V0_TILE = TB200_FPC_V0_TILE;
constexpr int FPC_V0_WARPS = 4;

struct FpcV0TileArgs
  {
  const void* in;          // device: component c, element j at in[(j * stride + c)]
  uint32_t n;
  uint32_t stride;
  int nstreams;
  int e1, e2;
  uint8_t* out;            // nstreams slots of out_stride bytes: 5-byte header + groups
  uint64_t out_stride;
  uint32_t* nbytes;        // [nstreams]
  uint32_t ntiles;         // per stream
  uint32_t run;            // tiles per run: a warp takes a run of consecutive tiles, the table scan links runs
  uint32_t nruns;          // per stream
  uint32_t* ticket;        // zeroed
  uint64_t* desc;          // [nstreams * ntiles], zeroed: byte-count look-back (per tile)
  uint32_t* state;         // [nstreams * nruns], zeroed: 0 nothing, 1 local table, 2 inclusive table published
  uint32_t* records;       // [nstreams * nruns] records of rec_words words: local[tw] mask[mw] inclusive[tw]
  uint32_t rec_words;
  uint8_t* scratch;        // [warps of the grid][run] slots of `slot` bytes: a run's tiles until its offset is known
  uint32_t slot;
  };

__host__ __device__ constexpr uint32_t fpc_v0_tile_out_bytes(int wbytes, int group, int hdr)
  { return ((FPC_V0_TILE / group) * hdr + FPC_V0_TILE * wbytes + 63u) & ~15u; }      // + what warp_copy_smem_to_global reads behind the last byte

template <typename W>
__global__ void __launch_bounds__(FPC_V0_WARPS * 32)
fpc_encode_v0_tiles_kernel(const FpcV0TileArgs a)
  {
  using TR = FpcTraits<W>;
  static_assert(sizeof(W) == 4, "records are 32-bit words");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const unsigned gt = lanemask_gt();
  const uint32_t nt1 = 1u << a.e1, nt2 = 1u << a.e2, tw = nt1 + nt2, mw = (tw + 31u) >> 5;
  const uint32_t tw_pad = (tw + 3u) & ~3u, mw_pad = (mw + 3u) & ~3u;
  constexpr uint32_t OUTB = fpc_v0_tile_out_bytes(TR::WBYTES, TR::GROUP, TR::HDR);
  const size_t per_warp = (size_t)(2u * tw_pad + 2u * mw_pad) * 4u + OUTB;
  uint8_t* my = smem_raw + (size_t)warp * per_warp;
  W* T = reinterpret_cast<W*>(my);                         // state in front of the tile, then the encoder's working tables
  W* LT = T + tw_pad;                                      // last writer per context inside the tile
  uint32_t* M = reinterpret_cast<uint32_t*>(LT + tw_pad);  // contexts written inside the tile
  uint32_t* AM = M + mw_pad;                               // contexts filled during the look-back
  uint8_t* ob = reinterpret_cast<uint8_t*>(AM + mw_pad);
  const uint32_t total_tiles = a.nruns * (uint32_t)a.nstreams;       // tickets: runs
  const int h = a.e2 >> 1;
  const uint32_t lowmask = (1u << h) - 1u;

  for (;;)
    {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1u);
    t = __shfl_sync(FULL, t, 0);
    if (t >= total_tiles) break;
    const uint32_t c = t % (uint32_t)a.nstreams, k = t / (uint32_t)a.nstreams;    // run k of stream c: its predecessors hold lower tickets
    const uint32_t j0 = k * a.run * FPC_V0_TILE;
    const uint32_t cnt = a.n - j0 < a.run * FPC_V0_TILE ? a.n - j0 : a.run * FPC_V0_TILE;
    const W* src = reinterpret_cast<const W*>(a.in) + c;
    // the chain's registers in front of the run (fpc.c:104-113: everything starts at zero)
    const W pv1 = j0 >= 1u ? src[(size_t)(j0 - 1u) * a.stride] : (W)0;
    const W pv2 = j0 >= 2u ? src[(size_t)(j0 - 2u) * a.stride] : (W)0;
    const W pv3 = j0 >= 3u ? src[(size_t)(j0 - 3u) * a.stride] : (W)0;
    const uint32_t ta0 = (uint32_t)((W)(pv1 - pv2) >> (TR::BITS - a.e2)), tb0 = (uint32_t)((W)(pv2 - pv3) >> (TR::BITS - a.e2));
    const uint32_t idx = k * (uint32_t)a.nstreams + c;
    uint32_t* rec = a.records + (size_t)idx * a.rec_words;

    // ---- 1. local table ----
    for (uint32_t i = lane; i < mw; i += 32) { M[i] = 0; AM[i] = 0; }
    for (uint32_t i = lane; i < tw; i += 32) T[i] = 0;
    __syncwarp();
      {
      W carry_v = pv1;
      uint32_t carry_ta = ta0, carry_tb = tb0;
      W vnext = lane < cnt ? src[(size_t)(j0 + lane) * a.stride] : (W)0;       // two windows ahead: the loads are strided and far
      W vnext2 = lane + 32u < cnt ? src[(size_t)(j0 + lane + 32u) * a.stride] : (W)0;
      for (uint32_t i0 = 0; i0 < cnt; i0 += 32)
        {
        const uint32_t j = i0 + lane;
        const bool act = j < cnt;
        const W v = vnext;
        vnext = vnext2;
        vnext2 = j + 64u < cnt ? src[(size_t)(j0 + j + 64u) * a.stride] : (W)0;
        W vprev = __shfl_up_sync(FULL, v, 1);
        if (lane == 0) vprev = carry_v;
        const uint32_t c1 = (uint32_t)(vprev >> (TR::BITS - a.e1));
        const unsigned m1 = __match_any_sync(FULL, c1);
        const W s = v - vprev;
        const uint32_t tt = (uint32_t)(s >> (TR::BITS - a.e2));
        uint32_t ta = __shfl_up_sync(FULL, tt, 1);
        uint32_t tb = __shfl_up_sync(FULL, tt, 2);
        if (lane == 0) { ta = carry_ta; tb = carry_tb; }
        if (lane == 1) { tb = carry_ta; }
        const uint32_t c2 = nt1 + (((tb & lowmask) << h) ^ ta);
        const unsigned m2 = __match_any_sync(FULL, c2);
        const unsigned actmask = __ballot_sync(FULL, act);
        // the last active element of every context wins (what a serial pass leaves)
        if (act && ((m1 & gt & actmask) == 0)) { LT[c1] = v; atomicOr(&M[c1 >> 5], 1u << (c1 & 31u)); }
        if (act && ((m2 & gt & actmask) == 0)) { LT[c2] = s; atomicOr(&M[c2 >> 5], 1u << (c2 & 31u)); }
        carry_v = __shfl_sync(FULL, v, 31);
        carry_tb = __shfl_sync(FULL, tt, 30);
        carry_ta = __shfl_sync(FULL, tt, 31);
        __syncwarp();
        }
      }
    // publish it
    for (uint32_t q = lane; q < tw; q += 32) if ((M[q >> 5] >> (q & 31u)) & 1u) __stcg(rec + q, (uint32_t)LT[q]);
    for (uint32_t i = lane; i < mw; i += 32) __stcg(rec + tw + i, M[i]);
    __threadfence();
    __syncwarp();
    if (lane == 0 && k != 0u) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a.state + idx), "r"(1u) : "memory");

    // ---- 2. the state in front of the tile ----
    for (int64_t p = (int64_t)k - 1; p >= 0; --p)
      {
      const uint32_t pidx = (uint32_t)p * (uint32_t)a.nstreams + c;
      uint32_t st = 0;
      if (lane == 0)
        for (;;)
          {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(st) : "l"(a.state + pidx) : "memory");
          if (st != 0u) break;
          __nanosleep(40);
          }
      st = __shfl_sync(FULL, st, 0);
      const uint32_t* prec = a.records + (size_t)pidx * a.rec_words;
      if (st == 2u)
        { // the full state behind tile p: whatever is still open comes from it
        const uint32_t* inc = prec + tw + mw;
        for (uint32_t q = lane; q < tw; q += 32) if (!((AM[q >> 5] >> (q & 31u)) & 1u)) T[q] = (W)__ldcg(inc + q);
        break;
        }
      // tile p's own writes: the contexts not seen yet.  Lane i owns mask word i (tables of up to
      // 1056 entries: 33 words, the last one handled by every lane alike), so the words arrive in one
      // round trip and the value loads below do not wait for one another.
      const uint32_t have_l = lane < mw ? AM[lane] : 0u, pm_l = lane < mw ? __ldcg(prec + tw + lane) : 0u;
      uint32_t have_x = 0, pm_x = 0;
      if (mw > 32u) { have_x = AM[32]; pm_x = __ldcg(prec + tw + 32u); }
      const uint32_t new_l = pm_l & ~have_l, new_x = pm_x & ~have_x;
#pragma unroll 4
      for (uint32_t i = 0; i < mw; ++i)
        {
        const uint32_t nw = i < 32u ? __shfl_sync(FULL, new_l, (int)i) : new_x;
        if (nw == 0u) continue;
        const uint32_t q = 32u * i + lane;
        if ((nw >> lane) & 1u) T[q] = (W)__ldcg(prec + q);
        }
      __syncwarp();
      if (lane < mw) AM[lane] = have_l | pm_l;
      if (mw > 32u && lane == 0) AM[32] = have_x | pm_x;
      __syncwarp();
      }
    __syncwarp();
    // the state behind this tile, for its successors
      {
      uint32_t* inc = rec + tw + mw;
      for (uint32_t q = lane; q < tw; q += 32) __stcg(inc + q, (uint32_t)(((M[q >> 5] >> (q & 31u)) & 1u) ? LT[q] : T[q]));
      __threadfence();
      __syncwarp();
      if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a.state + idx), "r"(2u) : "memory");
      }

    // ---- 3. encode the run's tiles: bytes to the warp's scratch slots, byte counts published at once ----
    // (Placing every tile as soon as it is encoded would chain the runs: the first tile of a run would
    // wait for the LAST tile of the run before it.  The counts of a run are all out when its encoding
    // ends, which is about when its successor's ends.)
    uint8_t* out = a.out + (size_t)c * a.out_stride;
    uint8_t* scr = a.scratch + (size_t)(blockIdx.x * FPC_V0_WARPS + warp) * a.run * a.slot;
    uint64_t* desc = a.desc + (size_t)c * a.ntiles;
    const uint32_t tile0 = k * a.run;
    uint32_t nb0 = 0;
    for (uint32_t i0 = 0, i = 0; i0 < cnt; i0 += FPC_V0_TILE, ++i)
      {
      const uint32_t jt = j0 + i0;
      const uint32_t tc = cnt - i0 < FPC_V0_TILE ? cnt - i0 : FPC_V0_TILE;
      const W q1 = jt >= 1u ? src[(size_t)(jt - 1u) * a.stride] : (W)0;
      const W q2 = jt >= 2u ? src[(size_t)(jt - 2u) * a.stride] : (W)0;
      const W q3 = jt >= 3u ? src[(size_t)(jt - 3u) * a.stride] : (W)0;
      const uint32_t nb = fpc_encode_warp<W, true, const W*, true>(src + (size_t)jt * a.stride, a.stride, tc, ob, T, T + nt1, a.e1, a.e2, q1,
                                                                   (uint32_t)((W)(q1 - q2) >> (TR::BITS - a.e2)), (uint32_t)((W)(q2 - q3) >> (TR::BITS - a.e2)));
      __syncwarp();
      lookback_publish(desc, tile0 + i, nb);
      if (i == 0u) nb0 = nb;
      uint4* sv = reinterpret_cast<uint4*>(scr + (size_t)i * a.slot);
      for (uint32_t v4 = lane; v4 < (nb + 15u) >> 4; v4 += 32) __stcg(sv + v4, *reinterpret_cast<const uint4*>(ob + 16u * v4));
      __syncwarp();
      }
    // ---- 4. the run's offset, then its tiles to their places ----
    uint64_t at = lookback_walk(desc, tile0, nb0);
    if (tile0 == 0u && lane == 0)
      {
      out[0] = (uint8_t)(((a.e1 >> 1) << 4) | (a.e2 >> 1));                 // fpc.c:120
      out[1] = (uint8_t)(a.n >> 24); out[2] = (uint8_t)(a.n >> 16);          // fpc.c:123-126
      out[3] = (uint8_t)(a.n >> 8);  out[4] = (uint8_t)a.n;
      }
    for (uint32_t i0 = 0, i = 0; i0 < cnt; i0 += FPC_V0_TILE, ++i)
      {
      const uint32_t nb = i == 0u ? nb0 : (uint32_t)(lb_load(desc + tile0 + i) & LB_VAL);     // this warp published it
      if (i != 0u && lane == 0) lb_store(desc + tile0 + i, LB_INC | (at + nb));               // successors stop here
      warp_copy_global(out + 5 + at, scr + (size_t)i * a.slot, nb);
      at += nb;
      }
    if (tile0 + (cnt + FPC_V0_TILE - 1u) / FPC_V0_TILE == a.ntiles && lane == 0) a.nbytes[c] = (uint32_t)(at + 5u);
    __syncwarp();
    }
  }

// ---------------------------------------------------------------------------------------------
// Lane-serial decoder core.  `Fetch` returns the next `nb` residual bytes as a big-endian value.
// ---------------------------------------------------------------------------------------------
template <typename W> struct FpcLaneState
  {
  W pred1, pred2, last;
  uint32_t c1, c2;
  };

// One value: select the predictor by code, rebuild v, update both tables exactly like the
// reference decoder (fpc.c:308-326 / :977-995).  TSTRIDE = distance between table entries
// (32 for lane-interleaved shared-memory tables, 1 for plain arrays).
template <typename W, int TSTRIDE>
__device__ __forceinline__ W fpc_decode_value(FpcLaneState<W>& st, W x, bool use2, W* T1, W* T2, int e1, int e2, uint32_t m2)
  {
  using TR = FpcTraits<W>;
  const W v = x ^ (use2 ? st.pred2 : st.pred1);
  T1[(size_t)st.c1 * TSTRIDE] = v;
  st.c1 = (uint32_t)(v >> (TR::BITS - e1));
  st.pred1 = T1[(size_t)st.c1 * TSTRIDE];
  const W s = v - st.last;
  T2[(size_t)st.c2 * TSTRIDE] = s;
  st.c2 = ((st.c2 << (e2 >> 1)) ^ (uint32_t)(s >> (TR::BITS - e2))) & m2;
  st.pred2 = v + T2[(size_t)st.c2 * TSTRIDE];
  st.last = v;
  return v;
  }

// ---------------------------------------------------------------------------------------------
// K4: chunked decode fused with the SoA -> AoS transpose.
// CTA = NCOMP * R warps.  Warp w owns component c = w % NCOMP of 32 consecutive chunk ranges
// (lane = range), so a CTA owns 32*R ranges x NCOMP components and its output region is one
// contiguous slab of the AoS array.  Per sub-block of SB values:
//   a. the warp refreshes every lane's byte window from global memory with coalesced loads
//   b. every lane decodes SB values of its own chunk out of its window
//   c. values go to a [range][SB*NCOMP (+1 pad)] shared tile, then out as contiguous rows.
// ---------------------------------------------------------------------------------------------
struct FpcDecodeArgs
  {
  const uint8_t* sizes;    // u16 LE [nranges * ncomp]
  const uint8_t* payload;
  uint64_t payload_bytes;
  uint64_t n;
  uint32_t nranges;
  uint32_t ntiles;
  int log2S, e1, e2;
  void* out;               // device AoS
  uint64_t* desc;          // look-back descriptors, zeroed
  uint32_t* ticket;        // zeroed
  };

#ifndef TB200_FPC_DEC_WORDBUF
#define TB200_FPC_DEC_WORDBUF 1
#endif

template <typename W, int SB> struct FpcWindow
  {
  using TR = FpcTraits<W>;
  // A lane's window starts at the 16-byte boundary at or below its read position and must hold
  // everything one sub-block can consume (code words + SB full residuals) plus the words the
  // unaligned big-endian fetch touches behind the last byte.
  static constexpr int BYTES = 15 + (SB / TR::GROUP) * TR::HDR + SB * TR::WBYTES + (TR::WBYTES == 4 ? 8 : 12);
  static constexpr int VECS = (BYTES + 15) / 16;
  };

// The encoder's window holds the carried tail of the previous sub-block (< 16 bytes) plus everything
// one sub-block can produce (code words + SB full residuals); nothing is read past it.
template <typename W, int SB> struct FpcEncWindow
  {
  using TR = FpcTraits<W>;
  static constexpr int BYTES = 15 + (SB / TR::GROUP) * TR::HDR + SB * TR::WBYTES;
  static constexpr int VECS = (BYTES + 15) / 16;
  };

// EXP = (e1 << 8) | e2 compiles the predictor exponents in (the archive default (2,4) runs this
// way: shifts and masks become immediates); EXP = 0 takes them from the arguments.
template <typename W, int NCOMP, int R, int SB, int EXP>
__global__ void __launch_bounds__(NCOMP * R * 32)
fpc_decode_kernel(const FpcDecodeArgs a)
  {
  const int e1 = EXP ? (EXP >> 8) : a.e1, e2 = EXP ? (EXP & 255) : a.e2;
  using TR = FpcTraits<W>;
  using WIN = FpcWindow<W, SB>;
  constexpr int NWARPS = NCOMP * R;
  constexpr int NTHREADS = NWARPS * 32;
  constexpr int WV = WIN::VECS;                           // 16-byte vectors per lane window
  constexpr int WPV = sizeof(W) / 4;                      // 32-bit words per value
  constexpr int ROWLEN = SB * NCOMP * WPV;                // words per staged row
  constexpr int ROWW = fpc_stage_row_words(SB, NCOMP, sizeof(W));   // odd row stride: lane = row, so bank = lane + word
  static_assert(NTHREADS % ROWLEN == 0 && NTHREADS / ROWLEN == R, "one thread per row word, R rows per pass");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t nt1 = 1u << e1, nt2 = 1u << e2;
  uint32_t* win = reinterpret_cast<uint32_t*>(smem_raw);                               // [NWARPS][32][WV*4]
  uint32_t* stagebuf = win + (size_t)NWARPS * 32 * WV * 4;                             // [32*R][ROWW]
  W* tables = reinterpret_cast<W*>(stagebuf + (size_t)32 * R * ROWW);                  // [NWARPS][nt1+nt2][32]
  __shared__ uint32_t sh_tile;
  __shared__ uint32_t sh_scan[NTHREADS];
  __shared__ uint32_t sh_wsum[NWARPS];
  __shared__ uint64_t sh_base;

  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  if (threadIdx.x == 0) sh_tile = atomicAdd(a.ticket, 1u);
  __syncthreads();
  const uint32_t tile = sh_tile;

  // chunk offsets: block-exclusive scan of the tile's sizes in stream order, tile base by look-back
  const uint64_t nchunks = (uint64_t)a.nranges * NCOMP;
  const uint64_t g0 = (uint64_t)tile * NTHREADS;
  uint32_t mysz = 0;
  if (g0 + threadIdx.x < nchunks)
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
  if (lane == 31) sh_wsum[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0, tsum = 0;
#pragma unroll
  for (int w = 0; w < NWARPS; ++w) { const uint32_t t = sh_wsum[w]; if (w < (int)warp) wbase += t; tsum += t; }
  sh_scan[threadIdx.x] = wbase + incl - mysz;
  if (warp == 0)
    {
    const uint64_t excl = lookback_exclusive(a.desc, tile, tsum);
    if (lane == 0) sh_base = excl;
    }
  __syncthreads();

  // this lane's chunk
  const uint32_t c = warp % NCOMP, rgrp = warp / NCOMP;
  const uint32_t klocal = rgrp * 32 + lane;
  const uint64_t k = (uint64_t)tile * (32 * R) + klocal;
  const uint32_t S = 1u << a.log2S;
  uint32_t cnt = 0;
  // Byte positions are kept relative to tb16, the 16-byte boundary at or below the tile's first
  // payload byte, so all per-lane arithmetic is 32-bit.  relA = position of the next unread byte.
  const uint8_t* tb = a.payload + sh_base;
  const uint32_t A = (uint32_t)reinterpret_cast<uintptr_t>(tb) & 15u;
  const uint8_t* tb16 = tb - A;
  uint32_t relA = A;
  if (k < a.nranges)
    {
    const uint64_t lo = k << a.log2S;
    cnt = (uint32_t)((a.n - lo < S) ? (a.n - lo) : S);
    relA = A + sh_scan[klocal * NCOMP + c];
    }
  // first byte past the payload, relative to tb16 (window copies never read at or beyond it)
  const uint64_t end64 = (uint64_t)(a.payload + a.payload_bytes - tb16);
  const uint32_t endoff = a.payload_bytes <= sh_base ? 0u : (end64 > 0x7fffffffull ? 0x7fffffffu : (uint32_t)end64);
  const uint32_t last16 = endoff ? ((endoff - 1u) & ~15u) : 0u;
  // number of sub-blocks = that of the fullest chunk in the CTA (range 0 of the tile is never shorter)
  const uint64_t lo0 = ((uint64_t)tile * (32 * R)) << a.log2S;
  const uint32_t cnt0 = (uint32_t)((a.n - lo0 < S) ? (a.n - lo0) : S);
  // every range of the tile is complete (S values) unless the tile holds the stream's last range
  const bool full_tile = ((uint64_t)tile + 1) * (32 * R) < a.nranges;

  W* T1 = tables + (size_t)warp * (nt1 + nt2) * 32 + lane;
  W* T2 = T1 + (size_t)nt1 * 32;
  for (uint32_t i = 0; i < nt1 + nt2; ++i) T1[(size_t)i * 32] = 0;
  FpcLaneState<W> st; st.pred1 = 0; st.pred2 = 0; st.last = 0; st.c1 = 0; st.c2 = 0;
  const uint32_t m2 = nt2 - 1;

  const uint32_t* wrow = win + ((size_t)warp * 32 + lane) * WV * 4;
  const uint32_t wwarp_s = (uint32_t)__cvta_generic_to_shared(win + (size_t)warp * 32 * WV * 4);
  uint32_t* srow = stagebuf + (size_t)klocal * ROWW + c * WPV;
  uint8_t* gout = reinterpret_cast<uint8_t*>(a.out);
  const bool out_aligned = (reinterpret_cast<uintptr_t>(gout) & 3u) == 0;

  // Window refill: the warp's 32 windows are contiguous in shared memory ([lane][WV] vectors), so
  // copy number ci lands at vector ci; its source is vector ci % WV of lane ci / WV's window.
  // 16-byte cp.async copies (global -> shared without a register round trip), all in flight at once;
  // bytes at or beyond the end of the payload are not read (src-size operand, zero filled).
  uint32_t fsrc[WV];                                       // copy `it` of this lane: source lane | 16 * vector << 8
#pragma unroll
  for (int it = 0; it < WV; ++it)
    {
    const uint32_t ci = (uint32_t)it * 32u + lane;
    const uint32_t L = ci / (uint32_t)WV;
    fsrc[it] = L | ((16u * (ci - L * (uint32_t)WV)) << 8);
    }
  auto fill = [&]()
    {
    const uint32_t arel = relA & ~15u;
    if (__all_sync(FULL, arel + 16u * WV <= endoff))
      { // every window of the warp ends inside the payload (all but the stream's last tiles)
#pragma unroll
      for (int it = 0; it < WV; ++it)
        {
        const uint32_t off = __shfl_sync(FULL, arel, fsrc[it] & 31u) + (fsrc[it] >> 8);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(wwarp_s + 16u * ((uint32_t)it * 32u + lane)), "l"(tb16 + off) : "memory");
        }
      }
    else
      {
#pragma unroll
      for (int it = 0; it < WV; ++it)
        {
        const uint32_t ci = (uint32_t)it * 32u + lane;
        const uint32_t off = __shfl_sync(FULL, arel, fsrc[it] & 31u) + (fsrc[it] >> 8);
        const int32_t rem = (int32_t)endoff - (int32_t)off;
        const uint32_t ssz = rem >= 16 ? 16u : (rem > 0 ? (uint32_t)rem : 0u);
        const uint32_t offc = off < last16 ? off : last16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(wwarp_s + 16u * ci), "l"(tb16 + offc), "r"(ssz) : "memory");
        }
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
    };

  fill();
  for (uint32_t i0 = 0; i0 < cnt0; i0 += SB)
    {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // b. decode up to SB values of this lane's chunk out of its window
    const uint32_t bp0 = relA & 15u;
    uint32_t bp = bp0;
    const uint32_t todo = (i0 < cnt) ? ((cnt - i0 < (uint32_t)SB) ? cnt - i0 : (uint32_t)SB) : 0;
    // 32-bit values keep their read position in BITS (pos8, relative to the window start): the
    // shift counts of the extraction are then plain register operands.
    uint32_t pos8 = 8u * bp0;
#if TB200_FPC_DEC_WORDBUF
    // The lane walks its window through a three-word big-endian register buffer (q0 = word holding
    // the next unread byte, q1, q2 = the two behind it; pb8 = bits of q0 already consumed).  A fetch
    // is two funnel shifts on q0:q1; when it crosses into q1 the buffer advances and ONE word is
    // loaded - into q2, i.e. a whole word before it can be needed - so a value costs about one
    // shared-memory read instead of the two of an unaligned fetch, off the dependency chain.
    uint32_t q0 = 0, q1 = 0, q2 = 0, pb8 = pos8 & 31u;
    const uint32_t* qn = wrow + (bp0 >> 2) + 3;
    if (sizeof(W) == 4)
      {
      q0 = __byte_perm(qn[-3], 0u, 0x0123u); q1 = __byte_perm(qn[-2], 0u, 0x0123u); q2 = __byte_perm(qn[-1], 0u, 0x0123u);
      }
#endif
    auto take = [&](uint32_t nb8) -> uint32_t
      { // the next nb8 / 8 (0..4) bytes as a big-endian number
#if TB200_FPC_DEC_WORDBUF
      const uint32_t t = __funnelshift_l(q1, q0, pb8);
      const uint32_t x = __funnelshift_lc(t, 0u, nb8);
      pb8 += nb8;
      if (pb8 >= 32u)
        {
        q0 = q1; q1 = q2;
        q2 = __byte_perm(*qn++, 0u, 0x0123u);
        }
      pb8 &= 31u;
      return x;
#else
      const uint32_t* w = wrow + (pos8 >> 5);
      const uint32_t t = __funnelshift_r(w[0], w[1], pos8);             // shift count taken modulo 32
      pos8 += nb8;
      return __funnelshift_lc(__byte_perm(t, 0u, 0x0123u), 0u, nb8);
#endif
      };
    auto decode_group32 = [&](uint32_t g, auto checked)
      { // fpc.c:248-326.  The eight 3-bit codes are split once per group: bit 3j of `two` = value j
        // takes the DFCM prediction (code > 4), and the residual lengths in bits (8 * (code, or
        // code - 4): clearing bit 2 of codes 5..7) sit in bits 3j+3..3j+5 of `len8`.
      constexpr bool CHECK = decltype(checked)::value;
      const uint32_t bc = take(24u);
      const uint32_t two = (bc >> 2) & ((bc >> 1) | bc) & 0x249249u;
      const uint32_t len8 = (bc & ~(two << 2)) << 3;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        {
        const uint32_t nb8 = (len8 >> (3 * jj)) & 0x38u;
        const bool use2 = (two >> (3 * jj)) & 1u;
        const W x = (W)take(nb8);
        const uint32_t idx = g * 8u + jj;
        if (!CHECK || idx < todo)
          {
          const W v = fpc_decode_value<W, 32>(st, x, use2, T1, T2, e1, e2, m2);
          srow[idx * NCOMP * WPV] = (uint32_t)v;
          }
        }
      };
    auto decode_group64 = [&](uint32_t g, auto checked)
      {
      constexpr bool CHECK = decltype(checked)::value;
      // code word
      uint32_t bc;
        {
        const uint32_t wi = bp >> 2;
        const uint32_t be = __byte_perm(wrow[wi], wrow[wi + 1], 0x0123u + (bp & 3u) * 0x1111u);
        bc = be >> (32 - 8 * TR::HDR);
        bp += TR::HDR;
        }
#pragma unroll
      for (int jj = 0; jj < TR::GROUP; ++jj)
        {
        const uint32_t code = (bc >> (TR::CBITS * jj)) & ((1u << TR::CBITS) - 1u);
        const bool use2 = code > (uint32_t)TR::BASE2;
        const uint32_t nb = use2 ? code - TR::BASE2 : code;
        const uint32_t wi = bp >> 2;
        const uint32_t sel = 0x0123u + (bp & 3u) * 0x1111u;
        const uint32_t w0 = wrow[wi], w1 = wrow[wi + 1], w2 = wrow[wi + 2];
        const uint64_t be = ((uint64_t)__byte_perm(w0, w1, sel) << 32) | __byte_perm(w1, w2, sel);
        const W x = nb ? (W)(be >> (64 - 8 * nb)) : (W)0;
        bp += nb;
        const uint32_t idx = g * TR::GROUP + jj;
        if (!CHECK || idx < todo)
          {
          const W v = fpc_decode_value<W, 32>(st, x, use2, T1, T2, e1, e2, m2);
          srow[idx * NCOMP * WPV] = (uint32_t)v;
          srow[idx * NCOMP * WPV + 1] = (uint32_t)((uint64_t)v >> 32);
          }
        }
      };
    auto decode_group = [&](uint32_t g, auto checked)
      {
      if (sizeof(W) == 4) decode_group32(g, checked); else decode_group64(g, checked);
      };
    if (__all_sync(FULL, todo == (uint32_t)SB))
      { // every lane has a full sub-block: no per-value bounds checks
#pragma unroll 1
      for (uint32_t g = 0; g < (uint32_t)SB / TR::GROUP; ++g) decode_group(g, std::false_type{});
      }
    else
      {
#pragma unroll 1
      for (uint32_t g = 0; g < (uint32_t)SB / TR::GROUP; ++g)
        {
        if (g * TR::GROUP >= todo) break;
        decode_group(g, std::true_type{});
        }
      }
    if (sizeof(W) == 4)
      {
#if TB200_FPC_DEC_WORDBUF
      bp = 4u * (uint32_t)(qn - wrow - 3) + (pb8 >> 3);
#else
      bp = pos8 >> 3;
#endif
      }
    relA += bp - bp0;
    // the next window crosses L2 -> shared memory while the slab is flushed
    if (i0 + SB < cnt0) fill();
    __syncthreads();

    // c. flush the staged slab: row r = range (tile*32R + r), values [i0, i0+SB) x NCOMP, contiguous
    if (full_tile && out_aligned)
      { // thread t moves word t % ROWLEN of rows t / ROWLEN, + R, + 2R, ...: consecutive lanes read
        // consecutive shared-memory words (no bank conflicts: the row stride is odd) and write
        // consecutive global words (whole 128-byte lines per warp store)
      const uint32_t w = threadIdx.x % (uint32_t)ROWLEN, r0 = threadIdx.x / (uint32_t)ROWLEN;
      const uint32_t row_bytes = (uint32_t)(NCOMP * sizeof(W)) << a.log2S;
      uint8_t* to = gout + ((((uint64_t)tile * (32 * R)) << a.log2S) + i0) * (NCOMP * sizeof(W)) + (size_t)r0 * row_bytes + 4u * w;
      const uint32_t* from = stagebuf + (size_t)r0 * ROWW + w;
      const size_t step = (size_t)R * row_bytes;                      // the pointer walks from row to row: one 64-bit add per store
      constexpr int UN = 8;
#pragma unroll 1
      for (int it0 = 0; it0 < 32; it0 += UN)
        {
        uint32_t v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) v[u] = from[(size_t)(it0 + u) * R * ROWW];
#pragma unroll
        for (int u = 0; u < UN; ++u) { __stcs(reinterpret_cast<uint32_t*>(to), v[u]); to += step; }
        }
      }
    else
      {
      for (uint32_t r = warp; r < 32u * R; r += NWARPS)
        {
        const uint64_t kr = (uint64_t)tile * (32 * R) + r;
        if (kr >= a.nranges) break;
        const uint64_t lo = kr << a.log2S;
        const uint32_t rc = (uint32_t)((a.n - lo < S) ? (a.n - lo) : S);
        if (i0 >= rc) continue;
        const uint32_t nw = ((rc - i0 < (uint32_t)SB) ? rc - i0 : (uint32_t)SB) * NCOMP * WPV;
        const uint32_t* sr = stagebuf + (size_t)r * ROWW;
        uint32_t* go = reinterpret_cast<uint32_t*>(gout + (lo + i0) * (NCOMP * sizeof(W)));
        for (uint32_t q = lane; q < nw; q += 32) go[q] = sr[q];
        }
      }
    __syncthreads();
    }
  }

// ---------------------------------------------------------------------------------------------
// K3: chunked encode, LANE PER CHUNK - the mirror image of K4.
//
// The warp-cooperative routine above (32 consecutive values per warp step) spends most of its
// ~180 instructions per step on cross-lane bookkeeping (match.any, shuffles, ballots, scattered byte
// stores).  Chunks are independent, so the cheaper way to use a warp is the decoder's: every lane
// runs the plain serial predictor of the reference on its own chunk (tables lane-interleaved in
// shared memory), and the memory system is served by the warp as a whole:
//   a. the CTA stages a slab of SB values of all its ranges with 16-byte cp.async copies
//      (this is trico_transpose_*_aos_to_soa, transpose_aos_to_soa.c:8-82, fused)
//   b. every lane encodes SB values of its chunk into a byte window in shared memory (32-bit
//      accumulator, one word store per four output bytes)
//   c. the warp writes the completed 16-byte vectors of its 32 windows to the chunks' scratch
//      slots in global memory (the CTA reuses its slots tile after tile, so they live in L2)
//   d. at the end of the tile: chunk sizes -> block scan -> decoupled look-back -> every chunk is
//      copied once from its slot to its final offset; the u16 size table and the stream's
//      payload_bytes field are written by the same kernel.
// CTA = NCOMP * R warps; warp w owns component w % NCOMP of 32 consecutive ranges (lane = range).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fpc_emit(uint32_t* wrow, uint32_t& acc, uint32_t& fill, uint32_t& wp, uint32_t le, uint32_t n)
  { // appends n (0..4) bytes; `le` holds them in memory order in its low bytes, upper bytes zero
  const uint32_t lo = acc | (le << (8u * fill));
  const uint32_t nf = fill + n;
  if (nf >= 4u) { wrow[wp++] = lo; acc = __funnelshift_rc(le, 0u, 32u - 8u * fill); fill = nf - 4u; }
  else { acc = lo; fill = nf; }
  }

// warp copy of n bytes between global buffers; src 16-byte aligned (read around L1), dst arbitrary
__device__ __forceinline__ void warp_copy_global(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n)
  {
  const unsigned lane = lane_id();
  uint32_t head = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
  if (head > n) head = n;
  if (lane < head) dst[lane] = __ldcg(src + lane);
  const uint32_t nvec = (n - head) >> 4;
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(src + (head & ~3u));
  const unsigned sh = (head & 3u) * 8u;
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
  constexpr int UN = 2;
  for (uint32_t i0 = 0; i0 < nvec; i0 += 32 * UN)
    {
    uint32_t w[UN][5];
#pragma unroll
    for (int u = 0; u < UN; ++u)
      {
      const uint32_t i = i0 + lane + 32 * u;
      if (i < nvec)
        {
        const uint32_t* q = sw + 4 * i;
#pragma unroll
        for (int j = 0; j < 5; ++j) w[u][j] = __ldcg(q + j);
        }
      }
#pragma unroll
    for (int u = 0; u < UN; ++u)
      {
      const uint32_t i = i0 + lane + 32 * u;
      if (i < nvec)
        dv[i] = make_uint4(__funnelshift_r(w[u][0], w[u][1], sh), __funnelshift_r(w[u][1], w[u][2], sh),
                           __funnelshift_r(w[u][2], w[u][3], sh), __funnelshift_r(w[u][3], w[u][4], sh));
      }
    }
  const uint32_t done = head + (nvec << 4);
  if (done + lane < n) dst[done + lane] = __ldcg(src + done + lane);
  }

#ifndef TB200_FPC_ENC_PREFETCH
#define TB200_FPC_ENC_PREFETCH 1
#endif
#ifndef TB200_FPC_ENC_TMA
#define TB200_FPC_ENC_TMA 1         // the slab arrives as one bulk copy (TMA, 1-D) per row, completion on an mbarrier
#endif
#ifndef TB200_FPC_ENC_WARPS
#define TB200_FPC_ENC_WARPS 15      // resident warps per SM the register allocation aims at
#endif
// DEFER: the tile's chunks are not copied out at its end but in small pieces during the NEXT tile,
// two chunks per warp and sub-block through a per-warp bounce buffer filled by cp.async while the
// lanes encode (needs two sets of scratch slots per CTA).  Neither the look-back of the tile - its
// predecessors finish at about the same time, so it used to wait for the slowest of the wave - nor
// the read-back of the slots is waited for any more.
// mbarrier + 1-D bulk copy (TMA) helpers of the encoder's slab staging: one `cp.async.bulk` per staged
// row replaces 24 16-byte cp.async copies (and their address arithmetic in every thread); the bytes
// are counted on an mbarrier every thread polls for itself, so the CTA barrier in front of the encode
// goes as well.
__device__ __forceinline__ void mbar_init(uint32_t mbar_s, uint32_t count)
  {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar_s), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar_s, uint32_t bytes)
  {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_s), "r"(bytes) : "memory");
  }
__device__ __forceinline__ void mbar_wait(uint32_t mbar_s, uint32_t parity)
  {
  uint32_t done;
  do
    {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mbar_s), "r"(parity) : "memory");
    } while (!done);
  }
__device__ __forceinline__ void bulk_g2s(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t mbar_s, uint64_t policy)
  {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               :: "r"(dst_s), "l"(src), "r"(bytes), "r"(mbar_s), "l"(policy) : "memory");
  }

template <typename W, int NCOMP, int R, int SB, int EXP, bool DEFER>
#ifdef TB200_FPC_ENC_MAXNREG
__global__ void __maxnreg__(TB200_FPC_ENC_MAXNREG)
#else
__global__ void __launch_bounds__(NCOMP * R * 32, TB200_FPC_ENC_WARPS / (NCOMP * R))
#endif
fpc_encode_lanes_kernel(const FpcEncodeArgs a)
  {
  const int e1 = EXP ? (EXP >> 8) : a.e1, e2 = EXP ? (EXP & 255) : a.e2;
  using TR = FpcTraits<W>;
  using WIN = FpcEncWindow<W, SB>;
  constexpr int NWARPS = NCOMP * R;
  constexpr int NTHREADS = NWARPS * 32;                   // == chunks per tile
  constexpr int WV = WIN::VECS;                           // 16-byte vectors per lane window
  constexpr int WPV = sizeof(W) / 4;
  constexpr int ROWV = SB * NCOMP * WPV / 4;              // 16-byte vectors per staged row
  constexpr int ROWW = fpc_stage_row_words_enc(SB, NCOMP, sizeof(W));   // rows stay 16-byte aligned (16-byte cp.async)
  static_assert((SB * NCOMP * WPV) % 4 == 0, "staged rows must be whole vectors");
  static_assert(SB % TR::GROUP == 0, "sub-blocks hold whole groups");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t nt1 = 1u << e1, nt2 = 1u << e2;
  uint32_t* win = reinterpret_cast<uint32_t*>(smem_raw);                               // [NWARPS][32][WV*4]
  uint32_t* stagebuf = win + (size_t)NWARPS * 32 * WV * 4;                             // [32*R][ROWW]
  W* tables = reinterpret_cast<W*>(stagebuf + (size_t)32 * R * ROWW);                  // [NWARPS][nt1+nt2][32]
  __shared__ uint32_t sh_tile;
  __shared__ uint32_t sh_size[NTHREADS];                  // stream order: index = range_local * NCOMP + component
  __shared__ uint32_t sh_off[NTHREADS];
  __shared__ uint32_t sh_wsum[NWARPS];
  __shared__ uint64_t sh_base;
  // DEFER: sizes, offsets, total and base of the tile that is being drained
  __shared__ uint32_t sh_size_p[DEFER ? NTHREADS : 1];
  __shared__ uint32_t sh_off_p[DEFER ? NTHREADS : 1];
  __shared__ uint32_t sh_tsum_p;
  __shared__ uint64_t sh_base_p;
  __shared__ __align__(8) uint64_t sh_mbar;               // counts the bytes of the slab in flight (TMA staging)

  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t c = warp % NCOMP, rgrp = warp / NCOMP;
  const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(&sh_mbar);
  uint32_t slab_phase = 0;                                // parity of the slab the CTA waits for next (uniform)
  if (TB200_FPC_ENC_TMA && threadIdx.x == 0) mbar_init(mbar_s, 1);   // the tile loop's first barrier publishes it
  const uint32_t klocal = rgrp * 32 + lane;
  const uint32_t S = 1u << a.log2S;
  const uint32_t h2 = (uint32_t)e2 >> 1, m2 = nt2 - 1;
  W* T1 = tables + (size_t)warp * (nt1 + nt2) * 32 + lane;
  W* T2 = T1 + (size_t)nt1 * 32;
  uint32_t* wrow = win + ((size_t)warp * 32 + lane) * WV * 4;
  const uint4* wwarp4 = reinterpret_cast<const uint4*>(win + (size_t)warp * 32 * WV * 4);
  const uint32_t* srow = stagebuf + (size_t)klocal * ROWW + c * WPV;
  const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stagebuf);
  const size_t set_bytes = (size_t)NTHREADS * a.slot;                                  // one set of scratch slots
  uint8_t* const scr0 = a.scratch + (size_t)blockIdx.x * (DEFER ? 2 : 1) * set_bytes;
  uint8_t* cta_scr = scr0;
  uint8_t* my_scr = cta_scr + (size_t)(klocal * NCOMP + c) * a.slot;
  // DEFER state (uniform over the CTA unless noted)
  uint8_t* bounce_d = reinterpret_cast<uint8_t*>(tables + (size_t)NWARPS * (nt1 + nt2) * 32) + (size_t)warp * a.slot;   // per warp
  const uint32_t bounce_s = (uint32_t)__cvta_generic_to_shared(bounce_d);
  bool pv_valid = false, pv_have_base = false, pv_loaded = false;                       // pv_loaded: per warp
  uint32_t pv_tile = 0, pv_set = 0, cur_set = 0, pv_next = 0, pv_cur = 0;              // pv_next, pv_cur: per warp
  const uint8_t* gin = reinterpret_cast<const uint8_t*>(a.in);
  const bool in_aligned = (reinterpret_cast<uintptr_t>(gin) & 15u) == 0;
  // L2 residency: the input is read once (evict first); the scratch slots are written now and read
  // back at the end of the tile (evict last when written, evict first when read back)
  uint64_t pol_first, pol_last;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  auto st_scratch = [&](uint8_t* p, const uint4 v)
    { asm volatile("st.global.cg.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol_last) : "memory"); };
  auto ld_scratch = [&](const uint4* p)
    {
    uint4 w;
    asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p), "l"(pol_first) : "memory");
    return w;
    };
  // DEFER: the bounce buffer's chunk -> its final place in the payload
  auto pv_write = [&]()
    {
    const uint32_t nbytes = sh_size_p[pv_cur];
    if (nbytes) warp_copy_smem_to_global(a.payload + sh_base_p + sh_off_p[pv_cur], bounce_d, nbytes);
    __syncwarp();
    pv_loaded = false;
    };
  // DEFER: the warp's next chunk of the previous tile, scratch slot -> bounce buffer (asynchronously)
  auto pv_fetch = [&]()
    {
    if (pv_next >= (uint32_t)NTHREADS) return;
    const uint32_t q = pv_next;
    const uint32_t nv = (sh_size_p[q] + 15u) >> 4;
    const uint8_t* sp = scr0 + pv_set * set_bytes + (size_t)q * a.slot;
    for (uint32_t i = lane; i < nv; i += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(bounce_s + 16u * i), "l"(sp + 16u * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    pv_cur = q; pv_next += NWARPS; pv_loaded = true;
    };
  // DEFER: base offset of the previous tile (warp 0; the others see it after the next barrier)
  auto pv_lookback = [&]()
    {
    if (warp != 0) return;
    const uint64_t excl = lookback_walk(a.desc, pv_tile, sh_tsum_p);
    if (lane == 0)
      {
      sh_base_p = excl;
      if (pv_tile == a.ntiles - 1)
        {
        *a.total = excl + sh_tsum_p;
        store_u64_bytes(a.total_field, excl + sh_tsum_p);
        }
      }
    };
  // DEFER: whatever is left of the previous tile, waiting for every copy (end of a tile / of the kernel)
  auto pv_finish = [&]()
    {
    if (!pv_valid) return;
    if (!pv_have_base) { pv_lookback(); __syncthreads(); pv_have_base = true; }
    while (pv_loaded || pv_next < (uint32_t)NTHREADS)
      {
      if (pv_loaded)
        {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        pv_write();
        }
      pv_fetch();
      }
    pv_valid = false;
    };

  for (;;)
    {
    __syncthreads();
    if (threadIdx.x == 0) sh_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t tile = sh_tile;
    if (tile >= a.ntiles) break;

    const uint64_t k = (uint64_t)tile * (32 * R) + klocal;
    uint32_t cnt = 0;
    if (k < a.nranges)
      {
      const uint64_t lo = k << a.log2S;
      cnt = (uint32_t)((a.n - lo < S) ? (a.n - lo) : S);
      }
    const uint64_t lo0 = ((uint64_t)tile * (32 * R)) << a.log2S;
    const uint32_t cnt0 = (uint32_t)((a.n - lo0 < S) ? (a.n - lo0) : S);
    const bool fast_in = ((uint64_t)tile + 1) * (32 * R) < a.nranges && in_aligned;   // every range of the tile is complete
    if (DEFER)
      {
      cta_scr = scr0 + cur_set * set_bytes;
      my_scr = cta_scr + (size_t)(klocal * NCOMP + c) * a.slot;
      }

    for (uint32_t i = 0; i < nt1 + nt2; ++i) T1[(size_t)i * 32] = 0;
    W pred1 = 0, pred2 = 0, last = 0;
    uint32_t c1 = 0, c2 = 0;
    uint32_t acc = 0, fill = 0, wp = 0, flushed = 0;

    // a. slab of SB values per range: row r of the staging tile = range tile*32R + r
    auto stage_in = [&](uint32_t i0)
      {
      if (fast_in)
        {
        const uint8_t* tin = gin + (lo0 + i0) * (NCOMP * sizeof(W));
        const uint32_t row_bytes = (uint32_t)(NCOMP * sizeof(W)) << a.log2S;
#if TB200_FPC_ENC_TMA
        if (warp == 0)
          { // lane = row: 16 * ROWV bytes each, all counted on the CTA's mbarrier
          if (lane == 0) mbar_expect_tx(mbar_s, 32u * R * ROWV * 16u);
          __syncwarp();
#pragma unroll
          for (uint32_t r = lane; r < 32u * R; r += 32)
            bulk_g2s(stage_s + r * (uint32_t)ROWW * 4u, tin + (size_t)r * row_bytes, (uint32_t)ROWV * 16u, mbar_s, pol_first);
          }
#else
        for (uint32_t ci = threadIdx.x; ci < 32u * R * ROWV; ci += NTHREADS)
          {
          const uint32_t r = ci / (uint32_t)ROWV, v = ci - r * (uint32_t)ROWV;
          asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(stage_s + (r * ROWW + 4u * v) * 4u), "l"(tin + (size_t)r * row_bytes + 16u * v), "l"(pol_first) : "memory");
          }
#endif
#if TB200_FPC_ENC_PREFETCH
        // the slab after this one -> L2 (one 128-byte line per thread), so that its copy, issued one
        // sub-block from now, finds the data on the chip instead of waiting for DRAM
        if (i0 + SB < S)
          {
          constexpr uint32_t LPR = ROWV / 8;                           // lines per row segment
          const uint32_t r = threadIdx.x / LPR, l = threadIdx.x - r * LPR;
          asm volatile("prefetch.global.L2 [%0];" :: "l"(tin + (size_t)r * row_bytes + (size_t)SB * (NCOMP * sizeof(W)) + 128u * l));
          }
#endif
        }
      else
        {
        for (uint32_t r = warp; r < 32u * R; r += NWARPS)
          {
          const uint64_t kr = (uint64_t)tile * (32 * R) + r;
          if (kr >= a.nranges) break;
          const uint64_t lo = kr << a.log2S;
          const uint32_t rc = (uint32_t)((a.n - lo < S) ? (a.n - lo) : S);
          if (i0 >= rc) continue;
          const uint32_t nw = ((rc - i0 < (uint32_t)SB) ? rc - i0 : (uint32_t)SB) * NCOMP * WPV;
          const uint32_t* gi = reinterpret_cast<const uint32_t*>(gin + (lo + i0) * (NCOMP * sizeof(W)));
          for (uint32_t q = lane; q < nw; q += 32) stagebuf[(size_t)r * ROWW + q] = gi[q];
          }
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
      };

    stage_in(0);
    for (uint32_t i0 = 0; i0 < cnt0; i0 += SB)
      {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (TB200_FPC_ENC_TMA && fast_in)
        { // every thread sees the slab's bytes arrive for itself: no CTA barrier in front of the encode
        mbar_wait(mbar_s, slab_phase);
        slab_phase ^= 1u;
        }
      else __syncthreads();
      if (DEFER && pv_valid)
        { // a piece of the previous tile: the chunk that arrived in the bounce buffer leaves, the next one is requested
        if (!pv_have_base && i0 >= (uint32_t)SB) pv_lookback();      // one sub-block after the tile's end its predecessors have published
        if (pv_have_base) { if (pv_loaded) pv_write(); pv_fetch(); }
        }

      // b. encode up to SB values of this lane's chunk
      const uint32_t todo = (i0 < cnt) ? ((cnt - i0 < (uint32_t)SB) ? cnt - i0 : (uint32_t)SB) : 0;
      auto encode_group = [&](uint32_t g, auto checked)
        {
        constexpr bool CHECK = decltype(checked)::value;
        W xs[TR::GROUP];
        uint32_t nbs[TR::GROUP];
        uint32_t bc = 0;
#pragma unroll
        for (int jj = 0; jj < TR::GROUP; ++jj)
          {
          const uint32_t idx = g * TR::GROUP + jj;
          uint32_t code = 1, nb = 1;                                // pad slot of the last group: code 1 + one zero byte (fpc.c:196-204, :789-794)
          W x = 0;
          if (!CHECK || idx < todo)
            {
            const W v = *reinterpret_cast<const W*>(srow + (size_t)idx * NCOMP * WPV);
            const W x1 = v ^ pred1, x2 = v ^ pred2;
            T1[(size_t)c1 * 32] = v;
            c1 = (uint32_t)(v >> (TR::BITS - e1));
            pred1 = T1[(size_t)c1 * 32];
            const W st = v - last;
            T2[(size_t)c2 * 32] = st;
            c2 = ((c2 << h2) ^ (uint32_t)(st >> (TR::BITS - e2))) & m2;
            pred2 = v + T2[(size_t)c2 * 32];
            last = v;
            // code selection, fpc.c:146-189 / :635-782
            const int n1 = sig_bytes(x1);
            int n2 = sig_bytes(x2); if (n2 == 0) n2 = 1;
            const bool use2 = (n1 >= 2) && (n2 < n1);
            code = use2 ? TR::BASE2 + n2 : n1;
            nb = use2 ? n2 : n1;
            x = use2 ? x2 : x1;
            }
          bc |= code << (TR::CBITS * jj);
          xs[jj] = x; nbs[jj] = nb;
          }
        // code word (big-endian), then the residual bytes, most significant first
        if (TR::HDR == 3) fpc_emit(wrow, acc, fill, wp, __byte_perm(bc, 0u, 0x4012u), 3u);
        else              fpc_emit(wrow, acc, fill, wp, bc, 1u);
#pragma unroll
        for (int jj = 0; jj < TR::GROUP; ++jj)
          {
          const uint32_t nb = nbs[jj];
          if (sizeof(W) == 4)
            {
            const uint32_t xl = __funnelshift_lc(0u, (uint32_t)xs[jj], 32u - 8u * nb);     // left-justified, 0 when nb == 0
            fpc_emit(wrow, acc, fill, wp, __byte_perm(xl, 0u, 0x0123u), nb);
            }
          else
            {
            const uint64_t xl = nb ? ((uint64_t)xs[jj] << (8u * (8u - nb))) : 0ull;
            const uint32_t first = __byte_perm((uint32_t)(xl >> 32), 0u, 0x0123u), second = __byte_perm((uint32_t)xl, 0u, 0x0123u);
            fpc_emit(wrow, acc, fill, wp, first, nb < 4u ? nb : 4u);
            fpc_emit(wrow, acc, fill, wp, second, nb > 4u ? nb - 4u : 0u);
            }
          }
        };
      if (__all_sync(FULL, todo == (uint32_t)SB))
        { // every lane has a full sub-block: no per-value bounds checks, no pad slots
#pragma unroll 1
        for (uint32_t g = 0; g < (uint32_t)SB / TR::GROUP; ++g) encode_group(g, std::false_type{});
        }
      else
        {
#pragma unroll 1
        for (uint32_t g = 0; g < (uint32_t)SB / TR::GROUP; ++g)
          {
          if (g * TR::GROUP >= todo) break;
          encode_group(g, std::true_type{});
          }
        }
      __syncthreads();                               // the staging tile may be overwritten; window words are visible
      if (DEFER && pv_valid)
        {
        if (!pv_have_base && i0 >= (uint32_t)SB) pv_have_base = true;       // warp 0 stored the base before the barrier
        if (pv_have_base)
          {
          if (pv_loaded)
            {
            asm volatile("cp.async.wait_group 0;" ::: "memory");          // requested a whole encode ago
            __syncwarp();
            pv_write();
            }
          pv_fetch();
          }
        }
      if (i0 + SB < cnt0) stage_in(i0 + SB);         // next slab crosses L2 -> shared memory while the windows drain

      // c. completed vectors of the warp's 32 windows -> the chunks' scratch slots (warp-wide, so the
      //    stores are whole sectors; every lane flushing its own 16-byte pieces was measured slower)
      const uint32_t nf = wp >> 2;
#pragma unroll
      for (int it = 0; it < WV; ++it)
        {
        const uint32_t ci = (uint32_t)it * 32u + lane;
        const uint32_t L = ci / (uint32_t)WV, v = ci - L * (uint32_t)WV;
        const uint32_t nfL = __shfl_sync(FULL, nf, L), posL = __shfl_sync(FULL, flushed, L);
        if (v < nfL)
          st_scratch(cta_scr + (size_t)((rgrp * 32 + L) * NCOMP + c) * a.slot + posL + 16u * v, wwarp4[ci]);
        }
      __syncwarp();
      if (nf)
        { // keep the incomplete tail vector at the front of the window
        const uint32_t rem = wp & 3u;
        uint32_t t0 = 0, t1 = 0, t2 = 0;
        if (rem > 0) t0 = wrow[4 * nf];
        if (rem > 1) t1 = wrow[4 * nf + 1];
        if (rem > 2) t2 = wrow[4 * nf + 2];
        if (rem > 0) wrow[0] = t0;
        if (rem > 1) wrow[1] = t1;
        if (rem > 2) wrow[2] = t2;
        flushed += 16u * nf; wp = rem;
        }
      }

    // d. tail of every chunk, sizes, offsets, assembly
    uint32_t total = 0;
    if (cnt)
      {
      if (fill) wrow[wp] = acc;
      const uint32_t tailb = 4u * wp + fill;                          // < 16
      total = flushed + tailb;
      if (tailb) st_scratch(my_scr + flushed, *reinterpret_cast<const uint4*>(wrow));
      }
    sh_size[klocal * NCOMP + c] = total;
    __syncthreads();
      {
      const uint32_t mine = sh_size[threadIdx.x];
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
        {
        const uint32_t up = __shfl_up_sync(FULL, incl, o);
        if (lane >= (unsigned)o) incl += up;
        }
      if (lane == 31) sh_wsum[warp] = incl;
      __syncthreads();
      uint32_t wbase = 0, tsum = 0;
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) { const uint32_t t = sh_wsum[w]; if (w < (int)warp) wbase += t; tsum += t; }
      sh_off[threadIdx.x] = wbase + incl - mine;
      if (DEFER)
        { // u16 size table, stream order
        const uint64_t gq = (uint64_t)tile * NTHREADS + threadIdx.x;
        if (gq < (uint64_t)a.nranges * NCOMP)
          {
          uint8_t* sz = a.sizes + 2 * gq;
          sz[0] = (uint8_t)mine; sz[1] = (uint8_t)(mine >> 8);
          }
        if (warp == 0) lookback_publish(a.desc, tile, tsum);      // later tiles can add this one up; its own base is looked up during the next tile
        pv_finish();                                              // what is left of the tile before this one
        __syncthreads();                                          // nobody reads the previous tile's tables or bounce buffers any more
        sh_size_p[threadIdx.x] = mine;
        sh_off_p[threadIdx.x] = wbase + incl - mine;
        if (threadIdx.x == 0) sh_tsum_p = tsum;
        pv_valid = true; pv_have_base = false; pv_loaded = false;
        pv_tile = tile; pv_set = cur_set; pv_next = warp;
        cur_set ^= 1u;
        continue;
        }
      if (warp == 0)
        {
        const uint64_t excl = lookback_exclusive(a.desc, tile, tsum);
        if (lane == 0)
          {
          sh_base = excl;
          if (tile == a.ntiles - 1)
            {
            *a.total = excl + tsum;
            store_u64_bytes(a.total_field, excl + tsum);
            }
          }
        }
      // u16 size table, stream order
      const uint64_t gq = (uint64_t)tile * NTHREADS + threadIdx.x;
      if (gq < (uint64_t)a.nranges * NCOMP)
        {
        uint8_t* sz = a.sizes + 2 * gq;
        sz[0] = (uint8_t)mine; sz[1] = (uint8_t)(mine >> 8);
        }
      }
    constexpr int BV = 5;                                  // vectors per lane of the bounce path: chunks up to 2560 bytes
    const bool bounce_path = a.slot <= 32u * BV * 16u && a.slot <= 32u * WV * 16u;
    uint4 cur[BV], nxt[BV];
    auto fetch = [&](uint4 (&r)[BV], uint32_t q)
      {
      const uint32_t nv = (sh_size[q] + 15u) >> 4;
      const uint4* s4 = reinterpret_cast<const uint4*>(cta_scr + (size_t)q * a.slot);
#pragma unroll
      for (int u = 0; u < BV; ++u) if (lane + 32u * u < nv) r[u] = ld_scratch(s4 + lane + 32 * u);
      };
    if (bounce_path) fetch(cur, warp);                     // the first chunk's bytes travel while the look-back finishes
    __syncthreads();
    const uint64_t base = sh_base;
    if (bounce_path)
      { // slot -> registers (the next chunk's loads are in flight while this one is written) -> the warp's
        // window area -> final offset; the shared-memory hop turns the alignment shift into cheap LDS work
      uint8_t* bounce = reinterpret_cast<uint8_t*>(win + (size_t)warp * 32 * WV * 4);
      for (uint32_t q = warp; q < (uint32_t)NTHREADS; q += NWARPS)
        {
        if (q + NWARPS < (uint32_t)NTHREADS) fetch(nxt, q + NWARPS);
        const uint32_t nbytes = sh_size[q];
        const uint32_t nv = (nbytes + 15u) >> 4;
#pragma unroll
        for (int u = 0; u < BV; ++u) if (lane + 32u * u < nv) reinterpret_cast<uint4*>(bounce)[lane + 32 * u] = cur[u];
        __syncwarp();
        if (nbytes) warp_copy_smem_to_global(a.payload + base + sh_off[q], bounce, nbytes);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < BV; ++u) cur[u] = nxt[u];
        }
      }
    else
      for (uint32_t q = warp; q < (uint32_t)NTHREADS; q += NWARPS)
        {
        const uint32_t nbytes = sh_size[q];
        if (nbytes) warp_copy_global(a.payload + base + sh_off[q], cta_scr + (size_t)q * a.slot, nbytes);
        }
    }
  if (DEFER) pv_finish();
  }

// ---------------------------------------------------------------------------------------------
// K4L: legacy (reference v0) stream decoder.  The whole stream is ONE serial chain
// (fpc.c:246-327): value j needs the decoded value j-1 for both table look-ups, so a stream is
// one dependency chain of ~50 cycles per value whatever the hardware.  What the warp can do is
// keep everything else off that chain (block b decodes stream b):
//   * the compressed bytes travel through a 4 KiB ring in shared memory, filled 2 KiB ahead with
//     16-byte cp.async copies (the old kernel read every residual byte from global memory);
//   * per step of 32 values all lanes walk the code words of the step (4 groups of 8 floats,
//     16 groups of 2 doubles), lane j then gathers the residual of value j;
//   * lane 0 alone runs the predictor over the 32 values, taking each residual from its lane by
//     a shuffle (issued ahead, off the chain) and handing each value back the same way;
//   * the 32 values leave as one strided warp store.
// Tables: shared memory when they fit ((4,10) floats: 4 KiB), else global scratch zeroed by the
// host ((20,20) doubles: 16 MiB per stream - every look-up is then an L2 round trip).
// ---------------------------------------------------------------------------------------------
struct FpcLegacyDecodeArgs
  {
  const uint8_t* const* streams;   // device array of nstreams pointers to streams (hash_info, n, groups)
  int nstreams;
  void* out;                       // element j of stream c goes to out[j * stride + c]
  uint32_t stride;
  void* gtables;                   // nullptr or nstreams * table words of zeroed scratch
  uint64_t gtable_words;           // per stream
  uint32_t* counts;                // [nstreams] decoded value counts (from the stream headers)
  uint32_t expect;                 // values the caller has room for per stream
  const uint64_t* extents;         // [nstreams] bytes of the buffer readable from each stream's first byte (0: unknown)
  };

constexpr uint32_t FPC_LEGACY_RING = 4096;

template <typename W>
__global__ void __launch_bounds__(64)
fpc_decode_legacy_kernel(const FpcLegacyDecodeArgs a)
  {
  // Two warps.  Warp 1 FEEDS: it keeps the ring filled, parses the code words of step s and gathers
  // its residuals into one half of a double buffer, and writes the values of step s - 2 out.  Lane 0
  // of warp 0 runs the CHAIN of step s - 1 meanwhile - nothing but the two dependent table reads,
  // the xor and the hash updates per value (fpc.c:308-326).  One CTA barrier per 32 values.
  using TR = FpcTraits<W>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;                                        // FPC_LEGACY_RING bytes
  __shared__ W sh_x[2][32], sh_v[2][32];
  __shared__ uint32_t sh_u2[2][32];
  const unsigned c = blockIdx.x, lane = lane_id(), warp = threadIdx.x >> 5;
  const uint8_t* p = a.streams[c];
  const int e1 = (p[0] >> 4) << 1, e2 = (p[0] & 15) << 1;          // fpc.c:214-217
  const uint32_t n = ((uint32_t)p[1] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 8) | p[4];
  const size_t nt1 = (size_t)1 << e1, nt2 = (size_t)1 << e2;
  W* T1; W* T2;
  if (a.gtables == nullptr) { T1 = reinterpret_cast<W*>(smem_raw + FPC_LEGACY_RING); for (size_t i = threadIdx.x; i < nt1 + nt2; i += 64) T1[i] = 0; }
  else T1 = reinterpret_cast<W*>(a.gtables) + (size_t)c * a.gtable_words;
  T2 = T1 + nt1;
  if (threadIdx.x == 0) a.counts[c] = n;
  const uint32_t todo = n < a.expect ? n : a.expect;
  const uint32_t nsteps = (todo + 31u) / 32u;
  // feeder state: the ring holds stream bytes [ring_lo, ring_hi) of the 16-byte aligned view of the stream
  const uint8_t* base = p - (reinterpret_cast<uintptr_t>(p) & 15u);
  const uint64_t readable = a.extents && a.extents[c] ? a.extents[c] + (uint64_t)(p - base) : ~0ull;   // from base
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  uint64_t bp = (uint64_t)(p - base) + 5;                          // next unread byte, relative to base
  uint64_t hi = 0;                                                 // bytes [0, hi) have been requested
  auto request = [&](uint64_t upto)
    { // cp.async of whole 16-byte vectors up to `upto` (a multiple of 2048); bytes beyond the buffer are zero filled
    for (uint64_t o = hi + 16ull * lane; o < upto; o += 512)
      {
      const uint32_t ssz = o + 16 <= readable ? 16u : (o < readable ? (uint32_t)(readable - o) : 0u);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(ring_s + (uint32_t)(o & (FPC_LEGACY_RING - 1))), "l"(base + (ssz ? o : 0)), "r"(ssz) : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
    hi = upto;
    };
  auto rb = [&](uint64_t o) -> uint32_t { return ring[o & (FPC_LEGACY_RING - 1)]; };
  if (warp == 1)
    {
    request(FPC_LEGACY_RING);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
  __syncthreads();

  FpcLaneState<W> st; st.pred1 = 0; st.pred2 = 0; st.last = 0; st.c1 = 0; st.c2 = 0;
  const uint32_t m2 = (uint32_t)nt2 - 1;
  W* out = reinterpret_cast<W*>(a.out) + c;
  constexpr int GPS = 32 / TR::GROUP;                              // groups per step
  const uint32_t gl = lane % TR::GROUP, gi = lane / TR::GROUP;
  for (uint32_t s = 0; s < nsteps + 2u; ++s)
    {
    if (warp == 1)
      {
      if (s < nsteps)
        { // ---- feed step s ----
        // keep two kilobytes ahead: the half of the ring behind the read position is free
        const uint64_t target = ((bp >> 11) + 2) << 11;
        if (target > hi) request(target);
        if (bp + 320 > hi - 2048) asm volatile("cp.async.wait_group 0;" ::: "memory");      // the step reaches into the newest half
        else asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        // code words of the step: every lane walks them, lane j keeps the place of residual j
        uint64_t gp = bp, my_at = 0;
        uint32_t my_nb = 0;
        bool my_use2 = false;
#pragma unroll 1
        for (int g = 0; g < GPS; ++g)
          {
          uint32_t bc = 0;
#pragma unroll
          for (int b = 0; b < TR::HDR; ++b) bc = (bc << 8) | rb(gp + b);
          uint32_t sum = 0;
#pragma unroll
          for (int jj = 0; jj < TR::GROUP; ++jj)
            {
            const uint32_t code = (bc >> (TR::CBITS * jj)) & ((1u << TR::CBITS) - 1u);
            const uint32_t nb = code > (uint32_t)TR::BASE2 ? code - TR::BASE2 : code;
            if ((uint32_t)g == gi && (uint32_t)jj == gl) { my_at = gp + TR::HDR + sum; my_nb = nb; my_use2 = code > (uint32_t)TR::BASE2; }
            sum += nb;
            }
          gp += TR::HDR + sum;
          }
        // residual of value 32 s + lane, big-endian
        W x = 0;
        for (uint32_t b = 0; b < my_nb; ++b) x = (W)(x << 8) | (W)rb(my_at + b);
        sh_x[s & 1u][lane] = x; sh_u2[s & 1u][lane] = my_use2 ? 1u : 0u;
        bp = gp;          // (the pad slots of a last, incomplete group do not matter any more)
        }
      if (s >= 2u)
        { // ---- values of step s - 2 leave ----
        const uint32_t i = 32u * (s - 2u) + lane;
        if (i < todo) out[(size_t)i * a.stride] = sh_v[s & 1u][lane];
        }
      }
    else if (lane == 0 && s >= 1u && s <= nsteps)
      { // ---- the chain of step s - 1 ----
      const uint32_t q = (s - 1u) & 1u;
      const uint32_t i0 = 32u * (s - 1u);
      const uint32_t cnt = todo - i0 < 32u ? todo - i0 : 32u;
      if (cnt == 32u)
        {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) sh_v[q][j] = fpc_decode_value<W, 1>(st, sh_x[q][j], sh_u2[q][j] != 0u, T1, T2, e1, e2, m2);
        }
      else
        for (uint32_t j = 0; j < cnt; ++j) sh_v[q][j] = fpc_decode_value<W, 1>(st, sh_x[q][j], sh_u2[q][j] != 0u, T1, T2, e1, e2, m2);
      }
    __syncthreads();
    }
  }

} // namespace tb200
