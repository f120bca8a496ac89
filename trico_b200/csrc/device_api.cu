// device_api.cu - thin extern "C" layer over the sm_100a kernels (include/trico_b200_device.h).
// Owns the CUDA stream, the per-context workspace and every kernel launch.  No CPU fallback.
#include "../../include/trico_b200_device.h"
#include "../../include/trico_b200_io.h"

#include "common.cuh"
#include "fpc.cuh"
#include "lz4.cuh"
#include "lz4_lanes.cuh"
#include "lz4_multi.cuh"
#include "lz4_v0.cuh"
#include "planes.cuh"
#include "stl.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

using namespace tb200;

static thread_local char g_err[512] = "";

static int fail(const char* what, cudaError_t e)
  {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return 0;
  }
static int fail_msg(const char* what)
  {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return 0;
  }
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(#call, e__); } while (0)

extern "C" const char* tb200_last_error(void) { return g_err; }

struct tb200_ctx
  {
  int device;
  cudaStream_t stream;
  bool own_stream;
  uint8_t* ws;            // zeroed scratch, handed out in slices: [ticket (16 B) | descriptors] per launch
  size_t ws_bytes, ws_used;
  uint8_t* big;           // large scratch (legacy tables, plane buffers)
  size_t big_bytes;
  uint64_t launches;
  int max_smem_optin;
  // batches of streams run on several CUDA streams: sub-contexts, created on first use (device_api_misc.inc)
  tb200_ctx* sub[31];
  int nsub;
  cudaEvent_t ev_fork, ev_join[31];
  };

extern "C" int tb200_device_count(void)
  {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
  }

extern "C" tb200_ctx* tb200_ctx_create(int device, void* cuda_stream)
  {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { fail("cudaGetDeviceCount (no CUDA device: the B200 path has no CPU fallback)", e); return nullptr; }
  if (device < 0 || device >= n) { fail_msg("tb200_ctx_create: bad device index"); return nullptr; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { fail("cudaSetDevice", e); return nullptr; }
  tb200_ctx* c = (tb200_ctx*)calloc(1, sizeof(tb200_ctx));
  c->device = device;
  if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
  else
    {
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { fail("cudaStreamCreate", e); free(c); return nullptr; }
    c->own_stream = true;
    }
  cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  // (Measured: reserving a persisting L2 carve-out for the encoders' evict_last scratch lines makes
  // every kernel slower - C2 FPC decode 0.59 -> 1.18 ms - so the device default, none, stays.)
  return c;
  }

extern "C" void tb200_ctx_destroy(tb200_ctx* c)
  {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < c->nsub; ++i) tb200_ctx_destroy(c->sub[i]);
  if (c->nsub) { cudaEventDestroy(c->ev_fork); for (int i = 0; i < c->nsub; ++i) cudaEventDestroy(c->ev_join[i]); }
  if (c->ws) cudaFree(c->ws);
  if (c->big) cudaFree(c->big);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  free(c);
  }

extern "C" void* tb200_ctx_stream(tb200_ctx* c) { return (void*)c->stream; }
extern "C" int tb200_ctx_device(tb200_ctx* c) { return c->device; }
extern "C" int tb200_ctx_make_current(tb200_ctx* c) { CK(cudaSetDevice(c->device)); return 1; }
extern "C" int tb200_set_device(int device) { CK(cudaSetDevice(device)); return 1; }
// returns the device buffers a context grew to the driver (idle workers of the archive layer)
extern "C" void tb200_ctx_trim(tb200_ctx* c)
  {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->ws) cudaFree(c->ws);
  if (c->big) cudaFree(c->big);
  c->ws = nullptr; c->ws_bytes = 0; c->ws_used = 0; c->big = nullptr; c->big_bytes = 0;
  }
extern "C" uint64_t tb200_ctx_launch_count(tb200_ctx* c) { return c->launches; }

extern "C" int tb200_ctx_sync(tb200_ctx* c)
  {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return 1;
  }

// workspace for one launch: 16-byte ticket block followed by `ntiles` 64-bit descriptors, zeroed.
// tb200_encode_stream sets this around the codec call: a launcher that ends in lz4_assemble_kernel hands
// the header fields to it (and says so); otherwise tb200_encode_stream launches the header kernel itself
struct PendingHeader { uint8_t* out; uint32_t count; int type, info, log2; uint64_t fixed_plus_table; bool done; };
static thread_local PendingHeader* tl_header = nullptr;
static Lz4StreamHeader take_header()
  {
  Lz4StreamHeader a;
  if (!tl_header) return a;
  a.hdr = tl_header->out; a.hdr_count = tl_header->count; a.hdr_type = (uint8_t)tl_header->type; a.hdr_info = (uint8_t)tl_header->info;
  a.hdr_log2 = (uint8_t)tl_header->log2; a.hdr_fixed_plus_table = tl_header->fixed_plus_table;
  tl_header->done = true;
  return a;
  }

// Every launch gets a FRESH zeroed slice of the workspace; the buffer is cleared as a whole when it
// has been used up (stream order: after every kernel that looked at the old slices).  A stream of a
// few hundred kilobytes needs well under a kilobyte, so a batch of small streams pays one memset per
// few hundred streams instead of one per stream (a launch-bound regime: DESIGN.md, batched streams).
static int ws_prepare(tb200_ctx* c, uint64_t ntiles, uint32_t** ticket, uint64_t** desc)
  {
  const size_t need = 16 + (size_t)ntiles * 8;
  const size_t take = (need + 255) & ~(size_t)255;
  if (take > c->ws_bytes)
    {
    // the previous buffer may still be in use by queued kernels
    CK(cudaStreamSynchronize(c->stream));
    if (c->ws) CK(cudaFree(c->ws));
    c->ws = nullptr; c->ws_bytes = 0; c->ws_used = 0;
    size_t cap = take + take / 2 + 4096;
    if (cap < (256u << 10)) cap = 256u << 10;            // small streams never come back here (cudaFree stalls every thread's stream)
    cap = (cap + 255) & ~(size_t)255;
    CK(cudaMalloc((void**)&c->ws, cap));
    c->ws_bytes = cap;
    CK(cudaMemsetAsync(c->ws, 0, cap, c->stream));
    }
  if (c->ws_used + take > c->ws_bytes)
    {
    CK(cudaMemsetAsync(c->ws, 0, c->ws_used, c->stream));
    c->ws_used = 0;
    }
  uint8_t* p = c->ws + c->ws_used;
  c->ws_used += take;
  *ticket = reinterpret_cast<uint32_t*>(p);
  *desc = reinterpret_cast<uint64_t*>(p + 16);
  return 1;
  }

static int big_prepare(tb200_ctx* c, size_t need, uint8_t** out)
  {
  if (need > c->big_bytes)
    {
    CK(cudaStreamSynchronize(c->stream));
    if (c->big) CK(cudaFree(c->big));
    c->big = nullptr; c->big_bytes = 0;
    size_t cap = need + need / 2 + 4096;             // streams of growing size must not reallocate every time
    if (cap < (64u << 20)) cap = 64u << 20;
    CK(cudaMalloc((void**)&c->big, cap));
    c->big_bytes = cap;
    }
  *out = c->big;
  return 1;
  }

// Per-launch host work that does not change from call to call - the shared-memory opt-in, the
// occupancy query - is remembered: streams of a few hundred kilobytes are launch-overhead bound.
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is process-wide per (kernel, device), so the
// opt-in is a process-global maximum behind a mutex and is only ever RAISED: two archives on two
// threads that need different sizes of the same kernel (16 KiB index blocks, 8 KiB colour blocks)
// can never lower each other's limit.  Only the occupancy query is memoised per thread.
#include <mutex>
struct optin_entry { const void* kernel; int device; size_t bytes; };
static std::mutex g_optin_lock;
static optin_entry g_optin[256];
static int g_noptin = 0;

struct launch_memo { const void* kernel; size_t smem; int threads, device, per_sm, sms; };
static launch_memo* memo_for(const void* kernel, size_t smem, int threads, int device)
  {
  static thread_local launch_memo table[64];
  launch_memo* m = &table[((reinterpret_cast<uintptr_t>(kernel) >> 4) ^ (smem >> 8)) % 64];
  if (!(m->kernel == kernel && m->smem == smem && m->threads == threads && m->device == device))
    { m->kernel = kernel; m->smem = smem; m->threads = threads; m->device = device; m->per_sm = -1; }
  return m;
  }

template <typename K>
static int set_smem(K kernel, size_t bytes, const tb200_ctx* c)
  {
  if (bytes > (size_t)c->max_smem_optin) return fail_msg("kernel needs more shared memory than the device offers");
  // (static + dynamic) above 48 KiB needs the opt-in; the kernels carry up to ~2 KiB of static shared memory
  if (bytes <= 40 * 1024) return 1;
  const void* key = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> guard(g_optin_lock);
  optin_entry* e = nullptr;
  for (int i = 0; i < g_noptin; ++i) if (g_optin[i].kernel == key && g_optin[i].device == c->device) { e = &g_optin[i]; break; }
  if (e && e->bytes >= bytes) return 1;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (!e && g_noptin < 256) { e = &g_optin[g_noptin++]; e->kernel = key; e->device = c->device; }
  if (e) e->bytes = bytes;          // table full: the attribute is simply set again next time
  return 1;
  }

// the assembly kernel of the LZ4 path (and of small FPC streams): streams whose blocks can be large get a ring
// of shared-memory stages, one scratch slot each, that bulk copies (TMA) fill ahead of the CTA
static int launch_lz4_assemble(tb200_ctx* c, const Lz4EncodeArgs& a, uint64_t nchunks, unsigned ntiles)
  {
  Lz4StreamHeader h = take_header();
  size_t smem = 0;
  static const int use_ring = getenv("TB200_LZ4_ASM_NO_RING") ? 0 : 1;
  if (use_ring && a.slot >= LZ4_ASM_BIG && (a.slot & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.scratch) & 15u) == 0)
    {
    h.stages = (size_t)a.slot * 4 <= (72u << 10) ? 4u : ((size_t)a.slot * 2 <= (72u << 10) ? 2u : 0u);
    smem = (size_t)h.stages * a.slot;
    if (smem && !set_smem(lz4_assemble_kernel, smem, c)) return 0;
    }
  lz4_assemble_kernel<<<ntiles, LZ4_ASM_THREADS, smem, c->stream>>>(a, nchunks, h);
  return 1;
  }

// resident CTAs per SM and SM count for a persistent grid
template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, const tb200_ctx* c, int* per_sm, int* sms)
  {
  launch_memo* m = memo_for(reinterpret_cast<const void*>(kernel), smem, threads, c->device);
  if (m->per_sm < 0)
    {
    int p = 0, n = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p, kernel, threads, smem));
    CK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, c->device));
    m->per_sm = p; m->sms = n;
    }
  *per_sm = m->per_sm; *sms = m->sms;
  return 1;
  }

// ------------------------------------------------------------------------------------------------
// stream layouts: trico/trico.h:11-34, writers trico/trico.c:215-858
// ------------------------------------------------------------------------------------------------
extern "C" int tb200_stream_layout(int type, int* wordsize, int* ncomp, int* per_count)
  {
  // codec, wordsize, ncomp, per_count
  static const int8_t T[21][4] = {
    {0,0,0,0},
    {1,4,3,1},{1,8,3,1},{2,4,1,3},{2,8,1,3},{1,4,2,1},{1,8,2,1},{1,4,2,1},{1,8,2,1},
    {1,4,3,1},{1,8,3,1},{1,4,3,1},{1,8,3,1},{2,4,1,1},{2,4,1,1},{1,4,1,1},{1,8,1,1},
    {2,1,1,1},{2,2,1,1},{2,4,1,1},{2,8,1,1}};
  if (type < 1 || type > 20) return 0;
  if (wordsize) *wordsize = T[type][1];
  if (ncomp) *ncomp = T[type][2];
  if (per_count) *per_count = T[type][3];
  return T[type][0];
  }

extern "C" int tb200_default_log2_chunk(int type, uint32_t count)
  {
  int w = 0, nc = 0, pc = 0;
  const int codec = tb200_stream_layout(type, &w, &nc, &pc);
  (void)count;
  if (codec == 1) return w == 4 ? 9 : 8;     // 512 floats / 256 doubles per chunk (ratio cost <= 1 %)
  if (codec == 2)
    {
    // 16 KiB plane blocks for index streams and the 1- and 4-byte attribute lists; 8 KiB for 8-byte
    // elements (8 planes share one CTA) and for colours and 2-byte lists: those are mostly noise or
    // very many short sequences, which a warp parses serially - twice the blocks in flight is
    // +30..75 % there for 0.0..0.4 % of ratio (u8 lists with long-range repeats lose 40 %: they stay
    // at 16 KiB).  TB200_LZ4_LOG2B overrides (experiments).
    static int env = -1;
    if (env < 0) { const char* e = getenv("TB200_LZ4_LOG2B"); env = e ? atoi(e) : 0; }
    if (env >= 8 && env <= 15) return w == 8 ? (env > 14 ? 14 : env) : env;
    if (w == 8 || w == 2 || type == 13 || type == 14) return 13;
    return 14;
    }
  return 0;
  }

extern "C" uint64_t tb200_v1_nchunks(int type, uint32_t count, int log2_chunk)
  {
  int w = 0, nc = 0, pc = 0;
  const int codec = tb200_stream_layout(type, &w, &nc, &pc);
  if (!codec) return 0;
  const uint64_t n = (uint64_t)count * pc;
  const uint64_t nr = (n + ((uint64_t)1 << log2_chunk) - 1) >> log2_chunk;
  return nr * (codec == 1 ? nc : w);
  }

extern "C" uint64_t tb200_v1_stream_bound(int type, uint32_t count, int log2_chunk)
  {
  int w = 0, nc = 0, pc = 0;
  const int codec = tb200_stream_layout(type, &w, &nc, &pc);
  if (!codec) return 0;
  const uint64_t nch = tb200_v1_nchunks(type, count, log2_chunk);
  const uint32_t S = 1u << log2_chunk;
  const uint64_t per = codec == 1 ? fpc_chunk_bound(S, w) : lz4_block_bound(S);
  return TB200_V1_FIXED_BYTES + nch * (2 + per) + 64;
  }

// ------------------------------------------------------------------------------------------------
// chunked FPC
// ------------------------------------------------------------------------------------------------
template <typename W, int NCOMP, int R, int SB, int EXP = -1, int DEFER = -1>
static int launch_fpc_encode_lanes(tb200_ctx* c, FpcEncodeArgs a)
  {
  using WIN = FpcEncWindow<W, SB>;
  constexpr int NWARPS = NCOMP * R;
  const uint32_t S = 1u << a.log2S;
  const uint32_t slot = (fpc_chunk_bound(S, sizeof(W)) + 16u + 15u) & ~15u;
  if (EXP < 0)   // the archive default exponents run a build with the shifts and masks compiled in
    return (a.e1 == 2 && a.e2 == 4) ? launch_fpc_encode_lanes<W, NCOMP, R, SB, 0x0204>(c, a) : launch_fpc_encode_lanes<W, NCOMP, R, SB, 0>(c, a);
  constexpr int EX = EXP < 0 ? 0 : EXP;
  if (DEFER < 0)
    { // float vec3 streams of default geometry drain every tile during the next one (a bounce buffer per
      // warp in shared memory, two sets of scratch slots); the other shapes would lose a resident CTA to it
    constexpr bool CAN = sizeof(W) == 4 && NCOMP == 3 && R == 1;
    static int off = -1;
    if (off < 0) { const char* e = getenv("TB200_FPC_ENC_NODEFER"); off = e && e[0] == '1'; }
    if (CAN && !off && slot <= 2560u) return launch_fpc_encode_lanes<W, NCOMP, R, SB, EX, CAN ? 1 : 0>(c, a);
    return launch_fpc_encode_lanes<W, NCOMP, R, SB, EX, 0>(c, a);
    }
  constexpr bool DF = DEFER > 0;
  a.ntiles = (a.nranges + 32 * R - 1) / (32 * R);
  if (!ws_prepare(c, a.ntiles + 1, &a.ticket, &a.desc)) return 0;
  if (!a.total_field) a.total_field = reinterpret_cast<uint8_t*>(a.desc + a.ntiles);
  const size_t smem = (size_t)NWARPS * 32 * WIN::VECS * 16 +
                      (size_t)32 * R * fpc_stage_row_words_enc(SB, NCOMP, sizeof(W)) * 4 +
                      (size_t)NWARPS * 32 * ((1u << a.e1) + (1u << a.e2)) * sizeof(W) +
                      (DF ? (size_t)NWARPS * slot : 0);
  if (!set_smem(fpc_encode_lanes_kernel<W, NCOMP, R, SB, EX, DF>, smem, c)) return 0;
  int per_sm = 0, sms = 0;
  if (!persistent_grid(fpc_encode_lanes_kernel<W, NCOMP, R, SB, EX, DF>, NWARPS * 32, smem, c, &per_sm, &sms)) return 0;
  if (per_sm < 1) return fail_msg("fpc_encode_lanes_kernel does not fit on an SM");
  // persistent grid, every CTA resident (the look-back relies on it); a CTA reuses its scratch slots
  static int ctas_cap = -1;                                                              // experiments
  if (ctas_cap < 0) { const char* e = getenv("TB200_FPC_ENC_CTAS"); ctas_cap = e ? atoi(e) : 0; }
  if (ctas_cap >= 1 && ctas_cap < per_sm) per_sm = ctas_cap;
  uint32_t grid = (uint32_t)per_sm * (uint32_t)sms;
  if (grid > a.ntiles) grid = a.ntiles;
  a.slot = slot;
  uint8_t* scratch = nullptr;
  if (!big_prepare(c, (size_t)grid * (DF ? 2 : 1) * NWARPS * 32 * a.slot + 64, &scratch)) return 0;
  a.scratch = scratch;
  fpc_encode_lanes_kernel<W, NCOMP, R, SB, EX, DF><<<grid, NWARPS * 32, smem, c->stream>>>(a);
  c->launches++;
  CK(cudaGetLastError());
  return 1;
  }

static int launch_fpc_encode_small(tb200_ctx* c, const FpcEncodeArgs& f, int wordsize, int ncomp)
  {
  const uint32_t S = 1u << f.log2S;
  const uint64_t nchunks = (uint64_t)f.nranges * ncomp;
  const unsigned ntiles = (unsigned)((nchunks + LZ4_ASM_TILE - 1) / LZ4_ASM_TILE);
  Lz4EncodeArgs a;                                  // the assembly kernel of the LZ4 path: chunk g at scratch + g * slot, u16 sizes
  a.in = nullptr; a.n = 0; a.nranges = 0; a.log2B = 0; a.dense_list = nullptr; a.dense_count = nullptr; a.dbg = nullptr;
  a.sizes = f.sizes; a.payload = f.payload; a.total = f.total;
  a.slot = (fpc_chunk_bound(S, wordsize) + 16u + 15u) & ~15u;
  if (!ws_prepare(c, (uint64_t)ntiles + 1, &a.ticket, &a.desc)) return 0;
  a.total_field = f.total_field ? f.total_field : reinterpret_cast<uint8_t*>(a.desc + ntiles);
  uint8_t* scratch = nullptr;
  if (!big_prepare(c, (size_t)nchunks * a.slot + 64, &scratch)) return 0;
  a.scratch = scratch;
  FpcChunksArgs k;
  k.in = f.in; k.n = f.n; k.nranges = f.nranges; k.log2S = f.log2S; k.e1 = f.e1; k.e2 = f.e2; k.ncomp = ncomp;
  k.sizes = f.sizes; k.scratch = scratch; k.slot = a.slot;
  const size_t tw = ((size_t)1 << f.e1) + ((size_t)1 << f.e2);
  const size_t smem = ((((tw * wordsize) + 15) & ~(size_t)15) + a.slot + 16) * FPC_CHUNKS_WARPS;
  const unsigned grid = (unsigned)((nchunks + FPC_CHUNKS_WARPS - 1) / FPC_CHUNKS_WARPS);
  if (wordsize == 4)
    {
    if (!set_smem(fpc_encode_chunks_kernel<uint32_t>, smem, c)) return 0;
    fpc_encode_chunks_kernel<uint32_t><<<grid, FPC_CHUNKS_WARPS * 32, smem, c->stream>>>(k);
    }
  else
    {
    if (!set_smem(fpc_encode_chunks_kernel<uint64_t>, smem, c)) return 0;
    fpc_encode_chunks_kernel<uint64_t><<<grid, FPC_CHUNKS_WARPS * 32, smem, c->stream>>>(k);
    }
  if (!launch_lz4_assemble(c, a, nchunks, ntiles)) return 0;
  c->launches += 2;
  CK(cudaGetLastError());
  return 1;
  }

extern "C" int tb200_fpc_encode(tb200_ctx* c, int wordsize, int ncomp, const void* d_in, uint64_t n, int log2_chunk,
                                int e1, int e2, uint8_t* d_sizes, uint8_t* d_payload, uint8_t* d_total_field, uint64_t* d_total)
  {
  CK(cudaSetDevice(c->device));
  if ((wordsize != 4 && wordsize != 8) || ncomp < 1 || ncomp > 3) return fail_msg("tb200_fpc_encode: bad wordsize/ncomp");
  if (e1 < 2 || e2 < 2 || e1 > 4 || e2 > 6 || (e1 & 1) || (e2 & 1)) return fail_msg("tb200_fpc_encode: chunk exponents must be even, e1 in 2..4, e2 in 2..6");
  if (log2_chunk < 5 || fpc_chunk_bound(1u << log2_chunk, wordsize) > 65535u) return fail_msg("tb200_fpc_encode: bad chunk size");
  FpcEncodeArgs a;
  a.in = d_in; a.n = n; a.log2S = log2_chunk; a.e1 = e1; a.e2 = e2;
  a.nranges = (uint32_t)((n + ((uint64_t)1 << log2_chunk) - 1) >> log2_chunk);
  a.ntiles = 0; a.ticket = nullptr; a.desc = nullptr; a.scratch = nullptr; a.slot = 0;
  a.sizes = d_sizes; a.payload = d_payload; a.total = d_total; a.total_field = d_total_field;
  if (a.nranges == 0)
    {
    CK(cudaMemsetAsync(d_total, 0, 8, c->stream));
    if (d_total_field) CK(cudaMemsetAsync(d_total_field, 0, 8, c->stream));
    return 1;
    }
  // K3S: a stream of a few thousand chunks is all latency for the lane-per-chunk kernel (every lane
  // encodes 512 values one after the other); a warp per chunk finishes in a tenth of the time
  static int small_max = -1;
  if (small_max < 0) { const char* e = getenv("TB200_FPC_SMALL_CHUNKS"); small_max = e ? atoi(e) : 8192; }
  const uint64_t nchunks = (uint64_t)a.nranges * ncomp;
  if (nchunks <= (uint64_t)small_max) return launch_fpc_encode_small(c, a, wordsize, ncomp);
  if (wordsize == 4)
    return ncomp == 3 ? launch_fpc_encode_lanes<uint32_t, 3, 1, 32>(c, a) : ncomp == 2 ? launch_fpc_encode_lanes<uint32_t, 2, 2, 32>(c, a) : launch_fpc_encode_lanes<uint32_t, 1, 4, 32>(c, a);
  return ncomp == 3 ? launch_fpc_encode_lanes<uint64_t, 3, 1, 16>(c, a) : ncomp == 2 ? launch_fpc_encode_lanes<uint64_t, 2, 2, 16>(c, a) : launch_fpc_encode_lanes<uint64_t, 1, 4, 16>(c, a);
  }

template <typename W, int NCOMP, int R, int SB, int EXP = -1>
static int launch_fpc_decode(tb200_ctx* c, FpcDecodeArgs a)
  {
  if (EXP < 0)
    return (a.e1 == 2 && a.e2 == 4) ? launch_fpc_decode<W, NCOMP, R, SB, 0x0204>(c, a) : launch_fpc_decode<W, NCOMP, R, SB, 0>(c, a);
  constexpr int EX = EXP < 0 ? 0 : EXP;
  using WIN = FpcWindow<W, SB>;
  constexpr int NWARPS = NCOMP * R;
  a.ntiles = (a.nranges + 32 * R - 1) / (32 * R);
  if (!ws_prepare(c, a.ntiles, &a.ticket, &a.desc)) return 0;
  const size_t smem = (size_t)NWARPS * 32 * WIN::VECS * 16 +
                      (size_t)32 * R * fpc_stage_row_words(SB, NCOMP, sizeof(W)) * 4 +
                      (size_t)NWARPS * 32 * ((1u << a.e1) + (1u << a.e2)) * sizeof(W);
  if (!set_smem(fpc_decode_kernel<W, NCOMP, R, SB, EX>, smem, c)) return 0;
  fpc_decode_kernel<W, NCOMP, R, SB, EX><<<a.ntiles, NWARPS * 32, smem, c->stream>>>(a);
  c->launches++;
  CK(cudaGetLastError());
  return 1;
  }

extern "C" int tb200_fpc_decode(tb200_ctx* c, int wordsize, int ncomp, const uint8_t* d_sizes, const uint8_t* d_payload,
                                uint64_t payload_bytes, uint64_t n, int log2_chunk, int e1, int e2, void* d_out)
  {
  CK(cudaSetDevice(c->device));
  if ((wordsize != 4 && wordsize != 8) || ncomp < 1 || ncomp > 3) return fail_msg("tb200_fpc_decode: bad wordsize/ncomp");
  if (e1 < 2 || e2 < 2 || e1 > 4 || e2 > 6 || (e1 & 1) || (e2 & 1)) return fail_msg("tb200_fpc_decode: chunk exponents must be even, e1 in 2..4, e2 in 2..6");
  if (log2_chunk < 5 || log2_chunk > 13) return fail_msg("tb200_fpc_decode: bad chunk size");
  if (n == 0) return 1;
  FpcDecodeArgs a;
  a.sizes = d_sizes; a.payload = d_payload; a.payload_bytes = payload_bytes; a.n = n;
  a.nranges = (uint32_t)((n + ((uint64_t)1 << log2_chunk) - 1) >> log2_chunk);
  a.ntiles = 0; a.log2S = log2_chunk; a.e1 = e1; a.e2 = e2; a.out = d_out; a.desc = nullptr; a.ticket = nullptr;
  if (wordsize == 4)
    return ncomp == 3 ? launch_fpc_decode<uint32_t, 3, 1, 32>(c, a) : ncomp == 2 ? launch_fpc_decode<uint32_t, 2, 2, 32>(c, a) : launch_fpc_decode<uint32_t, 1, 4, 32>(c, a);
  return ncomp == 3 ? launch_fpc_decode<uint64_t, 3, 1, 16>(c, a) : ncomp == 2 ? launch_fpc_decode<uint64_t, 2, 2, 16>(c, a) : launch_fpc_decode<uint64_t, 1, 4, 16>(c, a);
  }

// ------------------------------------------------------------------------------------------------
// reference-format (v0) FPC streams
// ------------------------------------------------------------------------------------------------
extern "C" uint64_t tb200_fpc_v0_bound(int wordsize, uint32_t n)
  {
  return 5 + (wordsize == 4 ? (uint64_t)n * 4 + 3 * (((uint64_t)n + 7) / 8) + 8 : (uint64_t)n * 8 + ((uint64_t)n + 1) / 2 + 2) + 16;
  }

static void norm_exponents(int* e1, int* e2)
  { // floating_point_stream_compression.c:88-93
  *e1 &= ~1; *e2 &= ~1;
  if (*e1 > 30) *e1 = 30;
  if (*e2 > 30) *e2 = 30;
  }

extern "C" int tb200_fpc_encode_v0(tb200_ctx* c, int wordsize, const void* d_in, uint32_t n, uint32_t stride, int nstreams,
                                   int e1, int e2, uint8_t* d_out, uint64_t out_stride, uint32_t* d_nbytes)
  {
  CK(cudaSetDevice(c->device));
  if (wordsize != 4 && wordsize != 8) return fail_msg("tb200_fpc_encode_v0: bad wordsize");
  norm_exponents(&e1, &e2);
  if (e1 < 2 || e2 < 2) return fail_msg("tb200_fpc_encode_v0: exponents below 2 are not supported");
  const size_t tw = ((size_t)1 << e1) + ((size_t)1 << e2);
  // K3L: long float streams whose tables fit shared memory are encoded tile-parallel (same bytes)
  static int tiled = -1;
  if (tiled < 0) { const char* e = getenv("TB200_FPC_V0_TILED"); tiled = e ? atoi(e) : 1; }
  if (tiled && wordsize == 4 && n >= 8u * FPC_V0_TILE && tw <= 4096)
    {
    FpcV0TileArgs t;
    t.in = d_in; t.n = n; t.stride = stride; t.nstreams = nstreams; t.e1 = e1; t.e2 = e2;
    t.out = d_out; t.out_stride = out_stride; t.nbytes = d_nbytes;
    t.ntiles = (n + FPC_V0_TILE - 1) / FPC_V0_TILE;
    const uint32_t mw = (uint32_t)((tw + 31) / 32);
    t.rec_words = (uint32_t)((2 * tw + mw + 3) & ~(size_t)3);
    const size_t tw_pad = (tw + 3) & ~(size_t)3, mw_pad = (mw + 3u) & ~3u;
    const size_t per_warp = (2 * tw_pad + 2 * mw_pad) * 4 + fpc_v0_tile_out_bytes(4, 8, 3);
    const size_t smem = per_warp * FPC_V0_WARPS;
    if (!set_smem(fpc_encode_v0_tiles_kernel<uint32_t>, smem, c)) return 0;
    int per_sm = 0, sms = 0;
    if (!persistent_grid(fpc_encode_v0_tiles_kernel<uint32_t>, FPC_V0_WARPS * 32, smem, c, &per_sm, &sms)) return 0;
    if (per_sm >= 1)
      {
      uint64_t grid = (uint64_t)per_sm * sms;                        // every CTA resident: the look-backs rely on it
      // tiles per run: up to 8, as long as every warp of the grid still gets about four runs
      const uint64_t warps = grid * FPC_V0_WARPS;
      uint64_t run = ((uint64_t)t.ntiles * nstreams) / (4 * warps);
      if (run < 1) run = 1;
      if (run > 16384u / FPC_V0_TILE) run = 16384u / FPC_V0_TILE;     // 16 K values per run
      if (const char* e = getenv("TB200_FPC_V0_RUN")) { const int v = atoi(e); if (v >= 1 && v <= 64) run = v; }
      t.run = (uint32_t)run;
      t.nruns = (uint32_t)((t.ntiles + run - 1) / run);
      const size_t total = (size_t)t.ntiles * nstreams, total_runs = (size_t)t.nruns * nstreams;
      // workspace: [ticket 256 B][desc: u64 per tile][state: u32 per run][records per run]
      const size_t off_desc = 256, off_state = off_desc + total * 8, off_rec = (off_state + total_runs * 4 + 255) & ~(size_t)255;
      const uint64_t want = (total_runs + FPC_V0_WARPS - 1) / FPC_V0_WARPS;
      if (grid > want) grid = want;
      t.slot = fpc_v0_tile_out_bytes(4, 8, 3);
      const size_t off_scr = (off_rec + total_runs * t.rec_words * 4 + 255) & ~(size_t)255;
      uint8_t* g = nullptr;
      if (!big_prepare(c, off_scr + (size_t)grid * FPC_V0_WARPS * run * t.slot, &g)) return 0;
      CK(cudaMemsetAsync(g, 0, off_rec, c->stream));
      t.scratch = g + off_scr;
      t.ticket = reinterpret_cast<uint32_t*>(g);
      t.desc = reinterpret_cast<uint64_t*>(g + off_desc);
      t.state = reinterpret_cast<uint32_t*>(g + off_state);
      t.records = reinterpret_cast<uint32_t*>(g + off_rec);
      fpc_encode_v0_tiles_kernel<uint32_t><<<(unsigned)grid, FPC_V0_WARPS * 32, smem, c->stream>>>(t);
      c->launches++;
      CK(cudaGetLastError());
      return 1;
      }
    }
  FpcLegacyEncodeArgs a;
  a.in = d_in; a.n = n; a.stride = stride; a.nstreams = nstreams; a.e1 = e1; a.e2 = e2;
  a.out = d_out; a.out_stride = out_stride; a.nbytes = d_nbytes; a.gtables = nullptr;
  size_t smem = tw * wordsize;
  if (smem > 64 * 1024)
    {
    uint8_t* g = nullptr;
    if (!big_prepare(c, tw * wordsize * nstreams, &g)) return 0;
    a.gtables = g;      // zeroed by the kernel's warp before use
    smem = 0;
    }
  if (wordsize == 4)
    {
    if (!set_smem(fpc_encode_legacy_kernel<uint32_t>, smem, c)) return 0;
    fpc_encode_legacy_kernel<uint32_t><<<nstreams, 32, smem, c->stream>>>(a);
    }
  else
    {
    if (!set_smem(fpc_encode_legacy_kernel<uint64_t>, smem, c)) return 0;
    fpc_encode_legacy_kernel<uint64_t><<<nstreams, 32, smem, c->stream>>>(a);
    }
  c->launches++;
  CK(cudaGetLastError());
  return 1;
  }

extern "C" int tb200_fpc_decode_v0(tb200_ctx* c, int wordsize, const uint8_t* d_base, const uint64_t* offsets, const uint64_t* lengths,
                                   const uint8_t* hash_info, int nstreams, uint32_t expect_n, void* d_out, uint32_t stride)
  {
  CK(cudaSetDevice(c->device));
  if (wordsize != 4 && wordsize != 8) return fail_msg("tb200_fpc_decode_v0: bad wordsize");
  if (nstreams < 1 || nstreams > 8) return fail_msg("tb200_fpc_decode_v0: bad stream count");
  size_t tw = 0;
  for (int s = 0; s < nstreams; ++s)
    {
    const size_t t = ((size_t)1 << ((hash_info[s] >> 4) << 1)) + ((size_t)1 << ((hash_info[s] & 15) << 1));
    if (t > tw) tw = t;
    }
  // pointer table, stream lengths and counts live at the start of the big scratch; tables (if global) follow
  const size_t head = 256;
  const bool global_tables = tw * wordsize > 48 * 1024;
  uint8_t* g = nullptr;
  if (!big_prepare(c, head + (global_tables ? tw * wordsize * nstreams : 0), &g)) return 0;
  struct { const uint8_t* ptrs[8]; uint64_t len[8]; } hdr;
  for (int s = 0; s < 8; ++s) { hdr.ptrs[s] = s < nstreams ? d_base + offsets[s] : nullptr; hdr.len[s] = (s < nstreams && lengths) ? lengths[s] : 0; }
  CK(cudaMemcpyAsync(g, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, c->stream));
  FpcLegacyDecodeArgs a;
  a.streams = reinterpret_cast<const uint8_t* const*>(g);
  a.extents = lengths ? reinterpret_cast<const uint64_t*>(g + 64) : nullptr;
  a.nstreams = nstreams; a.out = d_out; a.stride = stride;
  a.counts = reinterpret_cast<uint32_t*>(g + 192);
  a.expect = expect_n;
  a.gtables = nullptr; a.gtable_words = tw;
  size_t smem = FPC_LEGACY_RING + tw * wordsize;
  if (global_tables)
    {
    a.gtables = g + head;
    CK(cudaMemsetAsync(g + head, 0, tw * wordsize * nstreams, c->stream));
    smem = FPC_LEGACY_RING;
    }
  if (wordsize == 4)
    {
    if (!set_smem(fpc_decode_legacy_kernel<uint32_t>, smem, c)) return 0;
    fpc_decode_legacy_kernel<uint32_t><<<nstreams, 64, smem, c->stream>>>(a);
    }
  else
    {
    if (!set_smem(fpc_decode_legacy_kernel<uint64_t>, smem, c)) return 0;
    fpc_decode_legacy_kernel<uint64_t><<<nstreams, 64, smem, c->stream>>>(a);
    }
  c->launches++;
  CK(cudaGetLastError());
  // hdr is a stack object consumed by an async copy from pageable memory: CUDA stages pageable
  // sources before returning, so no synchronisation is needed here.
  return 1;
  }

#include "device_api_lz4.inc"
#include "device_api_misc.inc"
#include "device_api_comm.inc"
#include "device_api_stl.inc"
