/*
 * archive.c - host side of the trico C API (include/trico_b200.h), plain C.
 *
 * Owns the archive object, the wire framing of both container versions and the version dispatch;
 * every byte of codec work is done by sm_100a kernels reached through the thin C ABI of
 * include/trico_b200_device.h.  There is no CPU codec in this file and no fallback: if the device
 * layer fails, the call fails.
 *
 * Mirrors the behaviour of /root/reference/trico/trico.c (framing :12-124, open/close :126-189,
 * writers :215-858, counters :860-941, readers :943-1668, skip :1670) with one table-driven
 * writer and one table-driven reader instead of forty hand-expanded ones.
 */
#include "trico_b200.h"
#include "trico_b200_device.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define TRICO_MAGIC 0x6f637254u          /* "Trco", trico.c:94 */

static _Thread_local char g_err[256] = "";

/* TRICO_B200_TRACE=1: wall-clock time per phase of the one-shot stream path, printed at exit */
#include <time.h>
static double g_tr[8]; static unsigned long g_trn[8]; static int g_trace = -1;
static const char* const g_trname[8] = {"h2d", "encode launch", "size d2h + sync", "reserve", "stream d2h + sync", "decode upload", "decode launch", "decode d2h + sync"};
static double tr_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void tr_report(void) { for (int i = 0; i < 8; ++i) if (g_trn[i]) fprintf(stderr, "trico_b200 trace: %-20s %8lu calls %9.1f us each\n", g_trname[i], g_trn[i], 1e6 * g_tr[i] / g_trn[i]); }
static int tr_on(void) { if (g_trace < 0) { const char* e = getenv("TRICO_B200_TRACE"); g_trace = e && e[0] == '1'; if (g_trace) atexit(tr_report); } return g_trace; }
#define TR_BEGIN double tr_t0 = tr_on() ? tr_now() : 0
#define TR_MARK(i) do { if (g_trace) { const double t1 = tr_now(); g_tr[i] += t1 - tr_t0; g_trn[i]++; tr_t0 = t1; } } while (0)
static void set_err(const char* s) { snprintf(g_err, sizeof(g_err), "%s", s); }
static void set_dev_err(void) { snprintf(g_err, sizeof(g_err), "%s", tb200_last_error()); }
const char* trico_b200_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------------------------
 * Workers: a context (device + stream + kernel workspace) with two growable device buffers.
 * Archives borrow one for their lifetime; raw codec / transpose calls borrow one per call.
 * ------------------------------------------------------------------------------------------ */
#define PIPE_NBUF 3                      /* ring depth of the slab pipeline */
typedef struct
  {
  tb200_ctx* ctx;
  uint8_t* d_raw;  uint64_t raw_cap;     /* uncompressed elements */
  uint8_t* d_enc;  uint64_t enc_cap;     /* encoded stream bytes */
  uint64_t* d_scalar;                    /* 64 bytes of device scalars: [0] stream bytes, [1..3] slab totals, [7] status */
  int busy;
  /* slab pipeline (host <-> device copies overlapped with the kernels, see write_stream_pipelined) */
  int pipe_ready;
  void* s_h2d; void* s_d2h;
  void* ev_h2d[PIPE_NBUF]; void* ev_k[PIPE_NBUF]; void* ev_d2h[PIPE_NBUF];
  uint8_t* d_ring; uint64_t ring_cap;    /* 2 * PIPE_NBUF slots of ring_cap / (2 * PIPE_NBUF) bytes */
  uint8_t* d_table; uint64_t table_cap;  /* u16 chunk sizes of the stream in flight */
  uint64_t* h_scalar;                    /* page-locked, 64 bytes */
  } worker;

/* workers are heap objects in a table that grows on demand (the reference puts no bound on the
 * number of concurrently open archives); idle workers beyond KEEP_IDLE give their device memory back */
static worker** g_workers = NULL;
static int g_nworkers = 0, g_workers_cap = 0;
#define KEEP_IDLE 8
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static int env_device(void)
  {
  const char* s = getenv("TRICO_B200_DEVICE");
  if (!s) s = getenv("LOCAL_RANK");      /* one process per GPU under torchrun */
  int d = s ? atoi(s) : 0;
  int n = tb200_device_count();
  if (n > 0 && d >= n) d %= n;
  return d < 0 ? 0 : d;
  }

static worker* worker_acquire(void)
  {
  worker* w = NULL;
  pthread_mutex_lock(&g_lock);
  for (int i = 0; i < g_nworkers; ++i)
    if (!g_workers[i]->busy) { w = g_workers[i]; break; }
  if (!w)
    {
    if (g_nworkers == g_workers_cap)
      {
      const int cap = g_workers_cap ? 2 * g_workers_cap : 16;
      worker** t = (worker**)realloc(g_workers, (size_t)cap * sizeof(worker*));
      if (t) { g_workers = t; g_workers_cap = cap; }
      }
    tb200_ctx* ctx = g_nworkers < g_workers_cap ? tb200_ctx_create(env_device(), NULL) : NULL;
    if (ctx)
      {
      w = (worker*)calloc(1, sizeof(worker));
      if (w)
        {
        w->ctx = ctx;
        w->d_scalar = (uint64_t*)tb200_device_alloc(64);
        }
      if (!w || !w->d_scalar) { tb200_ctx_destroy(ctx); free(w); w = NULL; }
      else g_workers[g_nworkers++] = w;
      }
    if (!w) set_dev_err();
    }
  if (w) w->busy = 1;
  pthread_mutex_unlock(&g_lock);
  if (w) tb200_ctx_make_current(w->ctx);      /* streams, events and pinned buffers created from here on belong to its device */
  return w;
  }

static void worker_trim(worker* w)
  { /* called with the worker still marked busy: nobody else can take it meanwhile */
  tb200_ctx_make_current(w->ctx);
  tb200_ctx_sync(w->ctx);
  if (w->pipe_ready) { tb200_stream_sync(w->s_h2d); tb200_stream_sync(w->s_d2h); }
  if (w->d_raw) tb200_device_free(w->d_raw);
  if (w->d_enc) tb200_device_free(w->d_enc);
  if (w->d_ring) tb200_device_free(w->d_ring);
  if (w->d_table) tb200_device_free(w->d_table);
  w->d_raw = w->d_enc = w->d_ring = w->d_table = NULL;
  w->raw_cap = w->enc_cap = w->ring_cap = w->table_cap = 0;
  tb200_ctx_trim(w->ctx);
  }

static void worker_release(worker* w)
  {
  if (!w) return;
  pthread_mutex_lock(&g_lock);
  int idle = 0;
  for (int i = 0; i < g_nworkers; ++i) if (!g_workers[i]->busy) ++idle;
  pthread_mutex_unlock(&g_lock);
  if (idle >= KEEP_IDLE) worker_trim(w);
  pthread_mutex_lock(&g_lock);
  w->busy = 0;
  pthread_mutex_unlock(&g_lock);
  }

static int ensure(uint8_t** buf, uint64_t* cap, uint64_t need, worker* w)
  {
  if (need <= *cap) return 1;
  if (!tb200_ctx_sync(w->ctx)) { set_dev_err(); return 0; }
  if (*buf) tb200_device_free(*buf);
  *buf = NULL; *cap = 0;
  uint64_t want = need + need / 2 + 4096;         /* streams of growing size must not reallocate every time */
  if (want < (16u << 20)) want = 16u << 20;       /* ... and small ones never (cudaFree stalls every thread's stream) */
  uint8_t* p = (uint8_t*)tb200_device_alloc(want);
  if (!p) { set_dev_err(); return 0; }
  *buf = p; *cap = want;
  return 1;
  }

/* ------------------------------------------------------------------------------------------
 * Host buffers of writable archives.  The reference grows a malloc'd buffer with realloc
 * (trico.c:31-47); here the buffer is page-locked so the device can DMA the finished stream
 * straight into it, and closed archives park their buffer in a small pool (page-locking is
 * expensive, archives are usually written in a loop).  Without a CUDA device page-locking fails
 * and plain malloc is used, so an empty archive can still be created and inspected.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint8_t* p; uint64_t cap; int pinned; } hostbuf;
#define POOL_SLOTS 64
static hostbuf g_pool[POOL_SLOTS];
static uint64_t g_pool_bytes = 0;
#define POOL_MAX_BYTES (8ull << 30)

static hostbuf hostbuf_get(uint64_t need)
  {
  hostbuf b = {NULL, 0, 0};
  pthread_mutex_lock(&g_lock);
  int best = -1;
  for (int i = 0; i < POOL_SLOTS; ++i)
    if (g_pool[i].p && g_pool[i].cap >= need && (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = i;
  if (best >= 0) { b = g_pool[best]; g_pool[best].p = NULL; g_pool_bytes -= b.cap; }
  pthread_mutex_unlock(&g_lock);
  if (b.p) return b;
  const uint64_t cap = need ? need : 1;
  if (tb200_device_count() > 0 && cap >= (64u << 10))
    {
    tb200_set_device(env_device());          /* page-locking creates a context on the CURRENT device: not on GPU 0 for every rank */
    b.p = (uint8_t*)tb200_host_alloc_pinned(cap);
    b.pinned = b.p != NULL;
    }
  if (!b.p) { b.p = (uint8_t*)malloc(cap); b.pinned = 0; }
  b.cap = b.p ? cap : 0;
  return b;
  }

static void hostbuf_put(hostbuf b)
  {
  if (!b.p) return;
  if (b.pinned)
    {
    pthread_mutex_lock(&g_lock);
    int slot = -1;
    if (g_pool_bytes + b.cap <= POOL_MAX_BYTES)
      for (int i = 0; i < POOL_SLOTS; ++i) if (!g_pool[i].p) { slot = i; break; }
    if (slot >= 0) { g_pool[slot] = b; g_pool_bytes += b.cap; }
    pthread_mutex_unlock(&g_lock);
    if (slot >= 0) return;
    tb200_host_free_pinned(b.p);
    }
  else free(b.p);
  }

/* ------------------------------------------------------------------------------------------
 * archive object (trico.c:12-24)
 * ------------------------------------------------------------------------------------------ */
typedef struct
  {
  int writable;
  uint32_t version;
  /* writer: growable host buffer */
  uint8_t* buffer; uint64_t size; uint64_t cap; int buffer_pinned;
  /* reader: borrowed bytes (host or device) */
  const uint8_t* data; uint64_t data_size; uint64_t pos;
  int data_on_device;
  int next_type;
  int fpc_log2, lz4_log2;                /* 0 = default */
  int format;                            /* writer: 1 = chunked container (default), 0 = the reference's own format */
  worker* w;
  } archive;

static int buffer_reserve(archive* a, uint64_t extra)
  {
  if (!a->writable) return 0;            /* trico.c:33 */
  if (a->size + extra <= a->cap) return 1;
  uint64_t want = a->size + extra;
  if (want < a->cap * 2) want = a->cap * 2;
  hostbuf nb = hostbuf_get(want);
  if (!nb.p) return 0;                   /* trico.c:40 */
  if (a->size) memcpy(nb.p, a->buffer, a->size);
  hostbuf old = {a->buffer, a->cap, a->buffer_pinned};
  hostbuf_put(old);
  a->buffer = nb.p; a->cap = nb.cap; a->buffer_pinned = nb.pinned;
  return 1;
  }

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t get32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t get64(const uint8_t* p) { return (uint64_t)get32(p) | ((uint64_t)get32(p + 4) << 32); }

/* copies `n` archive bytes at `off` into host memory, wherever the archive lives */
static int fetch(archive* a, uint64_t off, void* dst, uint64_t n)
  {
  if (a->writable) return 0;             /* trico.c:67 */
  if (off + n > a->data_size) return 0;  /* trico.c:71 */
  if (!a->data_on_device) { memcpy(dst, a->data + off, n); return 1; }
  if (!tb200_memcpy_d2h(a->w->ctx, dst, a->data + off, n) || !tb200_ctx_sync(a->w->ctx)) { set_dev_err(); return 0; }
  return 1;
  }

static void peek_next_type(archive* a)
  { /* trico.c:100-109: the type byte of the next stream is consumed ahead of time */
  uint8_t t;
  if (a->pos < a->data_size && fetch(a, a->pos, &t, 1)) { a->next_type = t; a->pos += 1; }
  else a->next_type = trico_empty;
  }

void* trico_open_archive_for_writing(uint64_t initial_buffer_size)
  {
  archive* a = (archive*)calloc(1, sizeof(archive));
  if (!a) return NULL;
  a->writable = 1;
  hostbuf b = hostbuf_get(initial_buffer_size < 8 ? 8 : initial_buffer_size);
  if (!b.p) { free(a); return NULL; }    /* trico.c:141-145 */
  a->buffer = b.p; a->cap = b.cap; a->buffer_pinned = b.pinned;
  put32(a->buffer, TRICO_MAGIC);         /* trico.c:90-98 */
  put32(a->buffer + 4, 0);               /* version 0 until a chunked stream is appended */
  a->size = 8;
  const char* fmt = getenv("TRICO_B200_FORMAT");   /* lets UNMODIFIED callers (trico_encoder) write reference-readable archives */
  a->format = (fmt && fmt[0] == '0') ? 0 : 1;
  return a;
  }

void* trico_open_archive_for_reading(const uint8_t* data, uint64_t data_size)
  {
  archive* a = (archive*)calloc(1, sizeof(archive));
  if (!a) return NULL;
  a->data = data; a->data_size = data_size;
  a->data_on_device = tb200_pointer_is_device(data);
  if (a->data_on_device && !(a->w = worker_acquire())) { free(a); return NULL; }
  uint8_t hdr[8];
  if (!fetch(a, 0, hdr, 8) || get32(hdr) != TRICO_MAGIC)       /* trico.c:111-119 */
    { worker_release(a->w); free(a); return NULL; }
  a->version = get32(hdr + 4);
  if (a->version > 1) { set_err("unknown archive version"); worker_release(a->w); free(a); return NULL; }
  a->pos = 8;
  peek_next_type(a);
  return a;
  }

void trico_close_archive(void* h)
  {
  archive* a = (archive*)h;
  if (!a) return;
  if (a->w) { tb200_ctx_sync(a->w->ctx); worker_release(a->w); }
  hostbuf hb = {a->buffer, a->cap, a->buffer_pinned};
  hostbuf_put(hb);
  free(a);
  }

uint8_t* trico_get_buffer_pointer(void* h) { return ((archive*)h)->buffer; }
uint64_t trico_get_size(void* h) { return ((archive*)h)->size; }
uint32_t trico_get_version(void* h) { return ((archive*)h)->version; }
enum trico_stream_type trico_get_next_stream_type(void* h) { return (enum trico_stream_type)((archive*)h)->next_type; }

int trico_b200_set_chunking(void* h, int fpc_log2_values, int lz4_log2_bytes)
  {
  archive* a = (archive*)h;
  if (!a || !a->writable) return 0;
  if (fpc_log2_values && (fpc_log2_values < 5 || fpc_log2_values > 12)) return 0;
  if (lz4_log2_bytes && (lz4_log2_bytes < 8 || lz4_log2_bytes > 15)) return 0;
  a->fpc_log2 = fpc_log2_values; a->lz4_log2 = lz4_log2_bytes;
  return 1;
  }

/* 0: every stream of this archive is written in the reference's own layout (one FPC stream per
 * component with (4,10) / (20,20) tables, one LZ4 block per byte plane: trico.c:215-262, :323-378), so
 * an unmodified reference decoder reads it; 1 (default): the chunked container.  Must be chosen
 * before the first stream is written. */
int trico_b200_set_format(void* h, int version)
  {
  archive* a = (archive*)h;
  if (!a || !a->writable || (version != 0 && version != 1)) return 0;
  if (a->size > 8 && version != a->format) return 0;
  a->format = version;
  return 1;
  }

uint64_t trico_b200_launch_count(void* h)
  {
  archive* a = (archive*)h;
  return (a && a->w) ? tb200_ctx_launch_count(a->w->ctx) : 0;
  }

static int need_worker(archive* a)
  {
  if (!a->w) a->w = worker_acquire();
  return a->w != NULL;
  }

/* ------------------------------------------------------------------------------------------
 * Slab pipeline.  A host-resident stream larger than a few slabs is cut into chunk-aligned slabs;
 * slab i+1 crosses PCIe while slab i is in the kernels and slab i-1 travels back, on three CUDA
 * streams chained by events.  Chunks are independent, so the bytes are identical to the one-shot
 * path: only the schedule differs.  (The reference has no counterpart: it is one CPU thread.)
 * ------------------------------------------------------------------------------------------ */
static uint64_t slab_bytes(void)
  {
  const char* s = getenv("TRICO_B200_SLAB_MB");
  long mb = s ? atol(s) : 0;
  if (mb < 1 || mb > 1024) mb = 32;
  return (uint64_t)mb << 20;
  }

static int pipe_init(worker* w)
  {
  if (w->pipe_ready) return 1;
  w->s_h2d = tb200_stream_create();
  w->s_d2h = tb200_stream_create();
  w->h_scalar = (uint64_t*)tb200_host_alloc_pinned(64);
  int ok = w->s_h2d && w->s_d2h && w->h_scalar;
  for (int i = 0; i < PIPE_NBUF; ++i)
    {
    w->ev_h2d[i] = tb200_event_create_notiming();
    w->ev_k[i] = tb200_event_create_notiming();
    w->ev_d2h[i] = tb200_event_create_notiming();
    ok = ok && w->ev_h2d[i] && w->ev_k[i] && w->ev_d2h[i];
    }
  if (!ok) { set_dev_err(); return 0; }
  w->pipe_ready = 1;
  return 1;
  }

/* waits until nothing of the pipeline is in flight (also the error path: buffers are reused) */
static int pipe_drain(worker* w)
  {
  int ok = tb200_stream_sync(w->s_h2d);
  ok = tb200_ctx_sync(w->ctx) && ok;
  ok = tb200_stream_sync(w->s_d2h) && ok;
  return ok;
  }

static int pipe_buffers(worker* w, uint64_t slot_bytes, uint64_t table_bytes)
  {
  const uint64_t need = slot_bytes * 2 * PIPE_NBUF;
  if (need > w->ring_cap)
    {
    if (!pipe_drain(w)) { set_dev_err(); return 0; }
    if (w->d_ring) tb200_device_free(w->d_ring);
    w->d_ring = (uint8_t*)tb200_device_alloc(need);
    w->ring_cap = w->d_ring ? need : 0;
    if (!w->d_ring) { set_dev_err(); return 0; }
    }
  if (!ensure(&w->d_table, &w->table_cap, table_bytes + 64, w)) return 0;
  return 1;
  }

typedef struct
  {
  int codec, ws, nc, log2c;
  uint64_t n;               /* scalars per component (FPC) / elements (LZ4) */
  uint64_t unit;            /* scalars (elements) per chunk range */
  uint64_t slab_ranges;     /* ranges per slab */
  uint64_t nranges, nslabs;
  uint64_t range_raw;       /* raw bytes of one full range (all components / the whole element) */
  uint32_t nsub;            /* chunks per range */
  uint64_t chunk_bound;     /* worst-case bytes of one chunk */
  } slab_plan;

static int plan_slabs(slab_plan* p, int type, uint32_t count, int log2c)
  {
  int ws = 0, nc = 0, pc = 0;
  p->codec = tb200_stream_layout(type, &ws, &nc, &pc);
  if (!p->codec) return 0;
  p->ws = ws; p->nc = nc; p->log2c = log2c;
  p->n = (uint64_t)count * pc;
  p->unit = (uint64_t)1 << log2c;
  p->nranges = (p->n + p->unit - 1) >> log2c;
  p->nsub = (uint32_t)(p->codec == 1 ? nc : ws);
  p->range_raw = p->unit * (uint64_t)ws * (p->codec == 1 ? nc : 1);
  uint64_t r = slab_bytes() / p->range_raw;
  /* whole encode / decode tiles (12 resp. up to 128 ranges): 384 = lcm */
  if (p->codec == 1) r = r / 384 * 384;
  if (r < 384 && p->codec == 1) r = 384;
  if (r < 1) r = 1;
  p->slab_ranges = r;
  p->nslabs = (p->nranges + r - 1) / r;
  /* worst-case bytes of one chunk, recovered from the public whole-stream bound of one range */
  const uint32_t one_range_count = (uint32_t)((p->unit + pc - 1) / pc);
  const uint64_t b1 = tb200_v1_stream_bound(type, one_range_count, log2c);
  const uint64_t nchb = tb200_v1_nchunks(type, one_range_count, log2c);
  p->chunk_bound = (b1 - TB200_V1_FIXED_BYTES - 64 - 2 * nchb) / (nchb ? nchb : 1);
  return 1;
  }

static int use_pipeline(const slab_plan* p, uint64_t raw_bytes)
  {
  const char* s = getenv("TRICO_B200_NO_PIPELINE");
  if (s && s[0] == '1') return 0;
  return p->nslabs >= 3 && raw_bytes >= (1u << 20);
  }

static int write_stream_pipelined(archive* a, int type, const uint8_t* data, uint32_t count, const slab_plan* p)
  {
  worker* w = a->w;
  if (!pipe_init(w)) return 0;
  const uint64_t nch = p->nranges * p->nsub;
  const uint64_t slab_raw = p->slab_ranges * p->range_raw;
  const uint64_t slab_enc = p->slab_ranges * p->nsub * p->chunk_bound + 256;
  const uint64_t slot = ((slab_raw > slab_enc ? slab_raw : slab_enc) + 4096 + 255) & ~(uint64_t)255;
  if (!pipe_buffers(w, slot, 2 * nch)) return 0;
  /* nothing of an earlier call may still be using the ring or the scalars */
  if (!pipe_drain(w)) { set_dev_err(); return 0; }
  const uint64_t head = TB200_V1_FIXED_BYTES + 2 * nch;
  if (!buffer_reserve(a, head)) return 0;
  const uint64_t stream_off = a->size;
  uint64_t pay = 0;                                    /* payload bytes placed so far */
  void* cs = tb200_ctx_stream(w->ctx);
  int ok = 1;
  const int info = p->codec == 1 ? (((TB200_V1_E1 >> 1) << 4) | (TB200_V1_E2 >> 1)) : 0;

  for (uint64_t s = 0; ok && s <= p->nslabs; ++s)
    {
    if (s < p->nslabs)
      {
      const int k = (int)(s % PIPE_NBUF);
      const uint64_t r0 = s * p->slab_ranges;
      uint64_t r1 = r0 + p->slab_ranges; if (r1 > p->nranges) r1 = p->nranges;
      const uint64_t e0 = r0 << p->log2c;
      uint64_t e1 = r1 << p->log2c; if (e1 > p->n) e1 = p->n;
      const uint64_t esz = (uint64_t)p->ws * (p->codec == 1 ? p->nc : 1);
      uint8_t* d_in = w->d_ring + (uint64_t)k * slot;
      uint8_t* d_out = w->d_ring + (uint64_t)(PIPE_NBUF + k) * slot;
      if (s >= PIPE_NBUF) ok = ok && tb200_stream_wait_event(w->s_h2d, w->ev_k[k]);
      ok = ok && tb200_memcpy_h2d_on(w->s_h2d, d_in, data + e0 * esz, (e1 - e0) * esz);
      ok = ok && tb200_event_record_on(w->s_h2d, w->ev_h2d[k]);
      ok = ok && tb200_stream_wait_event(cs, w->ev_h2d[k]);
      if (s >= PIPE_NBUF) ok = ok && tb200_stream_wait_event(cs, w->ev_d2h[k]);
      uint8_t* d_sizes = w->d_table + 2 * r0 * p->nsub;
      if (p->codec == 1)
        ok = ok && tb200_fpc_encode(w->ctx, p->ws, p->nc, d_in, e1 - e0, p->log2c, TB200_V1_E1, TB200_V1_E2, d_sizes, d_out, NULL, w->d_scalar + 1 + k);
      else
        ok = ok && tb200_lz4_encode(w->ctx, p->ws, d_in, e1 - e0, p->log2c, d_sizes, d_out, NULL, w->d_scalar + 1 + k);
      ok = ok && tb200_memcpy_d2h(w->ctx, w->h_scalar + 1 + k, w->d_scalar + 1 + k, 8);
      ok = ok && tb200_event_record_on(cs, w->ev_k[k]);
      if (!ok) set_dev_err();
      }
    if (ok && s >= 1)
      { /* slab s-1 has been encoded: its size is known, send it home */
      const int k = (int)((s - 1) % PIPE_NBUF);
      ok = tb200_event_sync(w->ev_k[k]);
      if (!ok) { set_dev_err(); break; }
      const uint64_t tot = w->h_scalar[1 + k];
      if (tot > slot) { set_err("encoder returned an impossible size"); ok = 0; break; }
      if (stream_off + head + pay + tot > a->cap)
        { /* growing moves the buffer: copies in flight must land first */
        ok = tb200_stream_sync(w->s_d2h);
        if (!ok) { set_dev_err(); break; }
        a->size = stream_off + head + pay;
        if (!buffer_reserve(a, tot)) { a->size = stream_off; ok = 0; break; }
        a->size = stream_off;
        }
      ok = tb200_memcpy_d2h_on(w->s_d2h, a->buffer + stream_off + head + pay, w->d_ring + (uint64_t)(PIPE_NBUF + k) * slot, tot) &&
           tb200_event_record_on(w->s_d2h, w->ev_d2h[k]);
      if (!ok) set_dev_err();
      pay += tot;
      }
    }
  /* size table, then the fixed header from the host */
  ok = ok && tb200_memcpy_d2h_on(w->s_d2h, a->buffer + stream_off + TB200_V1_FIXED_BYTES, w->d_table, 2 * nch);
  if (!pipe_drain(w)) ok = 0;
  if (!ok) { if (g_err[0] == 0) set_dev_err(); return 0; }
  uint8_t* h = a->buffer + stream_off;
  h[0] = (uint8_t)type; put32(h + 1, count); h[5] = (uint8_t)info; h[6] = (uint8_t)p->log2c;
  put32(h + 7, (uint32_t)pay); put32(h + 11, (uint32_t)(pay >> 32));
  a->size = stream_off + head + pay;
  if (a->version == 0) { a->version = 1; put32(a->buffer + 4, 1); }
  return 1;
  }

/* v1 stream at a->data + start (host memory) -> host_dst (host memory) */
static int read_stream_pipelined(archive* a, uint64_t start, const uint8_t* head, void* host_dst, const slab_plan* p)
  {
  worker* w = a->w;
  if (!pipe_init(w)) return 0;
  const uint64_t nch = p->nranges * p->nsub;
  const uint64_t total = get64(head + 7);
  const int info = head[5];
  const uint8_t* table = a->data + start + TB200_V1_FIXED_BYTES;
  const uint8_t* payload = table + 2 * nch;
  const uint64_t slab_raw = p->slab_ranges * p->range_raw;
  const uint64_t slab_enc_max = p->slab_ranges * p->nsub * 65535ull;      /* what the u16 table can describe */
  /* payload offset of every slab: a host walk over the u16 sizes */
  uint64_t* poff = (uint64_t*)malloc((p->nslabs + 1) * sizeof(uint64_t));
  if (!poff) return 0;
  uint64_t acc = 0, worst = 0;
  for (uint64_t s = 0; s < p->nslabs; ++s)
    {
    poff[s] = acc;
    const uint64_t c0 = s * p->slab_ranges * p->nsub;
    uint64_t c1 = c0 + p->slab_ranges * p->nsub; if (c1 > nch) c1 = nch;
    const uint8_t* t = table + 2 * c0;
    uint64_t sum = 0;
    for (uint64_t c = 0; c < c1 - c0; ++c) sum += (uint64_t)t[2 * c] | ((uint64_t)t[2 * c + 1] << 8);
    acc += sum;
    if (sum > worst) worst = sum;
    }
  poff[p->nslabs] = acc;
  if (acc != total || worst > slab_enc_max) { free(poff); set_err("chunk size table disagrees with the stream header"); return 0; }
  const uint64_t slot = ((slab_raw > worst ? slab_raw : worst) + 4096 + 255) & ~(uint64_t)255;
  if (!pipe_buffers(w, slot, 2 * nch) || !pipe_drain(w)) { free(poff); return 0; }
  void* cs = tb200_ctx_stream(w->ctx);
  uint32_t* d_status = (uint32_t*)(w->d_scalar + 7);
  int ok = tb200_memset_d(w->ctx, d_status, 0, 8);
  ok = ok && tb200_memcpy_h2d(w->ctx, w->d_table, table, 2 * nch);
  const uint64_t esz = (uint64_t)p->ws * (p->codec == 1 ? p->nc : 1);
  for (uint64_t s = 0; ok && s < p->nslabs; ++s)
    {
    const int k = (int)(s % PIPE_NBUF);
    const uint64_t r0 = s * p->slab_ranges;
    uint64_t r1 = r0 + p->slab_ranges; if (r1 > p->nranges) r1 = p->nranges;
    const uint64_t e0 = r0 << p->log2c;
    uint64_t e1 = r1 << p->log2c; if (e1 > p->n) e1 = p->n;
    const uint64_t plen = poff[s + 1] - poff[s];
    /* keep the payload's alignment modulo 16 and leave room in front: the kernels read whole words */
    uint8_t* d_in = w->d_ring + (uint64_t)k * slot + 64 + ((uintptr_t)(payload + poff[s]) & 15u);
    uint8_t* d_out = w->d_ring + (uint64_t)(PIPE_NBUF + k) * slot;
    if (s >= PIPE_NBUF) ok = ok && tb200_stream_wait_event(w->s_h2d, w->ev_k[k]);
    ok = ok && tb200_memcpy_h2d_on(w->s_h2d, d_in, payload + poff[s], plen);
    ok = ok && tb200_event_record_on(w->s_h2d, w->ev_h2d[k]);
    ok = ok && tb200_stream_wait_event(cs, w->ev_h2d[k]);
    if (s >= PIPE_NBUF) ok = ok && tb200_stream_wait_event(cs, w->ev_d2h[k]);
    const uint8_t* d_sizes = w->d_table + 2 * r0 * p->nsub;
    if (p->codec == 1)
      ok = ok && tb200_fpc_decode(w->ctx, p->ws, p->nc, d_sizes, d_in, plen, e1 - e0, p->log2c, (info >> 4) << 1, (info & 15) << 1, d_out);
    else
      ok = ok && tb200_lz4_decode_async(w->ctx, p->ws, d_sizes, d_in, plen, e1 - e0, p->log2c, d_out, d_status);
    ok = ok && tb200_event_record_on(cs, w->ev_k[k]);
    ok = ok && tb200_stream_wait_event(w->s_d2h, w->ev_k[k]);
    ok = ok && tb200_memcpy_d2h_on(w->s_d2h, (uint8_t*)host_dst + e0 * esz, d_out, (e1 - e0) * esz);
    ok = ok && tb200_event_record_on(w->s_d2h, w->ev_d2h[k]);
    }
  free(poff);
  if (!ok) set_dev_err();
  ok = ok && tb200_memcpy_d2h(w->ctx, w->h_scalar + 7, d_status, 8);
  if (!pipe_drain(w)) { if (ok) set_dev_err(); ok = 0; }
  if (ok && (uint32_t)w->h_scalar[7] != 0) { set_err("malformed LZ4 block"); ok = 0; }
  return ok;
  }

/* ------------------------------------------------------------------------------------------
 * writer: one routine for all stream types (trico.c:215-858).  `count` is the value stored in
 * the stream header.
 * ------------------------------------------------------------------------------------------ */
/* Reference-format stream (trico.c:215-262 float/double components, :323-378 byte planes):
 *   u8 type, u32 count, then per component / plane: u32 nbytes + payload.
 * FPC components come from the tile-parallel v0 encoder (byte-identical to trico_compress), the
 * planes from tb200_lz4_encode_v0 (one valid LZ4 block per plane). */
static int write_stream_v0(archive* a, int type, const void* data, uint32_t count, int codec, int ws, int nc, int pc)
  {
  worker* w = a->w;
  const uint64_t n = (uint64_t)count * pc;                    /* values per component / bytes per plane */
  const int nsub = codec == 1 ? nc : ws;
  const uint64_t raw_bytes = n * ws * (codec == 1 ? nc : 1);
  if (n > 0x7E000000ull) { set_err("stream too long for the reference format"); return 0; }
  if (!buffer_reserve(a, 5)) return 0;
  a->buffer[a->size] = (uint8_t)type;
  put32(a->buffer + a->size + 1, count);
  a->size += 5;
  if (n == 0)
    { /* what the reference's codecs leave of an empty input: an FPC header + one group of pad slots
         (fpc.c:196-204), an LZ4 block that is one zero token */
    for (int s = 0; s < nsub; ++s)
      {
      if (codec == 1)
        {
        uint32_t nb = 0; uint8_t* p = NULL;
        if (ws == 4) trico_compress(&nb, &p, NULL, 0, 4, 10); else trico_compress_double_precision(&nb, &p, NULL, 0, 20, 20);
        if (!p || !buffer_reserve(a, 4 + (uint64_t)nb)) { free(p); return 0; }
        put32(a->buffer + a->size, nb); memcpy(a->buffer + a->size + 4, p, nb); a->size += 4 + (uint64_t)nb;
        free(p);
        }
      else
        {
        if (!buffer_reserve(a, 5)) return 0;
        put32(a->buffer + a->size, 1); a->buffer[a->size + 4] = 0; a->size += 5;
        }
      }
    return 1;
    }
  const void* d_in = data;
  if (!tb200_pointer_is_device(data))
    {
    if (!ensure(&w->d_raw, &w->raw_cap, raw_bytes + 64, w)) return 0;
    if (!tb200_memcpy_h2d(w->ctx, w->d_raw, data, raw_bytes)) { set_dev_err(); return 0; }
    d_in = w->d_raw;
    }
  const uint64_t per = ((codec == 1 ? tb200_fpc_v0_bound(ws, (uint32_t)n) : tb200_lz4_v0_bound(n)) + 255) & ~(uint64_t)255;
  if (!ensure(&w->d_enc, &w->enc_cap, 256 + per * nsub, w)) return 0;
  uint8_t* d_sizes = w->d_enc;                                /* u32 (FPC) / u64 (LZ4) per sub-stream */
  uint8_t* d_pay = w->d_enc + 256;
  int ok;
  if (codec == 1) ok = tb200_fpc_encode_v0(w->ctx, ws, d_in, (uint32_t)n, (uint32_t)nc, nc, ws == 4 ? 4 : 20, ws == 4 ? 10 : 20, d_pay, per, (uint32_t*)d_sizes);
  else            ok = tb200_lz4_encode_v0(w->ctx, ws, d_in, n, d_pay, per, (uint64_t*)d_sizes);
  if (!ok) { set_dev_err(); return 0; }
  uint64_t sizes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (!tb200_memcpy_d2h(w->ctx, sizes, d_sizes, 64) || !tb200_ctx_sync(w->ctx)) { set_dev_err(); return 0; }
  uint64_t total = 0;
  uint64_t nb[8];
  for (int s = 0; s < nsub; ++s)
    {
    nb[s] = codec == 1 ? ((const uint32_t*)sizes)[s] : sizes[s];
    if (nb[s] == 0 || nb[s] > per || nb[s] > 0xffffffffull) { set_err("encoder returned an impossible size"); return 0; }
    total += 4 + nb[s];
    }
  if (!buffer_reserve(a, total)) return 0;
  for (int s = 0; s < nsub; ++s)
    {
    put32(a->buffer + a->size, (uint32_t)nb[s]);
    if (!tb200_memcpy_d2h(w->ctx, a->buffer + a->size + 4, d_pay + per * s, nb[s])) { set_dev_err(); return 0; }
    a->size += 4 + nb[s];
    }
  if (!tb200_ctx_sync(w->ctx)) { set_dev_err(); return 0; }
  return 1;
  }

static int write_stream(void* h, int type, const void* data, uint32_t count)
  {
  archive* a = (archive*)h;
  if (!a || !a->writable) return 0;
  int ws = 0, nc = 0, pc = 0;
  const int codec = tb200_stream_layout(type, &ws, &nc, &pc);
  if (!codec) return 0;
  if (!need_worker(a)) return 0;
  worker* w = a->w;
  if (a->format == 0) return write_stream_v0(a, type, data, count, codec, ws, nc, pc);
  int log2c = codec == 1 ? a->fpc_log2 : a->lz4_log2;
  if (codec == 1 && log2c && ws == 8 && log2c > 11) log2c = 11;
  if (codec == 2 && log2c && ws == 8 && log2c > 14) log2c = 14;
  if (!log2c) log2c = tb200_default_log2_chunk(type, count);
  const uint64_t nscalars = (uint64_t)count * pc * (codec == 1 ? nc : 1);
  const uint64_t raw_bytes = nscalars * ws;
  const uint64_t bound = tb200_v1_stream_bound(type, count, log2c);
  const void* d_in = data;
  const int data_on_device = raw_bytes ? tb200_pointer_is_device(data) : 0;
  if (raw_bytes && !data_on_device)
    {
    slab_plan plan;
    if (plan_slabs(&plan, type, count, log2c) && use_pipeline(&plan, raw_bytes))
      return write_stream_pipelined(a, type, (const uint8_t*)data, count, &plan);
    }
  TR_BEGIN;
  if (raw_bytes && !data_on_device)
    {
    if (!ensure(&w->d_raw, &w->raw_cap, raw_bytes + 64, w)) return 0;
    if (!tb200_memcpy_h2d(w->ctx, w->d_raw, data, raw_bytes)) { set_dev_err(); return 0; }
    d_in = w->d_raw;
    }
  TR_MARK(0);
  if (!ensure(&w->d_enc, &w->enc_cap, bound, w)) return 0;
  if (!tb200_encode_stream(w->ctx, type, d_in, count, log2c, w->d_enc, w->enc_cap, w->d_scalar)) { set_dev_err(); return 0; }
  TR_MARK(1);
  uint64_t stream_bytes = 0;
  if (!tb200_memcpy_d2h(w->ctx, &stream_bytes, w->d_scalar, 8) || !tb200_ctx_sync(w->ctx)) { set_dev_err(); return 0; }
  TR_MARK(2);
  if (stream_bytes < TB200_V1_FIXED_BYTES || stream_bytes > bound) { set_err("encoder returned an impossible size"); return 0; }
  if (!buffer_reserve(a, stream_bytes)) return 0;
  TR_MARK(3);
  if (!tb200_memcpy_d2h(w->ctx, a->buffer + a->size, w->d_enc, stream_bytes) || !tb200_ctx_sync(w->ctx)) { set_dev_err(); return 0; }
  TR_MARK(4);
  a->size += stream_bytes;
  if (a->version == 0) { a->version = 1; put32(a->buffer + 4, 1); }
  return 1;
  }

int trico_write_vertices(void* a, const float* p, uint32_t n) { return write_stream(a, trico_vertex_float_stream, p, n); }
int trico_write_vertices_double(void* a, const double* p, uint32_t n) { return write_stream(a, trico_vertex_double_stream, p, n); }
int trico_write_vertex_normals(void* a, const float* p, uint32_t n) { return write_stream(a, trico_vertex_normal_float_stream, p, n); }
int trico_write_vertex_normals_double(void* a, const double* p, uint32_t n) { return write_stream(a, trico_vertex_normal_double_stream, p, n); }
int trico_write_triangle_normals(void* a, const float* p, uint32_t n) { return write_stream(a, trico_triangle_normal_float_stream, p, n); }
int trico_write_triangle_normals_double(void* a, const double* p, uint32_t n) { return write_stream(a, trico_triangle_normal_double_stream, p, n); }
int trico_write_triangles(void* a, const uint32_t* p, uint32_t n) { return write_stream(a, trico_triangle_uint32_stream, p, n); }
int trico_write_triangles_long(void* a, const uint64_t* p, uint32_t n) { return write_stream(a, trico_triangle_uint64_stream, p, n); }
int trico_write_uv_per_vertex(void* a, const float* p, uint32_t n) { return write_stream(a, trico_uv_per_vertex_float_stream, p, n); }
/* the float per-triangle writer stores three uv positions per triangle (trico.c:579) */
int trico_write_uv_per_triangle(void* a, const float* p, uint32_t n) { return write_stream(a, trico_uv_per_triangle_float_stream, p, n * 3u); }
/* double uv: correct tags 6/8 (the reference writes 5/7, trico.c:622,:627); count as given (trico.c:627) */
int trico_write_uv_per_vertex_double(void* a, const double* p, uint32_t n) { return write_stream(a, trico_uv_per_vertex_double_stream, p, n); }
int trico_write_uv_per_triangle_double(void* a, const double* p, uint32_t n) { return write_stream(a, trico_uv_per_triangle_double_stream, p, n); }
int trico_write_vertex_colors(void* a, const uint32_t* p, uint32_t n) { return write_stream(a, trico_vertex_color_stream, p, n); }
int trico_write_triangle_colors(void* a, const uint32_t* p, uint32_t n) { return write_stream(a, trico_triangle_color_stream, p, n); }
int trico_write_attributes_float(void* a, const float* p, uint32_t n) { return write_stream(a, trico_attribute_float_stream, p, n); }
int trico_write_attributes_double(void* a, const double* p, uint32_t n) { return write_stream(a, trico_attribute_double_stream, p, n); }
int trico_write_attributes_uint8(void* a, const uint8_t* p, uint32_t n) { return write_stream(a, trico_attribute_uint8_stream, p, n); }
int trico_write_attributes_uint16(void* a, const uint16_t* p, uint32_t n) { return write_stream(a, trico_attribute_uint16_stream, p, n); }
int trico_write_attributes_uint32(void* a, const uint32_t* p, uint32_t n) { return write_stream(a, trico_attribute_uint32_stream, p, n); }
int trico_write_attributes_uint64(void* a, const uint64_t* p, uint32_t n) { return write_stream(a, trico_attribute_uint64_stream, p, n); }

/* ------------------------------------------------------------------------------------------
 * counters (trico.c:860-941): the count is peeked, not consumed (read_inplace, trico.c:78)
 * ------------------------------------------------------------------------------------------ */
static uint32_t peek_count(void* h, unsigned long long typemask)
  {
  archive* a = (archive*)h;
  if (!a || a->writable) return 0;
  if (a->next_type < 1 || a->next_type > 20) return 0;       /* the type byte is untrusted: no shift by it */
  if (!((typemask >> a->next_type) & 1ull)) return 0;
  uint8_t b[4];
  if (!fetch(a, a->pos, b, 4)) return 0;
  return get32(b);
  }
#define BIT(t) (1ull << (t))
uint32_t trico_get_number_of_vertices(void* a) { return peek_count(a, BIT(1) | BIT(2)); }
uint32_t trico_get_number_of_triangles(void* a) { return peek_count(a, BIT(3) | BIT(4)); }
uint32_t trico_get_number_of_uvs(void* a) { return peek_count(a, BIT(5) | BIT(6) | BIT(7) | BIT(8)); }
uint32_t trico_get_number_of_normals(void* a) { return peek_count(a, BIT(9) | BIT(10) | BIT(11) | BIT(12)); }
uint32_t trico_get_number_of_colors(void* a) { return peek_count(a, BIT(13) | BIT(14)); }
uint32_t trico_get_number_of_attributes(void* a) { return peek_count(a, BIT(15) | BIT(16) | BIT(17) | BIT(18) | BIT(19) | BIT(20)); }

/* ------------------------------------------------------------------------------------------
 * reader: one routine for all stream types and both container versions (trico.c:943-1668).
 *   out == NULL            skip the stream
 *   alloc_result != 0      malloc the result and store it in *out (float/double attribute lists,
 *                          trico.c:1377, :1408); otherwise *out is the caller's buffer.
 * a->pos points just past the (already consumed) type byte.
 * ------------------------------------------------------------------------------------------ */
static int upload(archive* a, uint64_t off, uint64_t n, const uint8_t** d_ptr)
  {
  worker* w = a->w;
  if (a->data_on_device) { *d_ptr = a->data + off; return 1; }
  if (!ensure(&w->d_enc, &w->enc_cap, n + 256, w)) return 0;
  if (!tb200_memcpy_h2d(w->ctx, w->d_enc, a->data + off, n)) { set_dev_err(); return 0; }
  *d_ptr = w->d_enc;
  return 1;
  }

static int read_stream(void* h, int type, void** out, int alloc_result)
  {
  archive* a = (archive*)h;
  if (!a || a->writable) return 0;
  if (a->next_type != type) return 0;                        /* trico.c:946 */
  int ws = 0, nc = 0, pc = 0;
  const int codec = tb200_stream_layout(type, &ws, &nc, &pc);
  if (!codec) return 0;
  const uint64_t start = a->pos - 1;                          /* the type byte */
  uint8_t head[TB200_V1_FIXED_BYTES];
  if (!fetch(a, a->pos, head + 1, 4)) return 0;
  const uint32_t count = get32(head + 1);
  const uint64_t n = (uint64_t)count * pc;                    /* scalars per component / plane */
  const uint64_t nscalars = n * (codec == 1 ? nc : 1);
  const uint64_t raw_bytes = nscalars * ws;
  const int nsub = codec == 1 ? nc : ws;
  uint64_t end;                                               /* first byte after the stream */
  uint64_t sub_off[8]; uint32_t sub_len[8]; uint8_t sub_info[8];

  if (a->version == 0)
    { /* u32 nbytes + payload per component / plane (trico.c:953-979, :1094-1128) */
    uint64_t p = a->pos + 4;
    for (int s = 0; s < nsub; ++s)
      {
      uint8_t b[5];
      if (!fetch(a, p, b, 4)) return 0;
      sub_len[s] = get32(b);
      sub_off[s] = p + 4 - start;
      if (p + 4 + (uint64_t)sub_len[s] > a->data_size) return 0;
      sub_info[s] = 0;
      if (codec == 1)
        {
        if (sub_len[s] < 5 || !fetch(a, p + 4, b, 5)) return 0;
        sub_info[s] = b[0];
        const uint32_t stream_n = ((uint32_t)b[1] << 24) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 8) | b[4];
        if (stream_n != n) { set_err("component stream length disagrees with the stream header"); return 0; }   /* trico.c:982 */
        }
      p += 4 + (uint64_t)sub_len[s];
      }
    end = p;
    }
  else
    {
    head[0] = (uint8_t)type;
    if (!fetch(a, a->pos + 4, head + 5, TB200_V1_FIXED_BYTES - 5)) return 0;
    /* untrusted header fields: the chunk size first (it is a shift count), then the table and the
     * payload against the bytes that are really there, in subtraction form so nothing wraps */
    if (head[6] < 5 || head[6] > 15) { set_err("bad chunk size in stream header"); return 0; }
    const uint64_t nch = tb200_v1_nchunks(type, count, head[6]);
    const uint64_t total = get64(head + 7);
    const uint64_t left = a->data_size - start;                 /* >= TB200_V1_FIXED_BYTES: the header was fetched */
    if (nch > (left - TB200_V1_FIXED_BYTES) / 2) return 0;     /* trico.c:71 */
    if (total > left - TB200_V1_FIXED_BYTES - 2 * nch) return 0;
    end = start + TB200_V1_FIXED_BYTES + 2 * nch + total;
    }

  if (out != NULL && raw_bytes > 0)
    {
    if (!need_worker(a)) return 0;
    worker* w = a->w;
    void* host_dst = NULL;
    void* d_dst;
    if (alloc_result)
      {
      host_dst = malloc(raw_bytes);
      if (!host_dst) return 0;
      }
    else host_dst = *out;
    const int dst_on_device = !alloc_result && tb200_pointer_is_device(host_dst);
    slab_plan plan;
    if (a->version == 1 && !dst_on_device && !a->data_on_device && head[6] >= 5 && head[6] <= 15 &&
        plan_slabs(&plan, type, count, head[6]) && use_pipeline(&plan, raw_bytes))
      {
      if (!read_stream_pipelined(a, start, head, host_dst, &plan)) { if (alloc_result) free(host_dst); return 0; }
      if (alloc_result) *out = host_dst;
      a->pos = end;
      peek_next_type(a);
      return 1;
      }
    if (dst_on_device) d_dst = host_dst;
    else
      {
      if (!ensure(&w->d_raw, &w->raw_cap, raw_bytes + 64, w)) { if (alloc_result) free(host_dst); return 0; }
      d_dst = w->d_raw;
      }
    const uint8_t* d_stream = NULL;
    int ok = upload(a, start, end - start, &d_stream);
    if (ok)
      {
      if (a->version == 1)
        ok = tb200_decode_stream(w->ctx, head, d_stream, end - start, d_dst);
      else if (codec == 1)
        {
        uint64_t sub_room[8];                /* bytes of the uploaded stream that follow each component's first byte */
        for (int s = 0; s < nsub; ++s) sub_room[s] = (end - start) - sub_off[s];
        ok = tb200_fpc_decode_v0(w->ctx, ws, d_stream, sub_off, sub_room, sub_info, nsub, (uint32_t)n, d_dst, (uint32_t)nc);
        }
      else
        ok = tb200_lz4_decode_v0(w->ctx, nsub, d_stream, sub_off, sub_len, n, d_dst);
      if (!ok) set_dev_err();
      }
    if (ok && !dst_on_device) ok = tb200_memcpy_d2h(w->ctx, host_dst, d_dst, raw_bytes);
    if (ok) ok = tb200_ctx_sync(w->ctx);
    if (!ok) { if (g_err[0] == 0) set_dev_err(); if (alloc_result) free(host_dst); return 0; }
    if (alloc_result) *out = host_dst;
    }
  else if (out != NULL && alloc_result)
    *out = malloc(1);

  a->pos = end;
  peek_next_type(a);
  return 1;
  }

#define RD(a, t, p) read_stream((a), (t), (void**)(p), 0)
int trico_read_vertices(void* a, float** p) { return RD(a, trico_vertex_float_stream, p); }
int trico_read_vertices_double(void* a, double** p) { return RD(a, trico_vertex_double_stream, p); }
int trico_read_vertex_normals(void* a, float** p) { return RD(a, trico_vertex_normal_float_stream, p); }
int trico_read_vertex_normals_double(void* a, double** p) { return RD(a, trico_vertex_normal_double_stream, p); }
int trico_read_triangle_normals(void* a, float** p) { return RD(a, trico_triangle_normal_float_stream, p); }
int trico_read_triangle_normals_double(void* a, double** p) { return RD(a, trico_triangle_normal_double_stream, p); }
int trico_read_triangles(void* a, uint32_t** p) { return RD(a, trico_triangle_uint32_stream, p); }
int trico_read_triangles_long(void* a, uint64_t** p) { return RD(a, trico_triangle_uint64_stream, p); }
int trico_read_uv_per_vertex(void* a, float** p) { return RD(a, trico_uv_per_vertex_float_stream, p); }
int trico_read_uv_per_vertex_double(void* a, double** p) { return RD(a, trico_uv_per_vertex_double_stream, p); }
int trico_read_uv_per_triangle(void* a, float** p) { return RD(a, trico_uv_per_triangle_float_stream, p); }
int trico_read_uv_per_triangle_double(void* a, double** p) { return RD(a, trico_uv_per_triangle_double_stream, p); }
int trico_read_vertex_colors(void* a, uint32_t** p) { return RD(a, trico_vertex_color_stream, p); }
int trico_read_triangle_colors(void* a, uint32_t** p) { return RD(a, trico_triangle_color_stream, p); }
int trico_read_attributes_float(void* a, float** p) { return read_stream(a, trico_attribute_float_stream, (void**)p, 1); }
int trico_read_attributes_double(void* a, double** p) { return read_stream(a, trico_attribute_double_stream, (void**)p, 1); }
int trico_read_attributes_uint8(void* a, uint8_t** p) { return RD(a, trico_attribute_uint8_stream, p); }
int trico_read_attributes_uint16(void* a, uint16_t** p) { return RD(a, trico_attribute_uint16_stream, p); }
int trico_read_attributes_uint32(void* a, uint32_t** p) { return RD(a, trico_attribute_uint32_stream, p); }
int trico_read_attributes_uint64(void* a, uint64_t** p) { return RD(a, trico_attribute_uint64_stream, p); }

int trico_skip_next_stream(void* h)
  { /* trico.c:1670-1699 */
  archive* a = (archive*)h;
  if (!a) return 0;
  if (a->next_type == trico_empty) return 1;
  return read_stream(a, a->next_type, NULL, 0);
  }

/* ------------------------------------------------------------------------------------------
 * raw codec API, reference stream format (floating_point_stream_compression.c:86, :212, :576, :803)
 * ------------------------------------------------------------------------------------------ */
static void compress_any(uint32_t* nbytes_out, uint8_t** out, const void* input, uint32_t n, int e1, int e2, int ws)
  {
  *nbytes_out = 0; *out = NULL;
  worker* w = worker_acquire();
  if (!w) return;
  const uint64_t raw = (uint64_t)n * ws, bound = tb200_fpc_v0_bound(ws, n);
  const void* d_in = input;
  int ok = 1;
  if (raw && !tb200_pointer_is_device(input))
    {
    ok = ensure(&w->d_raw, &w->raw_cap, raw + 64, w) && tb200_memcpy_h2d(w->ctx, w->d_raw, input, raw);
    d_in = w->d_raw;
    }
  ok = ok && ensure(&w->d_enc, &w->enc_cap, bound, w);
  ok = ok && tb200_fpc_encode_v0(w->ctx, ws, d_in, n, 1, 1, e1, e2, w->d_enc, bound, (uint32_t*)w->d_scalar);
  uint32_t nb = 0;
  ok = ok && tb200_memcpy_d2h(w->ctx, &nb, w->d_scalar, 4) && tb200_ctx_sync(w->ctx);
  if (ok && nb <= bound)
    {
    uint8_t* p = (uint8_t*)malloc(nb ? nb : 1);      /* exact size, like the reference's final realloc (:209) */
    if (p && tb200_memcpy_d2h(w->ctx, p, w->d_enc, nb) && tb200_ctx_sync(w->ctx)) { *out = p; *nbytes_out = nb; }
    else free(p);
    }
  if (!*out) set_dev_err();
  worker_release(w);
  }

void trico_compress(uint32_t* nb, uint8_t** out, const float* input, const uint32_t n, uint32_t e1, uint32_t e2)
  { compress_any(nb, out, input, n, (int)e1, (int)e2, 4); }
void trico_compress_double_precision(uint32_t* nb, uint8_t** out, const double* input, const uint32_t n, uint64_t e1, uint64_t e2)
  { compress_any(nb, out, input, n, (int)(e1 > 30 ? 30 : e1), (int)(e2 > 30 ? 30 : e2), 8); }

/* The raw decoder gets a bare pointer, so the extent of the stream has to be found before it can
 * be handed to the device: a walk over the code words (no value is decoded here). */
static uint64_t fpc_stream_extent(const uint8_t* s, int ws)
  {
  const uint32_t n = ((uint32_t)s[1] << 24) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 8) | s[4];
  const uint8_t* p = s + 5;
  if (ws == 4)
    for (uint64_t i = 0; i < n || (n == 0 && i == 0); i += 8)
      {
      const uint32_t bc = ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2];
      p += 3;
      for (int j = 0; j < 8; ++j) { const uint32_t c = (bc >> (3 * j)) & 7u; p += c > 4 ? c - 4 : c; }
      }
  else
    for (uint64_t i = 0; i < n || (n == 0 && i == 0); i += 2)
      {
      const uint32_t bc = *p++;
      for (int j = 0; j < 2; ++j) { const uint32_t c = (bc >> (4 * j)) & 15u; p += c > 8 ? c - 8 : c; }
      }
  return (uint64_t)(p - s);
  }

static void decompress_any(uint32_t* n_out, void** out, const uint8_t* compressed, int ws)
  {
  *n_out = 0; *out = NULL;
  worker* w = worker_acquire();
  if (!w) return;
  uint8_t hdr[5];
  const int on_dev = tb200_pointer_is_device(compressed);
  int ok = 1;
  const uint8_t* d_stream = compressed;
  uint32_t n = 0;
  uint64_t extent_bytes = 0;
  if (on_dev)
    {
    ok = tb200_memcpy_d2h(w->ctx, hdr, compressed, 5) && tb200_ctx_sync(w->ctx);
    }
  else
    {
    memcpy(hdr, compressed, 5);
    const uint64_t extent = fpc_stream_extent(compressed, ws);
    extent_bytes = extent;
    ok = ensure(&w->d_enc, &w->enc_cap, extent + 256, w) && tb200_memcpy_h2d(w->ctx, w->d_enc, compressed, extent);
    d_stream = w->d_enc;
    }
  n = ((uint32_t)hdr[1] << 24) | ((uint32_t)hdr[2] << 16) | ((uint32_t)hdr[3] << 8) | hdr[4];
  const uint64_t raw = (uint64_t)n * ws;
  void* host = malloc(raw ? raw : 1);                 /* the reference allocates the result (:231, :822) */
  ok = ok && host && ensure(&w->d_raw, &w->raw_cap, raw + 64, w);
  const uint64_t zero = 0;
  if (ok && n) ok = tb200_fpc_decode_v0(w->ctx, ws, d_stream, &zero, on_dev ? NULL : &extent_bytes, hdr, 1, n, w->d_raw, 1) &&
                   tb200_memcpy_d2h(w->ctx, host, w->d_raw, raw) && tb200_ctx_sync(w->ctx);
  if (ok) { *out = host; *n_out = n; }
  else { set_dev_err(); free(host); }
  worker_release(w);
  }

void trico_decompress(uint32_t* n, float** out, const uint8_t* compressed) { decompress_any(n, (void**)out, compressed, 4); }
void trico_decompress_double_precision(uint32_t* n, double** out, const uint8_t* compressed) { decompress_any(n, (void**)out, compressed, 8); }

/* ------------------------------------------------------------------------------------------
 * transposes into caller-allocated arrays (transpose_aos_to_soa.c:8-147)
 * ------------------------------------------------------------------------------------------ */
static void transpose_any(int to_soa, int ws, int ncomp, void* aos, void* const* comp, uint32_t n)
  {
  if (n == 0) return;
  worker* w = worker_acquire();
  if (!w) return;
  /* ws == 1 means byte planes of `ncomp`-byte elements */
  const uint64_t comp_bytes = (uint64_t)n * ws;
  const uint64_t aos_bytes = comp_bytes * ncomp;
  const uint64_t comp_stride = (comp_bytes + 255) & ~(uint64_t)255;
  int ok = ensure(&w->d_raw, &w->raw_cap, aos_bytes + 64, w) && ensure(&w->d_enc, &w->enc_cap, comp_stride * ncomp, w);
  void* d_comp[8];
  for (int c = 0; c < ncomp; ++c) d_comp[c] = w->d_enc + comp_stride * c;
  if (ok && to_soa)
    {
    ok = tb200_memcpy_h2d(w->ctx, w->d_raw, aos, aos_bytes) &&
         tb200_deinterleave(w->ctx, ws, ncomp, w->d_raw, n, d_comp);
    for (int c = 0; ok && c < ncomp; ++c) ok = tb200_memcpy_d2h(w->ctx, comp[c], d_comp[c], comp_bytes);
    }
  else if (ok)
    {
    for (int c = 0; ok && c < ncomp; ++c) ok = tb200_memcpy_h2d(w->ctx, d_comp[c], comp[c], comp_bytes);
    ok = ok && tb200_interleave(w->ctx, ws, ncomp, w->d_raw, n, (const void* const*)d_comp) &&
         tb200_memcpy_d2h(w->ctx, aos, w->d_raw, aos_bytes);
    }
  ok = ok && tb200_ctx_sync(w->ctx);
  if (!ok) set_dev_err();
  worker_release(w);
  }

void trico_transpose_xyz_aos_to_soa(float** x, float** y, float** z, const float* v, uint32_t n)
  { void* c[3] = {*x, *y, *z}; transpose_any(1, 4, 3, (void*)v, c, n); }
void trico_transpose_xyz_soa_to_aos(float** v, const float* x, const float* y, const float* z, uint32_t n)
  { void* c[3] = {(void*)x, (void*)y, (void*)z}; transpose_any(0, 4, 3, *v, c, n); }
void trico_transpose_xyz_aos_to_soa_double_precision(double** x, double** y, double** z, const double* v, uint32_t n)
  { void* c[3] = {*x, *y, *z}; transpose_any(1, 8, 3, (void*)v, c, n); }
void trico_transpose_xyz_soa_to_aos_double_precision(double** v, const double* x, const double* y, const double* z, uint32_t n)
  { void* c[3] = {(void*)x, (void*)y, (void*)z}; transpose_any(0, 8, 3, *v, c, n); }
void trico_transpose_uv_aos_to_soa(float** u, float** v, const float* uv, uint32_t n)
  { void* c[2] = {*u, *v}; transpose_any(1, 4, 2, (void*)uv, c, n); }
void trico_transpose_uv_soa_to_aos(float** uv, const float* u, const float* v, uint32_t n)
  { void* c[2] = {(void*)u, (void*)v}; transpose_any(0, 4, 2, *uv, c, n); }
void trico_transpose_uv_aos_to_soa_double_precision(double** u, double** v, const double* uv, uint32_t n)
  { void* c[2] = {*u, *v}; transpose_any(1, 8, 2, (void*)uv, c, n); }
void trico_transpose_uv_soa_to_aos_double_precision(double** uv, const double* u, const double* v, uint32_t n)
  { void* c[2] = {(void*)u, (void*)v}; transpose_any(0, 8, 2, *uv, c, n); }
void trico_transpose_uint16_aos_to_soa(uint8_t** b1, uint8_t** b2, const uint16_t* in, uint32_t n)
  { void* c[2] = {*b1, *b2}; transpose_any(1, 1, 2, (void*)in, c, n); }
void trico_transpose_uint16_soa_to_aos(uint16_t** out, const uint8_t* b1, const uint8_t* b2, uint32_t n)
  { void* c[2] = {(void*)b1, (void*)b2}; transpose_any(0, 1, 2, *out, c, n); }
void trico_transpose_uint32_aos_to_soa(uint8_t** b1, uint8_t** b2, uint8_t** b3, uint8_t** b4, const uint32_t* in, uint32_t n)
  { void* c[4] = {*b1, *b2, *b3, *b4}; transpose_any(1, 1, 4, (void*)in, c, n); }
void trico_transpose_uint32_soa_to_aos(uint32_t** out, const uint8_t* b1, const uint8_t* b2, const uint8_t* b3, const uint8_t* b4, uint32_t n)
  { void* c[4] = {(void*)b1, (void*)b2, (void*)b3, (void*)b4}; transpose_any(0, 1, 4, *out, c, n); }
void trico_transpose_uint64_aos_to_soa(uint8_t** b1, uint8_t** b2, uint8_t** b3, uint8_t** b4, uint8_t** b5, uint8_t** b6, uint8_t** b7, uint8_t** b8, const uint64_t* in, uint32_t n)
  { void* c[8] = {*b1, *b2, *b3, *b4, *b5, *b6, *b7, *b8}; transpose_any(1, 1, 8, (void*)in, c, n); }
void trico_transpose_uint64_soa_to_aos(uint64_t** out, const uint8_t* b1, const uint8_t* b2, const uint8_t* b3, const uint8_t* b4, const uint8_t* b5, const uint8_t* b6, const uint8_t* b7, const uint8_t* b8, uint32_t n)
  { void* c[8] = {(void*)b1, (void*)b2, (void*)b3, (void*)b4, (void*)b5, (void*)b6, (void*)b7, (void*)b8}; transpose_any(0, 1, 8, *out, c, n); }
