/*
 * trico_oracle.c - TEST INFRASTRUCTURE ONLY (see trico_oracle.h).
 *
 * Plain-C CPU restatement of the reference hot path.  Written from the behaviour of the
 * reference, not from its text: the float and double codecs are one routine over a word-size
 * parameter, the 20 per-type writers/readers are one table-driven routine.
 * "fpc.c" = /root/reference/trico/floating_point_stream_compression.c
 */
#include "trico_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * FPC-style codec.
 *
 * Per value v (bit pattern as unsigned, fpc.c:131 / :620):
 *   FCM  : xor1 = v ^ T1[c1]            then T1[c1] = v,  c1' = top e1 bits of v
 *   DFCM : xor2 = v ^ (last + T2[c2])   then T2[c2] = v - last, c2' = ((c2 << e2/2) ^ top e2 bits of stride) & mask
 * (fpc.c:133-143 / :622-632).  The update order in the reference (store at the *old* hash, then
 * rehash, then fetch the prediction for the next value) is kept by fetching the prediction
 * lazily at the top of the next iteration: nothing else touches the tables in between.
 *
 * Code selection (fpc.c:146-189 / :635-782): n1 = significant bytes of xor1, n2 = of xor2 with
 * a minimum of one byte; xor2 is chosen only if n1 >= 2 and n2 < n1.
 * Float: 3-bit codes, 0..4 = n1, 5..7 = 4+n2, groups of 8 under a 24-bit big-endian word
 * (fpc.c:12-18).  Double: 4-bit codes, 0..8 = n1, 9..15 = 8+n2, groups of 2 under one byte
 * (fpc.c:421-425).  Residual bytes follow the code word MSB first (fpc.c:20-73).
 * Tail: missing slots carry code 1 and one zero byte (fpc.c:196-204, :789-794).
 * ------------------------------------------------------------------------------------------ */

static inline int sig_bytes(uint64_t x)
  {
  int n = 0;
  while (x) { ++n; x >>= 8; }
  return n;
  }

static void norm_exponents(uint32_t* e1, uint32_t* e2)
  {
  /* fpc.c:88-93 / :578-583: force even, cap at 30 */
  *e1 &= ~1u; *e2 &= ~1u;
  if (*e1 > 30) *e1 = 30;
  if (*e2 > 30) *e2 = 30;
  }

uint64_t oracle_fpc32_bound(uint64_t n) { return 5 + 4 * n + 3 * ((n + 7) / 8) + 8; }
uint64_t oracle_fpc64_bound(uint64_t n) { return 5 + 8 * n + (n + 1) / 2 + 2; }

/* wbits = 32 or 64; values are passed as uint64_t getters to share the routine */
static uint64_t fpc_compress(uint8_t* out, const void* in, uint32_t n, uint32_t e1, uint32_t e2, int wbits)
  {
  norm_exponents(&e1, &e2);
  const int group = wbits == 32 ? 8 : 2;
  const int cbits = wbits == 32 ? 3 : 4;
  const int base2 = wbits == 32 ? 4 : 8;     /* code of "xor2, 0 extra" */
  const uint64_t wmask = wbits == 32 ? 0xffffffffull : ~0ull;
  const uint64_t m1 = ((uint64_t)1 << e1) - 1, m2 = ((uint64_t)1 << e2) - 1;
  uint64_t* T1 = (uint64_t*)calloc((size_t)m1 + 1, 8);
  uint64_t* T2 = (uint64_t*)calloc((size_t)m2 + 1, 8);
  uint64_t c1 = 0, c2 = 0, last = 0;
  uint8_t* p = out;
  *p++ = (uint8_t)(((e1 >> 1) << 4) | (e2 >> 1));            /* fpc.c:120 */
  *p++ = (uint8_t)(n >> 24); *p++ = (uint8_t)(n >> 16);       /* fpc.c:123-126 */
  *p++ = (uint8_t)(n >> 8);  *p++ = (uint8_t)n;

  uint64_t res[8]; int code[8]; int nb[8];
  const uint64_t slots = n == 0 ? (uint64_t)group : (((uint64_t)n + group - 1) / group) * group;
  for (uint64_t i = 0; i < slots; ++i)
    {
    int j = (int)(i % group);
    if (i < n)
      {
      uint64_t v = wbits == 32 ? (uint64_t)((const uint32_t*)in)[i] : ((const uint64_t*)in)[i];
      uint64_t x1 = v ^ T1[c1];
      uint64_t x2 = v ^ ((last + T2[c2]) & wmask);
      uint64_t stride = (v - last) & wmask;
      T1[c1] = v;
      c1 = e1 ? ((c1 << e1) ^ (v >> (wbits - e1))) & m1 : 0;
      T2[c2] = stride;
      c2 = e2 ? ((c2 << (e2 / 2)) ^ (stride >> (wbits - e2))) & m2 : 0;
      last = v;
      int n1 = sig_bytes(x1), n2 = sig_bytes(x2);
      if (n2 == 0) n2 = 1;
      if (n1 >= 2 && n2 < n1) { code[j] = base2 + n2; nb[j] = n2; res[j] = x2; }
      else                    { code[j] = n1;         nb[j] = n1; res[j] = x1; }
      }
    else
      {
      /* pad slot: code 1 + one zero byte (fpc.c:196-204, :789-794).  For n == 0 the reference
       * emits one group whose slot 0 is uninitialised stack; the oracle pads slot 0 too. */
      code[j] = 1; nb[j] = 1; res[j] = 0;
      }
    if (j == group - 1)
      {
      uint32_t bc = 0;
      for (int k = 0; k < group; ++k) bc |= (uint32_t)code[k] << (cbits * k);
      if (wbits == 32) { *p++ = (uint8_t)(bc >> 16); *p++ = (uint8_t)(bc >> 8); *p++ = (uint8_t)bc; }
      else             { *p++ = (uint8_t)bc; }
      for (int k = 0; k < group; ++k)
        for (int b = nb[k] - 1; b >= 0; --b)
          *p++ = (uint8_t)(res[k] >> (8 * b));
      }
    }
  free(T1); free(T2);
  return (uint64_t)(p - out);
  }

uint64_t oracle_fpc32_compress(uint8_t* out, const uint32_t* in, uint32_t n, uint32_t e1, uint32_t e2)
  { return fpc_compress(out, in, n, e1, e2, 32); }
uint64_t oracle_fpc64_compress(uint8_t* out, const uint64_t* in, uint32_t n, uint32_t e1, uint32_t e2)
  { return fpc_compress(out, in, n, e1, e2, 64); }

/* Shared decoder.  out may be NULL (then only the length walk is done). Returns bytes consumed
 * through *consumed and the value count as result.  fpc.c:212-417 / :803-1164. */
static uint32_t fpc_decompress(void* out, const uint8_t* in, int wbits, uint64_t* consumed)
  {
  const uint8_t* p = in;
  uint32_t e1 = (uint32_t)(*p >> 4) << 1, e2 = (uint32_t)(*p & 15) << 1;   /* fpc.c:214-217 */
  ++p;
  uint32_t n = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  p += 4;
  const int group = wbits == 32 ? 8 : 2;
  const int cbits = wbits == 32 ? 3 : 4;
  const int base2 = wbits == 32 ? 4 : 8;
  const uint64_t wmask = wbits == 32 ? 0xffffffffull : ~0ull;
  const uint64_t m1 = ((uint64_t)1 << e1) - 1, m2 = ((uint64_t)1 << e2) - 1;
  uint64_t *T1 = NULL, *T2 = NULL;
  if (out) { T1 = (uint64_t*)calloc((size_t)m1 + 1, 8); T2 = (uint64_t*)calloc((size_t)m2 + 1, 8); }
  uint64_t c1 = 0, c2 = 0, last = 0;
  uint32_t i = 0;
  while (i < n)
    {
    uint32_t bc;
    if (wbits == 32) { bc = ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2]; p += 3; }
    else             { bc = *p++; }
    for (int k = 0; k < group; ++k)
      {
      int code = (int)((bc >> (cbits * k)) & ((1u << cbits) - 1));
      int use2 = code > base2;
      int nbytes = use2 ? code - base2 : code;
      uint64_t x = 0;
      for (int b = 0; b < nbytes; ++b) x = (x << 8) | *p++;
      if (i < n)
        {
        /* the reference tail loop stops at the first (code 1, byte 0) pad (fpc.c:346-351); with a
         * well-formed stream that is exactly slot n % group, which `i < n` expresses. */
        if (out)
          {
          uint64_t pred = use2 ? ((last + T2[c2]) & wmask) : T1[c1];
          uint64_t v = x ^ pred;
          uint64_t stride = (v - last) & wmask;
          T1[c1] = v;
          c1 = e1 ? ((c1 << e1) ^ (v >> (wbits - e1))) & m1 : 0;
          T2[c2] = stride;
          c2 = e2 ? ((c2 << (e2 / 2)) ^ (stride >> (wbits - e2))) & m2 : 0;
          last = v;
          if (wbits == 32) ((uint32_t*)out)[i] = (uint32_t)v; else ((uint64_t*)out)[i] = v;
          }
        ++i;
        }
      }
    }
  if (n == 0)
    { /* one all-pad group follows the header (see encoder) */
    p += (wbits == 32 ? 3 + 8 : 1 + 2);
    }
  free(T1); free(T2);
  if (consumed) *consumed = (uint64_t)(p - in);
  return n;
  }

uint32_t oracle_fpc32_decompress(uint32_t* out, const uint8_t* in) { return fpc_decompress(out, in, 32, NULL); }
uint32_t oracle_fpc64_decompress(uint64_t* out, const uint8_t* in) { return fpc_decompress(out, in, 64, NULL); }
uint64_t oracle_fpc32_stream_bytes(const uint8_t* in) { uint64_t c; fpc_decompress(NULL, in, 32, &c); return c; }
uint64_t oracle_fpc64_stream_bytes(const uint8_t* in) { uint64_t c; fpc_decompress(NULL, in, 64, &c); return c; }

/* ------------------------------------------------------------------------------------------
 * AoS <-> SoA (transpose_aos_to_soa.c:8-82) and byte planes (:84-147): plane k holds byte k
 * (least significant first) of every element.
 * ------------------------------------------------------------------------------------------ */
void oracle_deinterleave(void* soa, const void* aos, uint64_t n, int ncomp, int wordsize)
  {
  const uint8_t* a = (const uint8_t*)aos; uint8_t* s = (uint8_t*)soa;
  for (uint64_t i = 0; i < n; ++i)
    for (int c = 0; c < ncomp; ++c)
      memcpy(s + ((uint64_t)c * n + i) * wordsize, a + (i * ncomp + c) * wordsize, (size_t)wordsize);
  }

void oracle_interleave(void* aos, const void* soa, uint64_t n, int ncomp, int wordsize)
  {
  uint8_t* a = (uint8_t*)aos; const uint8_t* s = (const uint8_t*)soa;
  for (uint64_t i = 0; i < n; ++i)
    for (int c = 0; c < ncomp; ++c)
      memcpy(a + (i * ncomp + c) * wordsize, s + ((uint64_t)c * n + i) * wordsize, (size_t)wordsize);
  }

void oracle_planes_split(uint8_t* planes, const void* in, uint64_t n, int wordsize)
  {
  const uint8_t* a = (const uint8_t*)in;      /* little-endian host: byte k of element i is a[i*W+k] */
  for (uint64_t i = 0; i < n; ++i)
    for (int k = 0; k < wordsize; ++k)
      planes[(uint64_t)k * n + i] = a[i * wordsize + k];
  }

void oracle_planes_merge(void* out, const uint8_t* planes, uint64_t n, int wordsize)
  {
  uint8_t* a = (uint8_t*)out;
  for (uint64_t i = 0; i < n; ++i)
    for (int k = 0; k < wordsize; ++k)
      a[i * wordsize + k] = planes[(uint64_t)k * n + i];
  }

/* ------------------------------------------------------------------------------------------
 * LZ4 block format.  Sequence = token (hi nibble literal length, lo nibble match length - 4),
 * 255-extension bytes, literals, u16 LE offset, 255-extension bytes (lz4.c:1629-1649,
 * :1848-2059).  The block ends with a literals-only sequence.
 * ------------------------------------------------------------------------------------------ */
#define LZ4_MINMATCH 4
#define LZ4_LASTLITERALS 5      /* lz4.c:192 */
#define LZ4_MFLIMIT 12          /* lz4.c:193 */

static int64_t lz4_walk(uint8_t* dst, uint64_t dst_cap, const uint8_t* src, uint64_t src_len, int strict)
  {
  uint64_t ip = 0, op = 0;
  if (src_len == 0) return -1;
  for (;;)
    {
    if (ip >= src_len) return -2;
    unsigned token = src[ip++];
    uint64_t lit = token >> 4;
    if (lit == 15)
      {
      unsigned b;
      do { if (ip >= src_len) return -3; b = src[ip++]; lit += b; } while (b == 255);
      }
    if (ip + lit > src_len || op + lit > dst_cap) return -4;
    if (dst) memcpy(dst + op, src + ip, (size_t)lit);
    ip += lit; op += lit;
    if (ip == src_len)
      {
      if (strict && (token & 15) != 0) return -5;
      if (strict && op >= 1 && lit < LZ4_LASTLITERALS && op > lit) return -6; /* last 5 bytes must be literals */
      return (int64_t)op;
      }
    if (ip + 2 > src_len) return -7;
    uint64_t offset = (uint64_t)src[ip] | ((uint64_t)src[ip + 1] << 8);
    ip += 2;
    if (offset == 0 || offset > op) return -8;
    uint64_t mlen = token & 15;
    if (mlen == 15)
      {
      unsigned b;
      do { if (ip >= src_len) return -9; b = src[ip++]; mlen += b; } while (b == 255);
      }
    mlen += LZ4_MINMATCH;
    if (op + mlen > dst_cap) return -10;
    if (strict && op + LZ4_MFLIMIT > dst_cap) return -11;          /* match must start >= 12 bytes before the end */
    if (strict && op + mlen + LZ4_LASTLITERALS > dst_cap) return -12;
    if (dst) for (uint64_t k = 0; k < mlen; ++k) dst[op + k] = dst[op + k - offset];
    op += mlen;
    }
  }

int64_t oracle_lz4_decompress(uint8_t* dst, uint64_t dst_cap, const uint8_t* src, uint64_t src_len)
  { return lz4_walk(dst, dst_cap, src, src_len, 0); }

int64_t oracle_lz4_validate(const uint8_t* src, uint64_t src_len, uint64_t expect_raw)
  {
  uint8_t* tmp = (uint8_t*)malloc((size_t)expect_raw + 1);
  int64_t r = lz4_walk(tmp, expect_raw, src, src_len, 1);
  free(tmp);
  if (r >= 0 && (uint64_t)r != expect_raw) return -20;
  return r;
  }

uint64_t oracle_lz4_bound(uint64_t n) { return n + n / 255 + 16; }   /* lz4.h:171 */

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

static uint8_t* lz4_emit(uint8_t* op, const uint8_t* lit, uint64_t nlit, uint64_t offset, uint64_t mlen)
  {
  uint8_t* token = op++;
  uint64_t l = nlit;
  if (l >= 15) { *token = 15 << 4; l -= 15; while (l >= 255) { *op++ = 255; l -= 255; } *op++ = (uint8_t)l; }
  else *token = (uint8_t)(l << 4);
  memcpy(op, lit, (size_t)nlit); op += nlit;
  if (mlen)
    {
    *op++ = (uint8_t)offset; *op++ = (uint8_t)(offset >> 8);
    uint64_t m = mlen - LZ4_MINMATCH;
    if (m >= 15) { *token |= 15; m -= 15; while (m >= 255) { *op++ = 255; m -= 255; } *op++ = (uint8_t)m; }
    else *token |= (uint8_t)m;
    }
  return op;
  }

uint64_t oracle_lz4_compress(uint8_t* dst, const uint8_t* src, uint64_t n)
  {
  enum { HLOG = 14 };
  uint8_t* op = dst;
  uint64_t anchor = 0;
  if (n >= LZ4_MFLIMIT + 1)
    {
    int64_t* table = (int64_t*)malloc(sizeof(int64_t) << HLOG);
    for (int i = 0; i < (1 << HLOG); ++i) table[i] = -1;
    const uint64_t mflimit = n - LZ4_MFLIMIT;      /* last position a match may start at */
    const uint64_t matchlimit = n - LZ4_LASTLITERALS;
    uint64_t ip = 0;
    while (ip <= mflimit)
      {
      uint32_t seq = rd32(src + ip);
      uint32_t h = (seq * 2654435761u) >> (32 - HLOG);
      int64_t cand = table[h];
      table[h] = (int64_t)ip;
      if (cand >= 0 && ip - (uint64_t)cand <= 65535 && rd32(src + cand) == seq)
        {
        uint64_t mlen = 4;
        while (ip + mlen < matchlimit && src[cand + mlen] == src[ip + mlen]) ++mlen;
        op = lz4_emit(op, src + anchor, ip - anchor, ip - (uint64_t)cand, mlen);
        ip += mlen; anchor = ip;
        }
      else ++ip;
      }
    free(table);
    }
  op = lz4_emit(op, src + anchor, n - anchor, 0, 0);
  return (uint64_t)(op - dst);
  }

/* ------------------------------------------------------------------------------------------
 * Stream layouts (trico.h:11-34; trico.c writers :215-858).
 * ------------------------------------------------------------------------------------------ */
int oracle_stream_layout(int type, oracle_layout* lay)
  {
  static const oracle_layout T[21] = {
    {0,0,0,0},
    {1,4,3,1}, {1,8,3,1},            /* vertex float / double            trico.c:264,:429 */
    {2,4,1,3}, {2,8,1,3},            /* triangle u32 / u64 (3 per count) trico.c:323,:444 */
    {1,4,2,1}, {1,8,2,1},            /* uv per vertex float / double     trico.c:572,:620 */
    {1,4,2,1}, {1,8,2,1},            /* uv per triangle float / double   trico.c:577,:625 */
    {1,4,3,1}, {1,8,3,1},            /* vertex normal float / double     trico.c:269,:434 */
    {1,4,3,1}, {1,8,3,1},            /* triangle normal float / double   trico.c:274,:439 */
    {2,4,1,1}, {2,4,1,1},            /* vertex / triangle colour         trico.c:760,:765 */
    {1,4,1,1}, {1,8,1,1},            /* attribute float / double         trico.c:279,:301 */
    {2,1,1,1}, {2,2,1,1}, {2,4,1,1}, {2,8,1,1} /* attribute u8/u16/u32/u64 trico.c:630,:657,:755,:770 */
  };
  if (type < 1 || type > 20) return 0;
  *lay = T[type];
  return 1;
  }

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t get32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static void put64(uint8_t* p, uint64_t v) { put32(p, (uint32_t)v); put32(p + 4, (uint32_t)(v >> 32)); }
static uint64_t get64(const uint8_t* p) { return (uint64_t)get32(p) | ((uint64_t)get32(p + 4) << 32); }

uint64_t oracle_v0_write_header(uint8_t* out, uint32_t version)
  {
  put32(out, 0x6f637254u);       /* "Trco", trico.c:94 */
  put32(out + 4, version);
  return 8;
  }

uint64_t oracle_v0_stream_bound(int type, uint32_t count)
  {
  oracle_layout L;
  if (!oracle_stream_layout(type, &L)) return 0;
  uint64_t n = (uint64_t)count * L.per_count;
  if (L.codec == 1) return 5 + (uint64_t)L.ncomp * (4 + (L.wordsize == 4 ? oracle_fpc32_bound(n) : oracle_fpc64_bound(n)));
  return 5 + (uint64_t)L.wordsize * (4 + oracle_lz4_bound(n));
  }

/* v0 stream: u8 type, u32 count, then per component / byte plane: u32 nbytes + payload
 * (trico.c:218-260, :326-368).  float/double exponents are the reference's call-site constants
 * (4,10) and (20,20) (trico.c:231, :396). */
uint64_t oracle_v0_write_stream(uint8_t* out, int type, const void* data, uint32_t count)
  {
  oracle_layout L;
  if (!oracle_stream_layout(type, &L)) return 0;
  uint8_t* p = out;
  *p++ = (uint8_t)type;
  put32(p, count); p += 4;
  uint64_t n = (uint64_t)count * L.per_count;
  if (L.codec == 1)
    {
    uint8_t* soa = (uint8_t*)malloc((size_t)(n * L.ncomp * L.wordsize) + 8);
    oracle_deinterleave(soa, data, n, L.ncomp, L.wordsize);
    for (int c = 0; c < L.ncomp; ++c)
      {
      const uint8_t* comp = soa + (uint64_t)c * n * L.wordsize;
      uint64_t nb = L.wordsize == 4 ? oracle_fpc32_compress(p + 4, (const uint32_t*)comp, (uint32_t)n, 4, 10)
                                    : oracle_fpc64_compress(p + 4, (const uint64_t*)comp, (uint32_t)n, 20, 20);
      put32(p, (uint32_t)nb); p += 4 + nb;
      }
    free(soa);
    }
  else
    {
    uint8_t* planes = (uint8_t*)malloc((size_t)(n * L.wordsize) + 8);
    oracle_planes_split(planes, data, n, L.wordsize);
    for (int k = 0; k < L.wordsize; ++k)
      {
      uint64_t nb = oracle_lz4_compress(p + 4, planes + (uint64_t)k * n, n);
      put32(p, (uint32_t)nb); p += 4 + nb;
      }
    free(planes);
    }
  return (uint64_t)(p - out);
  }

uint64_t oracle_v0_read_stream(void* out, const uint8_t* in, uint64_t avail, int* type, uint32_t* count)
  {
  oracle_layout L;
  if (avail < 5) return 0;
  const uint8_t* p = in;
  int t = *p++;
  if (!oracle_stream_layout(t, &L)) return 0;
  uint32_t cnt = get32(p); p += 4;
  if (type) *type = t;
  if (count) *count = cnt;
  uint64_t n = (uint64_t)cnt * L.per_count;
  int nsub = L.codec == 1 ? L.ncomp : L.wordsize;
  uint8_t* tmp = out ? (uint8_t*)malloc((size_t)(n * nsub * (L.codec == 1 ? L.wordsize : 1)) + 8) : NULL;
  for (int s = 0; s < nsub; ++s)
    {
    if ((uint64_t)(p - in) + 4 > avail) { free(tmp); return 0; }
    uint32_t nb = get32(p); p += 4;
    if ((uint64_t)(p - in) + nb > avail) { free(tmp); return 0; }
    if (out)
      {
      if (L.codec == 1)
        {
        if (L.wordsize == 4) oracle_fpc32_decompress((uint32_t*)(tmp + (uint64_t)s * n * 4), p);
        else                 oracle_fpc64_decompress((uint64_t*)(tmp + (uint64_t)s * n * 8), p);
        }
      else if (oracle_lz4_decompress(tmp + (uint64_t)s * n, n, p, nb) != (int64_t)n) { free(tmp); return 0; }
      }
    p += nb;
    }
  if (out)
    {
    if (L.codec == 1) oracle_interleave(out, tmp, n, L.ncomp, L.wordsize);
    else              oracle_planes_merge(out, tmp, n, L.wordsize);
    free(tmp);
    }
  return (uint64_t)(p - in);
  }

/* ------------------------------------------------------------------------------------------
 * v1 chunked stream (our container):
 *   u8 type, u32 count, u8 codec_info, u8 log2_chunk, u64 payload_bytes, u16 sizes[nchunks], payload
 * FPC : chunk (k,c) = values [k*S, min((k+1)*S, n)) of component c, index k*ncomp+c, payload =
 *       reference FPC stream of those values minus its 5-byte header; codec_info = hash_info.
 * LZ4 : chunk (k,p) = bytes [k*B, ...) of byte plane p, index k*wordsize+p, payload = one LZ4 block.
 * ------------------------------------------------------------------------------------------ */
#define V1_FIXED 15

static uint64_t v1_nranges(uint64_t n, int log2_chunk) { return (n + ((uint64_t)1 << log2_chunk) - 1) >> log2_chunk; }

uint64_t oracle_v1_stream_bound(int type, uint32_t count, int log2_chunk)
  {
  oracle_layout L;
  if (!oracle_stream_layout(type, &L)) return 0;
  uint64_t n = (uint64_t)count * L.per_count;
  uint64_t S = (uint64_t)1 << log2_chunk;
  uint64_t nr = v1_nranges(n, log2_chunk);
  int nsub = L.codec == 1 ? L.ncomp : L.wordsize;
  uint64_t per = L.codec == 1 ? (L.wordsize == 4 ? oracle_fpc32_bound(S) : oracle_fpc64_bound(S)) : oracle_lz4_bound(S);
  return V1_FIXED + nr * nsub * (2 + per);
  }

uint64_t oracle_v1_write_stream(uint8_t* out, int type, const void* data, uint32_t count, int log2_chunk, int e1, int e2)
  {
  oracle_layout L;
  if (!oracle_stream_layout(type, &L)) return 0;
  uint64_t n = (uint64_t)count * L.per_count;
  uint64_t S = (uint64_t)1 << log2_chunk;
  uint64_t nr = v1_nranges(n, log2_chunk);
  int nsub = L.codec == 1 ? L.ncomp : L.wordsize;
  uint8_t* p = out;
  *p++ = (uint8_t)type; put32(p, count); p += 4;
  *p++ = L.codec == 1 ? (uint8_t)((((unsigned)e1 >> 1) << 4) | ((unsigned)e2 >> 1)) : 0;
  *p++ = (uint8_t)log2_chunk;
  uint8_t* total_field = p; p += 8;
  uint8_t* sizes = p; p += 2 * nr * nsub;
  uint8_t* payload0 = p;
  uint8_t* tmp = (uint8_t*)malloc((size_t)(L.codec == 1 ? oracle_fpc64_bound(S) : oracle_lz4_bound(S)) + 16);
  uint8_t* soa = (uint8_t*)malloc((size_t)(n * nsub * (L.codec == 1 ? L.wordsize : 1)) + 8);
  if (L.codec == 1) oracle_deinterleave(soa, data, n, L.ncomp, L.wordsize);
  else              oracle_planes_split(soa, data, n, L.wordsize);
  for (uint64_t k = 0; k < nr; ++k)
    {
    uint64_t lo = k * S, cnt = n - lo < S ? n - lo : S;
    for (int s = 0; s < nsub; ++s)
      {
      uint64_t nb;
      if (L.codec == 1)
        {
        const uint8_t* comp = soa + ((uint64_t)s * n + lo) * L.wordsize;
        uint64_t full = L.wordsize == 4 ? oracle_fpc32_compress(tmp, (const uint32_t*)comp, (uint32_t)cnt, (uint32_t)e1, (uint32_t)e2)
                                        : oracle_fpc64_compress(tmp, (const uint64_t*)comp, (uint32_t)cnt, (uint32_t)e1, (uint32_t)e2);
        nb = full - 5;
        memcpy(p, tmp + 5, (size_t)nb);
        }
      else
        {
        nb = oracle_lz4_compress(p, soa + (uint64_t)s * n + lo, cnt);
        }
      uint8_t* sz = sizes + 2 * (k * nsub + s);
      sz[0] = (uint8_t)nb; sz[1] = (uint8_t)(nb >> 8);
      p += nb;
      }
    }
  put64(total_field, (uint64_t)(p - payload0));
  free(tmp); free(soa);
  return (uint64_t)(p - out);
  }

uint64_t oracle_v1_read_stream(void* out, const uint8_t* in, uint64_t avail, int* type, uint32_t* count)
  {
  oracle_layout L;
  if (avail < V1_FIXED) return 0;
  const uint8_t* p = in;
  int t = *p++;
  if (!oracle_stream_layout(t, &L)) return 0;
  uint32_t cnt32 = get32(p); p += 4;
  unsigned info = *p++;
  int log2_chunk = *p++;
  uint64_t total = get64(p); p += 8;
  if (type) *type = t;
  if (count) *count = cnt32;
  uint64_t n = (uint64_t)cnt32 * L.per_count;
  uint64_t S = (uint64_t)1 << log2_chunk;
  uint64_t nr = v1_nranges(n, log2_chunk);
  int nsub = L.codec == 1 ? L.ncomp : L.wordsize;
  const uint8_t* sizes = p; p += 2 * nr * nsub;
  if ((uint64_t)(p - in) + total > avail) return 0;
  const uint8_t* end = p + total;
  if (out)
    {
    uint8_t* soa = (uint8_t*)malloc((size_t)(n * nsub * (L.codec == 1 ? L.wordsize : 1)) + 8);
    uint8_t* tmp = (uint8_t*)malloc((size_t)oracle_fpc64_bound(S) + 16);
    for (uint64_t k = 0; k < nr; ++k)
      {
      uint64_t lo = k * S, c = n - lo < S ? n - lo : S;
      for (int s = 0; s < nsub; ++s)
        {
        const uint8_t* sz = sizes + 2 * (k * nsub + s);
        uint64_t nb = (uint64_t)sz[0] | ((uint64_t)sz[1] << 8);
        if (L.codec == 1)
          {
          tmp[0] = (uint8_t)info;
          tmp[1] = (uint8_t)(c >> 24); tmp[2] = (uint8_t)(c >> 16); tmp[3] = (uint8_t)(c >> 8); tmp[4] = (uint8_t)c;
          memcpy(tmp + 5, p, (size_t)nb);
          if (L.wordsize == 4) oracle_fpc32_decompress((uint32_t*)(soa + ((uint64_t)s * n + lo) * 4), tmp);
          else                 oracle_fpc64_decompress((uint64_t*)(soa + ((uint64_t)s * n + lo) * 8), tmp);
          }
        else if (oracle_lz4_decompress(soa + (uint64_t)s * n + lo, c, p, nb) != (int64_t)c) { free(soa); free(tmp); return 0; }
        p += nb;
        }
      }
    if (L.codec == 1) oracle_interleave(out, soa, n, L.ncomp, L.wordsize);
    else              oracle_planes_merge(out, soa, n, L.wordsize);
    free(soa); free(tmp);
    }
  return (uint64_t)(end - in);
  }

/* =============================================================================================
 * Mesh front-end: what trico_read_stl does after reading the facets, and the
 * normal recomputation of trico_decoder.
 * ============================================================================================= */
#include <math.h>

typedef struct { float x, y, z; uint32_t id; } oracle_corner;

/* iostl.c:8-19 (less) with the corner id as the tie-break that makes the order total */
static int oracle_corner_cmp(const void* pa, const void* pb)
  {
  const oracle_corner* a = (const oracle_corner*)pa;
  const oracle_corner* b = (const oracle_corner*)pb;
  if (a->x != b->x) return a->x < b->x ? -1 : 1;
  if (a->y != b->y) return a->y < b->y ? -1 : 1;
  if (a->z != b->z) return a->z < b->z ? -1 : 1;
  return a->id < b->id ? -1 : (a->id > b->id ? 1 : 0);
  }

uint32_t oracle_stl_dedup(const uint8_t* facets, uint32_t ntriangles, float* vertices, uint32_t* triangles)
  {
  if (ntriangles == 0) return 0;                                     /* iostl.c:72-73 */
  const uint64_t n = (uint64_t)ntriangles * 3;
  oracle_corner* c = (oracle_corner*)malloc(n * sizeof(oracle_corner));
  for (uint32_t t = 0; t < ntriangles; ++t)                           /* iostl.c:78-102: corner = position + its slot */
    for (int j = 0; j < 3; ++j)
      {
      oracle_corner* q = c + (uint64_t)t * 3 + j;
      memcpy(&q->x, facets + (uint64_t)t * 50 + 12 + 12 * j, 12);    /* iostl.c:175-183 */
      q->id = t * 3 + (uint32_t)j;
      }
  qsort(c, n, sizeof(oracle_corner), oracle_corner_cmp);              /* iostl.c:104 */
  uint32_t nv = 0;                                                    /* iostl.c:107-137: a new vertex where neighbours differ */
  for (uint64_t i = 0; i < n; ++i)
    {
    if (i == 0 || !(c[i].x == c[i - 1].x && c[i].y == c[i - 1].y && c[i].z == c[i - 1].z))
      {
      vertices[(uint64_t)nv * 3] = c[i].x; vertices[(uint64_t)nv * 3 + 1] = c[i].y; vertices[(uint64_t)nv * 3 + 2] = c[i].z;
      ++nv;
      }
    triangles[c[i].id] = nv - 1;
    }
  free(c);
  return nv;
  }

void oracle_triangle_normals(const float* vertices, const uint32_t* triangles, uint32_t ntriangles, float* normals)
  {
  for (uint32_t t = 0; t < ntriangles; ++t)
    {
    const float* p0 = vertices + (uint64_t)triangles[(uint64_t)t * 3] * 3;
    const float* p1 = vertices + (uint64_t)triangles[(uint64_t)t * 3 + 1] * 3;
    const float* p2 = vertices + (uint64_t)triangles[(uint64_t)t * 3 + 2] * 3;
    /* volatile: every operation rounded to float on its own, whatever the compiler flags (main.c:455-464) */
    volatile float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
    volatile float bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
    volatile float m0 = ay * bz, m1 = az * by, m2 = az * bx, m3 = ax * bz, m4 = ax * by, m5 = ay * bx;
    volatile float nx = m0 - m1, ny = m2 - m3, nz = m4 - m5;
    volatile float s0 = nx * nx, s1 = ny * ny, s2 = nz * nz;
    volatile float s01 = s0 + s1;
    volatile float s = s01 + s2;
    const float length = (float)sqrt((double)s);                      /* main.c:465 */
    normals[(uint64_t)t * 3] = length ? nx / length : nx;
    normals[(uint64_t)t * 3 + 1] = length ? ny / length : ny;
    normals[(uint64_t)t * 3 + 2] = length ? nz / length : nz;
    }
  }
