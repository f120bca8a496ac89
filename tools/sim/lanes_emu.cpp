// lanes_emu.cpp - runs the DEVICE SOURCE of the lane-parallel LZ4 parser (trico_b200/csrc/lz4_lanes.cuh)
// on the CPU: 32 std::threads stand for the 32 lanes of the warp (warp_emu.hpp).  Every block of
// the input file is compressed exactly as lz4_encode_dense_kernel would, decoded again with a
// plain sequential LZ4 decoder and compared.  Test infrastructure only (tests/test_lanes_emu.py).
//
//   g++ -O1 -std=c++20 -pthread -DTB200_HOST_EMU -I tools/sim -I trico_b200/csrc tools/sim/lanes_emu.cpp -o lanes_emu
//   lanes_emu plane.bin [block_bytes] [out_sizes.txt]
#include "lz4_lanes.cuh"

#include <thread>
#include <vector>

using namespace tb200;

static long lz4_decode_plain(const uint8_t* s, size_t n, uint8_t* d, size_t cap)
  {
  size_t ip = 0, op = 0;
  while (ip < n)
    {
    const unsigned tok = s[ip++];
    size_t lit = tok >> 4;
    if (lit == 15) { unsigned b; do { if (ip >= n) return -1; b = s[ip++]; lit += b; } while (b == 255); }
    if (ip + lit > n || op + lit > cap) return -2;
    memcpy(d + op, s + ip, lit); ip += lit; op += lit;
    if (ip >= n) break;
    if (ip + 2 > n) return -3;
    const size_t off = s[ip] | (s[ip + 1] << 8); ip += 2;
    size_t ml = tok & 15;
    if (ml == 15) { unsigned b; do { if (ip >= n) return -4; b = s[ip++]; ml += b; } while (b == 255); }
    ml += 4;
    if (off == 0 || off > op || op + ml > cap) return -5;
    for (size_t i = 0; i < ml; ++i) d[op + i] = d[op + i - off];
    op += ml;
    }
  return (long)op;
  }

int main(int argc, char** argv)
  {
  if (argc < 2) { fprintf(stderr, "usage: lanes_emu plane.bin [block_bytes]\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  std::vector<uint8_t> in;
  uint8_t tmp[65536]; size_t got;
  while ((got = fread(tmp, 1, sizeof(tmp), f)) > 0) in.insert(in.end(), tmp, tmp + got);
  fclose(f);
  const uint32_t B = argc > 2 ? (uint32_t)atoi(argv[2]) : 16384u;
  FILE* fo = argc > 3 ? fopen(argv[3], "w") : nullptr;
  emu_warp warp;
  g_emu_warp = &warp;
  const size_t smem_bytes = lz4_lanes_smem(B);
  std::vector<uint8_t> smem(smem_bytes + 64), out(B + B / 255 + 64), back(B);
  // deliberately dirty shared memory: nothing may depend on what was there before
  for (size_t i = 0; i < smem.size(); ++i) smem[i] = (uint8_t)(i * 131u + 7u);
  uint8_t* base = smem.data() + ((16 - ((uintptr_t)smem.data() & 15)) & 15);
  size_t total = 0, nblocks = 0;
  int rc = 0;
  for (size_t b0 = 0; b0 < in.size(); b0 += B, ++nblocks)
    {
    const uint32_t n = (uint32_t)std::min<size_t>(B, in.size() - b0);
    uint8_t* buf = base;
    memcpy(buf, in.data() + b0, n);
    memset(buf + n, 0, LZ4L_PAD);
    uint16_t* T = reinterpret_cast<uint16_t*>(base + B + LZ4L_PAD);
    uint8_t* own = reinterpret_cast<uint8_t*>(T + (2u << LZ4L_HLOG));
    uint8_t* regions = own + (32u << LZ4L_OWNBITS);
    uint8_t* stage = regions + 32u * LZ4L_REGION;
    uint32_t nbytes[32];
    std::vector<std::thread> th;
    for (unsigned l = 0; l < 32; ++l)
      th.emplace_back([&, l]() { threadIdx.x = l; nbytes[l] = lz4_compress_lanes(buf, n, out.data(), T, own, regions, stage); });
    for (auto& t : th) t.join();
    for (unsigned l = 1; l < 32; ++l) if (nbytes[l] != nbytes[0]) { fprintf(stderr, "block %zu: lanes disagree on the size (%u vs %u)\n", nblocks, nbytes[l], nbytes[0]); rc = 1; }
    if (nbytes[0] > n + n / 255 + 16) { fprintf(stderr, "block %zu: %u bytes exceed the LZ4 bound\n", nblocks, nbytes[0]); rc = 1; }
    const long d = lz4_decode_plain(out.data(), nbytes[0], back.data(), n);
    if (d != (long)n || memcmp(back.data(), in.data() + b0, n) != 0) { fprintf(stderr, "block %zu: round trip failed (decoder returned %ld for %u bytes)\n", nblocks, d, n); rc = 1; }
    if (fo) { fprintf(fo, "%u\n", nbytes[0]); fflush(fo); }
    total += nbytes[0];
    if (warp.mismatches.load()) { fprintf(stderr, "block %zu: lanes left the common path\n", nblocks); rc = 1; break; }
    }
  if (fo) fclose(fo);
  printf("%zu blocks, %zu -> %zu bytes, ratio %.4f, %s\n", nblocks, in.size(), total, total ? (double)in.size() / total : 0.0, rc ? "FAILED" : "ok");
  return rc;
  }
