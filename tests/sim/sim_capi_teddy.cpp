// sim_capi_teddy.cpp — TEST HARNESS ONLY: runs the multi-literal flavour of the bitstream kernel
// (scan_teddy.cu = scan_bits.cu with CGX_TEDDY, compiled with -DCGX_CPU_SIM against simt_cpu.h) on
// the CPU SIMT emulator, with the tables the real host compiler produces (the blob layout of capi.cu).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "host/engine.h"
#include "scan_params.h"

namespace cgx {
int64_t scan_teddy_chunks(int64_t n);
void sim_launch_scan_teddy(const ScanArgs& a, unsigned grid);
}  // namespace cgx

using namespace cgx;

extern "C" {

// returns 0 ok, -1 compile error, -2 pattern is not a literal set the engine takes
int cgxsim_scan_teddy(const char* pat, size_t plen, const uint8_t* h, int64_t n, int64_t base, int64_t after, int mode,
                      int64_t* out, int64_t cap, uint64_t result[4], unsigned grid, int pad_byte, int launches) {
  std::unique_ptr<Compiled> c;
  std::string err;
  if (CompilePattern(std::string(pat, plen), c, err) != COMPILE_OK) return -1;
  if (c->kind != ENG_TEDDY || c->teddy.max_len > 32) return -2;
  const TeddyTables& t = c->teddy;
  std::vector<uint64_t> lit8(2 * (size_t)t.npat, 0);
  for (int id = 0; id < t.npat; id++) {
    const int len = t.offs[id + 1] - t.offs[id];
    for (int k = 0; k < 8 && k < len; k++) {
      lit8[id] |= (uint64_t)t.bytes[t.offs[id] + k] << (8 * k);
      lit8[t.npat + id] |= 0xFFull << (8 * k);
    }
  }
  std::vector<uint16_t> fp2(256, 0);
  for (int id = 0; id < t.npat; id++) fp2[t.bytes[t.offs[id] + 2]] |= (uint16_t)(1u << t.bucket_of[id]);
  const size_t padded = ((size_t)n + 15) / 16 * 16 + 64;
  uint8_t* hb = (uint8_t*)aligned_alloc(128, (padded + 127) / 128 * 128);
  memset(hb, pad_byte, (padded + 127) / 128 * 128);
  if (n) memcpy(hb, h, (size_t)n);
  const int64_t nchunks = scan_teddy_chunks(n);
  std::vector<unsigned long long> status((size_t)(nchunks > 0 ? nchunks : 1), 0ull);
  const size_t ngroups = (size_t)(nchunks + 31) / 32 + 1;
  std::vector<unsigned long long> gstatus(ngroups, 0ull), gacc(ngroups, 0ull);
  unsigned long long scratch[8] = {0};
  ScanArgs a;
  memset(&a, 0, sizeof a);
  a.h = hb;
  a.n = n;
  a.base = base;
  a.after = after;
  a.teddy.fp = t.fp_packed.data();
  a.teddy.lit8 = lit8.data();
  a.teddy.fp2 = fp2.data();
  a.teddy.bytes = t.bytes.data();
  a.teddy.offs = t.offs.data();
  a.teddy.order = t.order_simd.data();
  a.teddy.bucket_off = t.bucket_off.data();
  a.teddy.npat = t.npat;
  a.teddy.nbuckets = t.nbuckets;
  a.teddy.min_len = t.min_len;
  a.teddy.max_len = t.max_len;
  a.engine = SEL_TEDDY;
  a.delim = '\n';
  a.mode = mode;
  a.out = out;
  a.cap = cap;
  a.total = scratch;
  a.ticket = (unsigned*)(scratch + 4);
  a.status = status.data();
  a.nchunks = nchunks;
  a.gstatus = gstatus.data();
  a.gacc = gacc.data();
  unsigned long long result_dev[2] = {~0ull, ~0ull};
  a.result = mode == 0 ? result_dev : nullptr;
  for (int l = 0; l < (launches > 0 ? launches : 1); l++) {
    a.epoch = 7u + (unsigned)l;
    memset(scratch, 0, sizeof scratch);
    if (nchunks) sim_launch_scan_teddy(a, grid);
    if (nchunks && mode != 2 && *a.ticket != 0u) return -5;
    if (nchunks && mode == 0 && (result_dev[0] != scratch[0] || result_dev[1] != scratch[1])) return -6;
    for (size_t g = 0; g < gacc.size(); g++)
      if (gacc[g]) return -7;
    if (mode == 2) *a.ticket = 0u;
  }
  for (int k = 0; k < 4; k++) result[k] = scratch[k];
  free(hb);
  return 0;
}

}  // extern "C"
