// sim_capi.cpp — TEST HARNESS ONLY: runs the bitstream kernel SOURCE (scan_bits.cu compiled with
// -DCGX_CPU_SIM against tests/sim/simt_cpu.h) on the CPU SIMT emulator, with the tables the real
// host compiler produces.  Used by tests/test_sim_flat.py to debug warp-level logic without a GPU.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "host/engine.h"
#include "scan_params.h"

namespace cgx {
int64_t scan_flat_chunks(int64_t n);
void sim_launch_scan_flat(const ScanArgs& a, unsigned grid);
}  // namespace cgx

using namespace cgx;

extern "C" {

// text of cgx_jit_prog.h for the pattern (to build the specialised flavour of the emulated kernel)
int cgxsim_jit_header(const char* pat, size_t plen, char* out, size_t cap) {
  std::unique_ptr<Compiled> c;
  std::string err;
  if (CompilePattern(std::string(pat, plen), c, err) != COMPILE_OK) return -1;
  if (c->kind != ENG_DFA || !c->flat.bs_ok) return -2;
  const std::string h = JitHeader(c->flat);
  if (h.size() + 1 > cap) return -3;
  memcpy(out, h.c_str(), h.size() + 1);
  return (int)h.size();
}

// returns 0 ok, -1 compile error, -2 pattern not eligible for the bitstream engine
int cgxsim_scan(const char* pat, size_t plen, const uint8_t* h, int64_t n, int64_t base, int mode,
                int64_t* out, int64_t cap, uint64_t result[4], unsigned grid, int pad_byte, int launches) {
  std::unique_ptr<Compiled> c;
  std::string err;
  if (CompilePattern(std::string(pat, plen), c, err) != COMPILE_OK) return -1;
  if (c->kind != ENG_DFA || !c->flat.bs_ok) return -2;
  // 16-byte aligned copy, padded with bytes the kernel must never interpret
  const size_t padded = ((size_t)n + 15) / 16 * 16 + 64;
  uint8_t* hb = (uint8_t*)aligned_alloc(128, (padded + 127) / 128 * 128);
  memset(hb, pad_byte, (padded + 127) / 128 * 128);
  if (n) memcpy(hb, h, (size_t)n);
  const int64_t nchunks = scan_flat_chunks(n);
  std::vector<unsigned long long> status((size_t)(nchunks > 0 ? nchunks : 1), 0ull);
  const size_t ngroups = (size_t)(nchunks + 31) / 32 + 1;
  std::vector<unsigned long long> gstatus(ngroups, 0ull);
  std::vector<unsigned long long> gacc(ngroups, 0ull);
  unsigned long long scratch[8] = {0};
  ScanArgs a;
  memset(&a, 0, sizeof a);
  a.h = hb;
  a.n = n;
  a.base = base;
  a.dfa.trans = c->dfa.trans.data();
  a.dfa.eoi = c->dfa.eoi.data();
  a.dfa.nstates = c->dfa.nstates;
  for (int k = 0; k < 5; k++) a.dfa.start[k] = c->dfa.start[k];
  a.filter.kind = c->filter_kind;
  a.filter.nranges = c->nranges;
  for (int k = 0; k < 4; k++) {
    a.filter.lo[k] = c->rlo[k];
    a.filter.hi[k] = c->rhi[k];
  }
  a.filter.lut = c->lut;
  a.flat = c->flat;
  a.engine = SEL_DFA;
  a.skip_safe = c->skip_safe ? 1 : 0;
  a.delim = c->delim;
  a.mode = mode;
  a.out = out;
  a.cap = cap;
  a.total = scratch;
  a.ticket = (unsigned*)(scratch + 4);
  a.status = status.data();
  a.nchunks = nchunks;
  a.gstatus = gstatus.data();
  a.gacc = gacc.data();
  // what the host does between launches (capi.cu): words of another epoch read as empty, the
  // accumulators and the ticket are left at zero by the kernel.  `launches` > 1 runs the kernel
  // again on the SAME scratch, with stale look-back words from the launch before.
  unsigned long long result_dev[2] = {~0ull, ~0ull};
  a.result = mode == 0 ? result_dev : nullptr;
  for (int l = 0; l < (launches > 0 ? launches : 1); l++) {
    a.epoch = 7u + (unsigned)l;
    if (mode != 0) memset(scratch, 0, sizeof scratch);
    scratch[0] = scratch[1] = 0xDEADull;  // FindAll overwrites them
    if (mode != 0) scratch[0] = scratch[1] = 0;
    if (nchunks) sim_launch_scan_flat(a, grid);
    if (nchunks && mode != 2 && *a.ticket != 0u) return -5;  // the ticket counter must clean itself
    if (nchunks && mode == 0 && (result_dev[0] != scratch[0] || result_dev[1] != scratch[1])) return -6;
    for (size_t g = 0; g < gacc.size(); g++)
      if (gacc[g]) return -7;  // so must the group accumulators
    if (mode == 2) *a.ticket = 0u;
  }
  if (!nchunks) scratch[0] = scratch[1] = 0;
  result[0] = scratch[0];
  result[1] = scratch[1];
  result[2] = scratch[2];
  result[3] = scratch[3];
  free(hb);
  return 0;
}

}  // extern "C"
