// flat_caps.cu — capture-group offsets of FLAT deterministic patterns without an NFA simulation.
//
// Replaces, for patterns the bitstream kernel takes (scan_bits.cu; `(\w+)@(\w+)\.(\w+)`,
// `(\d+)\.(\d+)\.(\d+)\.(\d+)`, `([a-z]+)=(\d{1,3})` ...), the capture pass of
//   reference meta/findall.go:390 FindAllSubmatch -> nfa/pikevm.go:2186 SearchWithSlotTableCapturesAt
// A flat pattern is a concatenation of byte-class items (one byte, C+, C*, C?), and the host proved it
// deterministic (host/engine.cpp DecideBitstream): from a match start the leftmost-first match is
// the forced greedy walk.  A capture group that encloses whole items therefore opens and closes at
// ITEM BOUNDARIES of that walk (host/engine.h FlatCaps), so the group offsets of a match are the
// positions the walk stands at before items at[2g] and at[2g+1] — one lane per match replays the
// walk over the match's own bytes (a few dozen L2-resident bytes) instead of running the PikeVM.
// Groups under a quantifier keep the Pike captures kernel (pikevm_kernel.cu).
#include <cuda_runtime.h>

#include <cstdint>

#include "scan_params.h"

namespace cgx {

namespace {

struct FlatCapArgs {
  const uint8_t* h;
  int64_t n, base;
  const int64_t* matches;             // (start, end) pairs, absolute offsets
  const unsigned long long* d_total;  // device-side match count
  unsigned long long cap;
  int64_t* out;                       // nslots int64 per match
  int nslots, nops;
  uint8_t ops[24];                    // kind | class << 2, pattern order
  uint8_t at[34];                     // slot -> item index
  uint32_t cls[4][8];                 // 256-bit membership sets of the (<= 4) classes
};

__global__ void __launch_bounds__(128) flat_caps_kernel(const __grid_constant__ FlatCapArgs a) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long nmatches = *a.d_total < a.cap ? *a.d_total : a.cap;
  if (i >= nmatches) return;
  const int64_t s = a.matches[2 * i] - a.base, e = a.matches[2 * i + 1] - a.base;
  int64_t* o = a.out + i * (unsigned long long)a.nslots;
  o[0] = s + a.base;
  o[1] = e + a.base;
  int64_t pos = s;
  for (int k = 0; k <= a.nops; k++) {
    for (int sl = 2; sl < a.nslots; sl++)
      if (a.at[sl] == k) o[sl] = pos + a.base;
    if (k == a.nops) break;
    const int kind = a.ops[k] & 3, c = a.ops[k] >> 2;
    auto in_class = [&](int64_t p) -> bool {
      if (p >= e) return false;
      const unsigned b = a.h[p];
      return (a.cls[c][b >> 5] >> (b & 31)) & 1u;
    };
    if (kind == 0) {
      pos++;  // one byte of the class (the match exists: it is there)
    } else if (kind == 3) {
      if (in_class(pos)) pos++;
    } else {
      while (in_class(pos)) pos++;  // C+ / C*: the whole run (forced greedy)
    }
  }
}

}  // namespace

// grid covers `cap` matches; lanes past the device-side count exit immediately
cudaError_t launch_flat_captures(const uint8_t* h, int64_t n, int64_t base, const int64_t* matches,
                                 const unsigned long long* d_total, unsigned long long cap, const FlatDev& f,
                                 const uint8_t* at, int nslots, int64_t* out, cudaStream_t stream) {
  if (cap == 0) return cudaSuccess;
  FlatCapArgs a;
  a.h = h;
  a.n = n;
  a.base = base;
  a.matches = matches;
  a.d_total = d_total;
  a.cap = cap;
  a.out = out;
  a.nslots = nslots;
  a.nops = f.fwd_nops;
  for (int k = 0; k < 24; k++) a.ops[k] = f.fwd_ops[k];
  for (int k = 0; k < 34; k++) a.at[k] = at[k];
  for (int c = 0; c < 4; c++) {
    for (int w = 0; w < 8; w++) a.cls[c][w] = 0u;
    if (c < f.nclasses)
      for (int r = 0; r < f.cls_nranges[c]; r++)
        for (unsigned b = f.cls_lo[c][r]; b <= f.cls_hi[c][r]; b++) a.cls[c][b >> 5] |= 1u << (b & 31);
  }
  const int threads = 128;
  const unsigned long long blocks = (cap + threads - 1) / threads;
  flat_caps_kernel<<<(unsigned)blocks, threads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace cgx
