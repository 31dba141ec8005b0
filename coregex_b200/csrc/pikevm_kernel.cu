// pikevm_kernel.cu — capture resolution for FindAllSubmatchIndex on sm_100a.
//
// Replaces reference nfa/pikevm.go:2186-2432 (SearchWithSlotTableCapturesAt and friends) for the
// batch path: the scan kernel has already produced the leftmost-first match list, so instead of
// re-seeding a thread at every byte (the reference's unanchored loop, :2240-2256) each GPU lane
// takes ONE match and runs the anchored Pike simulation from its start with capture slots:
// thread lists in DFS priority order, first-arrival-wins de-duplication, break at the first Match
// thread of a generation (leftmost-first), slot save/restore frames during the epsilon closure
// (reference :1895-2005).  The highest-priority path from a match start is the same path the
// reference's unanchored search commits to, so the slots are identical.
#include <cstdint>
#include <cuda_runtime.h>

#include "scan_params.h"

namespace cgx {

namespace {

constexpr int MAXS = 16;    // capture slots (2 per group, 8 groups)

// Two sizes of the same kernel (the host proves which one a program fits, host/pike_pack.cpp):
//   small: <= 64 instructions (one visited word), <= 32 live threads, 96 closure frames
//   large: <= 512 instructions, <= 64 live threads, 520 closure frames — patterns whose groups hold
//          `.`, `\S`, negated or non-ASCII classes (UTF-8 byte automata), at ~13 KB of local memory a lane
struct Small {
  static constexpr int MAXT = 32, VW = 1, MAXSTK = 96;
  using Pc = uint8_t;
};
struct Large {
  static constexpr int MAXT = 64, VW = 8, MAXSTK = 520;
  using Pc = uint16_t;
};

template <class Z>
struct ThreadList {
  typename Z::Pc pc[Z::MAXT];
  int32_t slots[Z::MAXT][MAXS];  // relative to the match start; -1 = unset
  int n;
};

template <class Z>
struct Visited {
  unsigned long long w[Z::VW];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < Z::VW; k++) w[k] = 0ull;
  }
  // true when pc was already marked; marks it
  __device__ __forceinline__ bool test_and_set(int pc) {
    unsigned long long& x = w[Z::VW == 1 ? 0 : pc >> 6];
    const unsigned long long bit = 1ull << (pc & 63);
    const bool was = (x & bit) != 0;
    x |= bit;
    return was;
  }
};

__device__ __forceinline__ bool is_word(int b) {
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}

// reference nfa/pikevm.go:1646-1675 checkLookAssertion
__device__ __forceinline__ bool look_ok(int kind, const uint8_t* h, int64_t n, int64_t pos, bool text0) {
  const int prev = pos > 0 ? (int)h[pos - 1] : -1;
  const int next = pos < n ? (int)h[pos] : -1;
  switch (kind) {
    case 0: return pos == 0 && text0;  // a shard with base > 0 starts after a delimiter, not at \\A
    case 1: return pos == n;
    case 2: return pos == 0 || prev == '\n';
    case 3: return pos == n || next == '\n';
    case 4: return is_word(prev) != is_word(next);
    default: return is_word(prev) == is_word(next);
  }
}

struct Vm {
  const uint32_t* code;
  const uint32_t* sets;
  const uint8_t* h;
  int64_t n;
  int64_t s;  // match start
  int nslots;
  bool text0;  // position 0 of h is the start of the logical haystack
};

// epsilon closure of pc at position pos (relative rp = pos - s), appending to `tl`
template <class Z>
__device__ void add_thread(const Vm& vm, ThreadList<Z>& tl, Visited<Z>& visited, int pc0, int64_t pos,
                           int32_t* cur) {
  constexpr int MAXT = Z::MAXT, MAXSTK = Z::MAXSTK;
  // frame: low 16 bits pc (0xFFFF = restore frame), then slot and old value
  int32_t stk_a[MAXSTK];
  int32_t stk_b[MAXSTK];
  int sp = 0;
  stk_a[sp] = pc0;
  stk_b[sp++] = 0;
  while (sp > 0) {
    --sp;
    const int32_t fa = stk_a[sp], fb = stk_b[sp];
    if (fa < 0) {  // restore frame: slot = -fa-1
      cur[-fa - 1] = fb;
      continue;
    }
    const int pc = fa;
    if (pc == 0xFFFF) continue;
    if (visited.test_and_set(pc)) continue;
    const uint32_t w0 = vm.code[2 * pc], w1 = vm.code[2 * pc + 1];
    const int op = w0 & 255, arg = (int)(w0 >> 8);
    const int out = w1 & 0xFFFF, out1 = w1 >> 16;
    switch (op) {
      case 1:    // I_SET
      case 6: {  // I_MATCH
        if (tl.n < MAXT) {
          tl.pc[tl.n] = (typename Z::Pc)pc;
          for (int k = 0; k < vm.nslots; k++) tl.slots[tl.n][k] = cur[k];
          tl.n++;
        }
        break;
      }
      case 2:  // I_SPLIT: out preferred
        if (sp + 2 <= MAXSTK) {
          stk_a[sp] = out1; stk_b[sp++] = 0;
          stk_a[sp] = out; stk_b[sp++] = 0;
        }
        break;
      case 3:  // I_SAVE
        if (arg < vm.nslots && sp + 2 <= MAXSTK) {
          stk_a[sp] = -arg - 1; stk_b[sp++] = cur[arg];
          cur[arg] = (int32_t)(pos - vm.s);
          stk_a[sp] = out; stk_b[sp++] = 0;
        } else if (sp + 1 <= MAXSTK) {
          stk_a[sp] = out; stk_b[sp++] = 0;
        }
        break;
      case 4:  // I_ASSERT
        if (look_ok(arg, vm.h, vm.n, pos, vm.text0) && sp + 1 <= MAXSTK) {
          stk_a[sp] = out; stk_b[sp++] = 0;
        }
        break;
      case 5:  // I_NOP
        if (sp + 1 <= MAXSTK) {
          stk_a[sp] = out; stk_b[sp++] = 0;
        }
        break;
      default:
        break;
    }
  }
}

// nmatches = min(*d_total, cap): the match count is only known on the device
template <class Z>
__global__ void pike_captures_kernel(const uint8_t* h, int64_t n, int64_t base, const int64_t* matches,
                                     const unsigned long long* d_total, unsigned long long cap,
                                     const uint32_t* code, const uint32_t* sets, int start_pc, int nslots,
                                     int64_t* out, bool text_end) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long nmatches = *d_total < cap ? *d_total : cap;
  if (i >= nmatches) return;
  Vm vm{code, sets, h, n, matches[2 * i] - base, nslots, base == 0};
  if (text_end && vm.s == n) {
    // reference nfa/pikevm.go:2201-2206: a search that starts AT the end of the haystack answers
    // from matchesEmptyAt and builds its captures from no slots at all — the groups of the empty
    // match at the very end are reported unset (stdlib reports them as empty at n)
    int64_t* o = out + i * (unsigned long long)nslots;
    o[0] = o[1] = matches[2 * i];
    for (int k = 2; k < nslots; k++) o[k] = -1;
    return;
  }
  ThreadList<Z> a, b;
  ThreadList<Z>* cur = &a;
  ThreadList<Z>* nxt = &b;
  int32_t work[MAXS];
  for (int k = 0; k < MAXS; k++) work[k] = -1;
  Visited<Z> visited;
  visited.clear();
  cur->n = 0;
  add_thread(vm, *cur, visited, start_pc, vm.s, work);
  int64_t best_end = -1;
  int32_t best[MAXS];
  for (int k = 0; k < MAXS; k++) best[k] = -1;
  for (int64_t pos = vm.s;; pos++) {
    const int byte = pos < n ? (int)h[pos] : -1;
    visited.clear();
    nxt->n = 0;
    for (int t = 0; t < cur->n; t++) {
      const int pc = cur->pc[t];
      const uint32_t w0 = code[2 * pc], w1 = code[2 * pc + 1];
      if ((w0 & 255) == 6) {  // Match: leftmost-first -> lower-priority threads are cut
        best_end = pos;
        for (int k = 0; k < nslots; k++) best[k] = cur->slots[t][k];
        break;
      }
      if (byte >= 0) {
        const uint32_t* st = sets + 8 * (w0 >> 8);
        if ((st[byte >> 5] >> (byte & 31)) & 1u) {
          for (int k = 0; k < nslots; k++) work[k] = cur->slots[t][k];
          add_thread(vm, *nxt, visited, (int)(w1 & 0xFFFF), pos + 1, work);
        }
      }
    }
    if (byte < 0 || nxt->n == 0) break;
    ThreadList<Z>* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  int64_t* o = out + i * (unsigned long long)nslots;
  const int64_t s_abs = matches[2 * i];
  o[0] = s_abs;
  o[1] = best_end < 0 ? matches[2 * i + 1] : best_end + base;
  for (int k = 2; k < nslots; k += 2) {
    const bool set = best[k] >= 0 && best[k + 1] >= 0;  // reference :2411-2432
    o[k] = set ? s_abs + best[k] : -1;
    o[k + 1] = set ? s_abs + best[k + 1] : -1;
  }
}

}  // namespace

// grid covers `cap` matches; lanes past the device-side count exit immediately
cudaError_t launch_pike_captures(const uint8_t* h, int64_t n, int64_t base, const int64_t* matches,
                                 const unsigned long long* d_total, unsigned long long cap,
                                 const uint32_t* code, const uint32_t* sets, int start_pc, int nslots,
                                 int64_t* out, cudaStream_t stream, bool large, bool text_end) {
  if (cap == 0) return cudaSuccess;
  const int threads = 128;
  const unsigned long long blocks = (cap + threads - 1) / threads;
  if (large)
    pike_captures_kernel<Large><<<(unsigned)blocks, threads, 0, stream>>>(h, n, base, matches, d_total, cap, code,
                                                                          sets, start_pc, nslots, out, text_end);
  else
    pike_captures_kernel<Small><<<(unsigned)blocks, threads, 0, stream>>>(h, n, base, matches, d_total, cap, code,
                                                                          sets, start_pc, nslots, out, text_end);
  return cudaGetLastError();
}

}  // namespace cgx
