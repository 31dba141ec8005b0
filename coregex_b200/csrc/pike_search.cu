// pike_search.cu — PikeVM search kernel: the fallback engine for patterns whose automaton does not
// fit the table kernels (anchored DFA over 160 states) — what the reference runs when it selects
// UseNFA or when its lazy DFA gives up (reference meta/find_indices.go:1172 findIndicesNFAAtWithState,
// nfa/pikevm.go:1711 SearchWithSlotTableAt -> :1747 searchWithSlotTableUnanchored, :1895
// addSearchThread, :2009 stepSearchThread, :2066 addSearchThreadToNext).
//
// The active-state step is the reference's, thread for thread: lists in priority (DFS) order,
// first arrival wins (sparse-set membership = one bit per instruction), a Match thread cuts
// everything of lower priority, a new thread is seeded at every position until a match has been
// seen (leftmost), and a search thread carries nothing but its start position (reference
// `searchThread{state, startPos}`).  What changes is the batch shape: the haystack is cut into
// records at the pattern's delimiter byte (a byte no match can contain), each GPU lane owns the
// records that START inside its slice of the input and runs the reference's FindAll loop
// (meta/findall.go:176-290: search, emit, continue at the match end) on them.  Matches are written
// in global order by a count pass, a prefix sum over the slices and an emit pass.
// This engine is about coverage, not speed: it is latency-bound (per-lane thread lists live in
// local memory, bytes come through L1/L2), one to two orders of magnitude below the table kernels.
#include <cuda_runtime.h>

#include <cstdint>

#include "scan_params.h"

namespace cgx {

namespace {

constexpr int SLICE = 1024;  // bytes of input whose record starts one lane owns

struct PikeArgs {
  const uint8_t* h;
  int64_t n, base, after;
  const uint32_t* code;  // 2 words per instruction (host/pike_pack.h)
  const uint32_t* sets;  // 8 words per byte set
  int ninst, start_pc;
  int delim;  // record delimiter byte, or 256: the whole haystack is one record
  int64_t nslices;
  unsigned* counts;             // per slice (count pass out, then exclusive offsets within the block)
  unsigned long long* blocksum; // per block of slices: total, then exclusive prefix
  int64_t* out;
  int64_t cap;
  unsigned long long* total;    // [0] matches, [1] flag
};

__device__ __forceinline__ bool is_word(int b) {
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}
// reference nfa/pikevm.go:1646-1675 checkLookAssertion; bytes outside the buffer: a shard with
// base > 0 begins right after a delimiter, one with after > 0 ends with one
__device__ __forceinline__ bool look_ok(const PikeArgs& a, int kind, int64_t pos) {
  const int prev = pos > 0 ? (int)__ldg(a.h + pos - 1) : (a.base > 0 ? (int)a.delim : -1);
  const int next = pos < a.n ? (int)__ldg(a.h + pos) : (a.after > 0 ? (int)a.delim : -1);
  switch (kind) {
    case 0: return prev < 0;
    case 1: return next < 0;
    case 2: return prev < 0 || prev == '\n';
    case 3: return next < 0 || next == '\n';
    case 4: return is_word(prev) != is_word(next);
    default: return is_word(prev) == is_word(next);
  }
}

template <int NT, int NI>
struct Lane {
  uint16_t pc[2][NT];
  int32_t st[2][NT];   // start of the thread's match, relative to the record start
  int n[2];
  uint32_t vis[NI / 32];
  uint16_t stk[NI];
};

// epsilon closure of pc0 at position pos, appended to list `l` in priority order
// (reference nfa/pikevm.go:1895-2005 addSearchThread; capture instructions are plain epsilons here)
template <int NT, int NI>
__device__ void add_thread(const PikeArgs& a, Lane<NT, NI>& L, int l, int pc0, int64_t pos, int32_t start) {
  int sp = 0;
  L.stk[sp++] = (uint16_t)pc0;
  while (sp > 0) {
    const int pc = L.stk[--sp];
    if (pc == 0xFFFF) continue;
    if ((L.vis[pc >> 5] >> (pc & 31)) & 1u) continue;
    L.vis[pc >> 5] |= 1u << (pc & 31);
    const uint32_t w0 = __ldg(a.code + 2 * pc), w1 = __ldg(a.code + 2 * pc + 1);
    const int op = w0 & 255, arg = (int)(w0 >> 8);
    const int out = w1 & 0xFFFF, out1 = w1 >> 16;
    switch (op) {
      case 1:    // I_SET
      case 6: {  // I_MATCH
        const int k = L.n[l];
        if (k < NT) {
          L.pc[l][k] = (uint16_t)pc;
          L.st[l][k] = start;
          L.n[l] = k + 1;
        }
        break;
      }
      case 2:  // I_SPLIT: out preferred -> popped first
        if (sp + 2 <= NI) {
          L.stk[sp++] = (uint16_t)out1;
          L.stk[sp++] = (uint16_t)out;
        }
        break;
      case 4:  // I_ASSERT
        if (look_ok(a, arg, pos) && sp < NI) L.stk[sp++] = (uint16_t)out;
        break;
      case 3:  // I_SAVE
      case 5:  // I_NOP
        if (sp < NI) L.stk[sp++] = (uint16_t)out;
        break;
      default:
        break;
    }
  }
}

// leftmost-first match at or after `at`, not beyond the record end `rend` (position of the
// record's delimiter, or n).  reference nfa/pikevm.go:1747-1829.
template <int NT, int NI>
__device__ bool search(const PikeArgs& a, Lane<NT, NI>& L, int64_t rs, int64_t at, int64_t rend, int64_t& ms, int64_t& me) {
  bool matched = false;
  int c = 0;
  L.n[0] = L.n[1] = 0;
  for (int w = 0; w < NI / 32; w++) L.vis[w] = 0u;
  for (int64_t p = at;; p++) {
    // a new thread at the lowest priority while nothing has matched (the membership bits still
    // describe list c: it was built as the "next" list of the step before)
    if (!matched) add_thread(a, L, c, a.start_pc, p, (int32_t)(p - rs));
    if (L.n[c] == 0) {
      // nothing alive: done once something has matched or the record is exhausted; otherwise the
      // seed's closure was empty here (a look-around that fails at p, e.g. a leading `\b`) and the
      // search moves on (reference nfa/pikevm.go:1747-1829 keeps seeding while nothing has matched)
      if (matched || p >= rend) break;
      for (int w = 0; w < NI / 32; w++) L.vis[w] = 0u;
      continue;
    }
    const int byte = p < a.n ? (int)__ldg(a.h + p) : -1;
    const int nx = c ^ 1;
    L.n[nx] = 0;
    for (int w = 0; w < NI / 32; w++) L.vis[w] = 0u;
    for (int t = 0; t < L.n[c]; t++) {
      const int pc = L.pc[c][t];
      const uint32_t w0 = __ldg(a.code + 2 * pc);
      if ((w0 & 255) == 6) {  // Match: leftmost-first -> lower-priority threads are cut
        matched = true;
        ms = rs + L.st[c][t];
        me = p;
        break;
      }
      if (byte >= 0) {
        const uint32_t* st = a.sets + 8 * (w0 >> 8);
        if ((__ldg(st + (byte >> 5)) >> (byte & 31)) & 1u)
          add_thread(a, L, nx, (int)(__ldg(a.code + 2 * pc + 1) & 0xFFFF), p + 1, L.st[c][t]);
      }
    }
    c = nx;
    if (p >= rend) {
      // the delimiter (or the end of input) consumed: nothing survives it; a Match thread that
      // became reachable by it (e.g. `$`) sits in the new list at position p + 1 only if the byte
      // was consumed, which no set allows — so the lists are done
      break;
    }
  }
  return matched;
}

template <int NT, int NI, bool EMIT>
__global__ void __launch_bounds__(128) pike_search_kernel(const PikeArgs a) {
  const int64_t slice = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ unsigned s_scan[128];
  unsigned cnt = 0;
  unsigned long long idx = 0;
  if (EMIT && slice < a.nslices) idx = a.blocksum[blockIdx.x] + a.counts[slice];
  if (slice < a.nslices) {
    Lane<NT, NI> L;
    const int64_t lo = slice * SLICE, hi = lo + SLICE < a.n ? lo + SLICE : a.n;
    // first record that starts in [lo, hi)
    int64_t rs = lo;
    if (lo > 0 && (int)__ldg(a.h + lo - 1) != a.delim) {
      while (rs < a.n && (int)__ldg(a.h + rs) != a.delim) rs++;
      rs++;
    }
    // (the empty record after a trailing delimiter — or an empty haystack — starts at n: it belongs
    // to the last slice; only a pattern that matches the empty string finds anything there).  In a
    // shard that other bytes follow (after > 0) position n is the first record of the NEXT shard.
    const bool last_slice = slice == a.nslices - 1 && a.after == 0;
    while (rs < hi || (last_slice && rs == a.n)) {
      int64_t rend = rs;
      if (a.delim > 255) rend = a.n;
      while (rend < a.n && (int)__ldg(a.h + rend) != a.delim) rend++;
      // the reference's FindAll loop (meta/findall.go:221-290) on this record.  `last` = end of the
      // previous non-empty match: an empty match right there is skipped (:251-259).  It can never
      // equal a record start, so records are independent.
      int64_t pos = rs, last = -1;
      while (pos <= rend) {
        int64_t ms = -1, me = -1;
        if (!search<NT, NI>(a, L, rs, pos, rend, ms, me)) break;
        if (ms == me && ms == last) {
          pos++;
          continue;
        }
        if (EMIT) {
          if ((int64_t)idx < a.cap) {
            a.out[2 * idx] = ms + a.base;
            a.out[2 * idx + 1] = me + a.base;
          }
          idx++;
        }
        cnt++;
        if (ms != me) last = me;
        pos = ms == me ? me + 1 : (me > pos ? me : pos + 1);  // :270-279
      }
      rs = rend + 1;
      if (rend >= a.n) break;
    }
  }
  if (EMIT) return;
  // count pass: exclusive offsets within the block + the block's total
  s_scan[threadIdx.x] = cnt;
  __syncthreads();
  for (int d = 1; d < 128; d <<= 1) {
    const unsigned v = threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0u;
    __syncthreads();
    s_scan[threadIdx.x] += v;
    __syncthreads();
  }
  if (slice < a.nslices) a.counts[slice] = s_scan[threadIdx.x] - cnt;
  if (threadIdx.x == 127) a.blocksum[blockIdx.x] = s_scan[127];
}

// exclusive prefix over the block totals (one block; the number of blocks is n / 128 KiB)
__global__ void pike_blocksum_kernel(unsigned long long* blocksum, int64_t nblocks, unsigned long long* total) {
  __shared__ unsigned long long carry;
  __shared__ unsigned long long s[256];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < nblocks; b0 += 256) {
    const int64_t i = b0 + threadIdx.x;
    const unsigned long long v = i < nblocks ? blocksum[i] : 0ull;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
      const unsigned long long x = threadIdx.x >= d ? s[threadIdx.x - d] : 0ull;
      __syncthreads();
      s[threadIdx.x] += x;
      __syncthreads();
    }
    if (i < nblocks) blocksum[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 255) carry += s[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    total[0] = carry;
    total[1] = carry ? 1ull : 0ull;
  }
}

template <int NT, int NI>
cudaError_t run(const PikeArgs& a, bool emit, cudaStream_t st) {
  const unsigned blocks = (unsigned)((a.nslices + 127) / 128);
  pike_search_kernel<NT, NI, false><<<blocks, 128, 0, st>>>(a);
  pike_blocksum_kernel<<<1, 256, 0, st>>>(a.blocksum, (int64_t)blocks, a.total);
  if (emit) pike_search_kernel<NT, NI, true><<<blocks, 128, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

int64_t pike_search_slices(int64_t n) { return n > 0 ? (n + SLICE - 1) / SLICE : 1; }  // an empty haystack is one (empty) record
// scratch: counts[nslices] u32, then blocksum[ceil(nslices / 128)] u64 (8-byte aligned)
size_t pike_search_scratch_bytes(int64_t n) {
  const int64_t ns = pike_search_slices(n);
  return (size_t)((ns * 4 + 7) & ~(int64_t)7) + (size_t)((ns + 127) / 128 + 1) * 8;
}

// mode: ScanMode.  total: device u64[2] {matches, flag}.  Returns the number of kernels launched in *launches.
cudaError_t launch_pike_search(const uint8_t* h, int64_t n, int64_t base, int64_t after, const uint32_t* code,
                               const uint32_t* sets, int ninst, int nthreads, int start_pc, int delim, int mode,
                               int64_t* out, int64_t cap, void* scratch, unsigned long long* total, cudaStream_t st,
                               int* launches) {
  PikeArgs a;
  a.h = h;
  a.n = n;
  a.base = base;
  a.after = after;
  a.code = code;
  a.sets = sets;
  a.ninst = ninst;
  a.start_pc = start_pc;
  a.delim = delim;
  a.nslices = pike_search_slices(n);
  a.counts = (unsigned*)scratch;
  a.blocksum = (unsigned long long*)((char*)scratch + ((a.nslices * 4 + 7) & ~(int64_t)7));
  a.out = out;
  a.cap = cap;
  a.total = total;
  if (launches) *launches = 0;
  const bool emit = mode == M_FINDALL && out && cap > 0;
  if (launches) *launches = emit ? 3 : 2;
  if (ninst <= 128 && nthreads <= 64) return run<64, 128>(a, emit, st);
  if (ninst <= 512 && nthreads <= 256) return run<256, 512>(a, emit, st);
  return run<1024, 2048>(a, emit, st);
}

}  // namespace cgx
