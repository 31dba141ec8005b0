// scan_dfa.cu — sm_100a kernel for the prefilter + anchored-DFA strategies.
//
// Replaces, for a whole corpus at once:
//   reference meta/findall.go:176-290        findAllIndicesLoop (pos = end chaining)
//   reference meta/find_indices.go:1050-1088 DigitPrefilter loop (candidate -> SearchAtAnchored)
//   reference simd/memchr_digit_amd64.s:26   memchrDigitAVX2 (32 B/iter digit scan)
//   reference dfa/lazy/lazy.go:219-324       SearchAtAnchored (per-byte class + table walk)
//
// Shape: a persistent grid; each CTA repeatedly takes a 32 KB chunk ticket, pulls the chunk
// (+16 B before, +1 KB after) into shared memory with one TMA bulk copy, then
//   phase A  every warp classifies 1 KB tiles with SWAR compares (32 B per lane -> one 32-bit
//            candidate word per lane) into a shared bitmap  — the memchr_digit / first-byte scan;
//   phase B  each warp owns the lines that START in its 4 KB slice, compacts candidates into
//            full batches of 32, walks the shared-memory DFA table one candidate per lane, and
//            resolves the leftmost non-overlapping chain exactly like the reference loop
//            (fast path: no overlaps in the batch; slow path: one lane replays the loop);
//   phase C  matches are staged in shared memory, the CTA publishes its count, a warp does a
//            decoupled look-back to get the global output offset, and the staged (start,end)
//            pairs are written as int64 in global match order.
// Every corpus byte crosses HBM once; output is 16 B per match.
#include <cstdio>

#include "scan_common.cuh"
#include "scan_params.h"

namespace cgx {

namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int CH = 32768;         // chunk bytes owned by one CTA iteration
constexpr int OVER = 1024;        // bytes after the chunk that are classified too
constexpr int PRE = 16;           // bytes before the chunk kept in the window
constexpr int WIN = PRE + CH + OVER;
constexpr int TILE = 1024;        // bytes per warp classification step (32 B per lane)
constexpr int NTILES = (CH + OVER) / TILE;
constexpr int NWORDS = NTILES * 32;
constexpr int SUB = CH / WARPS;   // slice whose line starts a warp owns
constexpr int GROUP = 16;         // bitmap words compacted per step
constexpr int QCAP = GROUP * 32 + 32;
constexpr int STG = 320;          // staged matches per warp
constexpr int64_t INF = (int64_t)1 << 62;

struct Smem {
  uint64_t mbar;
  int64_t ls[WARPS + 1];
  unsigned wcount[WARPS];
  unsigned woverflow;
  unsigned long long cta_base;
  unsigned chunk;
  uint32_t cand[NWORDS];
  uint16_t queue[WARPS][QCAP];
  uint2 stage[WARPS][STG];
  alignas(128) uint8_t win[WIN];
  // followed by: uint16_t trans[nstates*256]; uint8_t eoi[nstates]; uint8_t lut[256]
};

struct Ctx {
  const ScanArgs& a;
  Smem& sm;
  const uint16_t* trans;  // shared
  const uint8_t* eoi;     // shared
  const uint8_t* lut;     // shared (F_LUT) or null
  int64_t cbeg;           // chunk begin (global)
  int64_t gw;             // global position of win[0]
  int lane, warp;
};

__device__ __forceinline__ uint8_t byte_at(const Ctx& c, int64_t p) {
  int64_t i = p - c.gw;
  if (i >= 0 && i < WIN) return c.sm.win[i];
  return __ldg(c.a.h + p);
}

__device__ __forceinline__ bool in_filter_set(const Ctx& c, uint8_t b) {
  if (c.a.filter.kind == F_LUT) return c.lut[b] != 0;
  bool r = false;
  for (int k = 0; k < c.a.filter.nranges; k++) r |= (b >= c.a.filter.lo[k] && b <= c.a.filter.hi[k]);
  return r;
}

__device__ __forceinline__ int start_kind(uint8_t b) {
  if (b == '\n') return 3;
  if (b == '\r') return 4;
  bool w = (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
  return w ? 1 : 0;
}

// Anchored leftmost-first walk from p0.  Returns the match end or -1.
__device__ __forceinline__ int64_t dfa_walk(const Ctx& c, int64_t p0) {
  unsigned s = c.a.dfa.start[0];
  if (c.a.dfa.kind_lut_needed) {
    int k = p0 == 0 ? 2 : start_kind(byte_at(c, p0 - 1));
    s = c.a.dfa.start[k];
  }
  int64_t last = -1;
  int64_t p = p0;
  const int64_t n = c.a.n;
  // fast loop while inside the shared window
  int64_t i = p - c.gw;
  int64_t wend = n - c.gw < WIN ? n - c.gw : WIN;
  while (s) {
    if (i >= wend) break;
    unsigned e = c.trans[(s << 8) + c.sm.win[i]];
    if (e & 0x8000u) last = c.gw + i;
    s = e & 0x7FFFu;
    i++;
  }
  p = c.gw + i;
  while (s) {  // beyond the window (long line) or at end of input
    if (p >= n) {
      if (c.eoi[s]) last = n;
      break;
    }
    unsigned e = c.trans[(s << 8) + __ldg(c.a.h + p)];
    if (e & 0x8000u) last = p;
    s = e & 0x7FFFu;
    p++;
  }
  return last;
}

// ---- phase A: candidate bitmap -------------------------------------------------------------------
__device__ __forceinline__ uint32_t classify32(const Ctx& c, const uint8_t* p32) {
  const uint4 v0 = *reinterpret_cast<const uint4*>(p32);
  const uint4 v1 = *reinterpret_cast<const uint4*>(p32 + 16);
  uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t m = 0;
  if (c.a.filter.kind == F_LUT) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint32_t x = w[k];
      uint32_t f = (c.lut[x & 255] ? 1u : 0u) | (c.lut[(x >> 8) & 255] ? 2u : 0u) |
                   (c.lut[(x >> 16) & 255] ? 4u : 0u) | (c.lut[x >> 24] ? 8u : 0u);
      m |= f << (4 * k);
    }
    return m;
  }
  const int nr = c.a.filter.nranges;
  for (int r = 0; r < nr; r++) {
    const uint32_t klo = swar_klo(c.a.filter.lo[r]), khi = swar_khi(c.a.filter.hi[r]);
#pragma unroll
    for (int k = 0; k < 8; k++) m |= pack4(swar_in_range(w[k], klo, khi)) << (4 * k);
  }
  return m;
}

__device__ void phase_a(const Ctx& c) {
  for (int t = c.warp; t < NTILES; t += WARPS) {
    const int rel = t * TILE + c.lane * 32;  // relative to cbeg
    const uint8_t* p = c.sm.win + PRE + rel;
    uint32_t m = classify32(c, p);
    if (c.a.filter.kind == F_RUNSTART) {
      uint32_t prev = in_filter_set(c, p[-1]) ? 1u : 0u;
      m = m & ~((m << 1) | prev);
    }
    // positions at or beyond n are never candidates
    int64_t gp = c.cbeg + rel;
    if (gp + 32 > c.a.n) {
      int64_t v = c.a.n - gp;
      m = v <= 0 ? 0u : (m & ((1u << v) - 1u));
    }
    c.sm.cand[t * 32 + c.lane] = m;
  }
}

// first position q >= from with q == 0 or byte(q-1) == delim, searched inside the window only
__device__ int64_t find_line_start(const Ctx& c, int64_t from) {
  if (from <= 0) return 0;
  const int64_t wend_g = c.gw + WIN < c.a.n ? c.gw + WIN : c.a.n;  // bytes valid in window
  for (int64_t q = from - 1 + c.lane; __any_sync(0xffffffffu, q < wend_g); q += 32) {
    bool hit = q < wend_g && c.sm.win[q - c.gw] == c.a.delim;
    unsigned b = __ballot_sync(0xffffffffu, hit);
    if (b) return q - c.lane + (__ffs(b) - 1) + 1;
  }
  // the haystack end also terminates the last line
  return INF;
}

// ---- phase B: verify + chain ---------------------------------------------------------------------
template <bool DIRECT>
struct Emitter {
  const Ctx& c;
  unsigned nkept = 0;        // warp-uniform
  bool overflow = false;     // warp-uniform (STAGE mode)
  unsigned long long goff;   // DIRECT mode: global index of this warp's first match
  __device__ Emitter(const Ctx& cc, unsigned long long g) : c(cc), goff(g) {}

  __device__ __forceinline__ void put(unsigned idx, int64_t s, int64_t e) {
    if (c.a.mode != M_FINDALL) return;
    if (DIRECT) {
      unsigned long long gi = goff + idx;
      if ((int64_t)gi < c.a.cap) {
        longlong2 v = make_longlong2(s + c.a.base, e + c.a.base);
        *reinterpret_cast<longlong2*>(c.a.out + 2 * gi) = v;
      }
    } else {
      if (idx < STG) c.sm.stage[c.warp][idx] = make_uint2((unsigned)(s - c.cbeg), (unsigned)(e - c.cbeg));
    }
  }
};

// One lane replays the reference loop from `pos` while the next candidate is <= last_cand
// (or, with to_line_end, until the current line ends).  Returns the new chain position.
template <bool DIRECT>
__device__ int64_t serial_chain(const Ctx& c, Emitter<DIRECT>& em, int64_t pos, int64_t last_cand,
                                bool to_line_end) {
  unsigned added = 0;
  if (c.lane == 0) {
    const int64_t n = c.a.n;
    while (pos < n) {
      // reference: digitPos = prefilter.Find(haystack, pos)
      int64_t d = pos;
      bool stop = false;
      while (d < n) {
        uint8_t b = byte_at(c, d);
        if (in_filter_set(c, b)) break;
        if (to_line_end && b == c.a.delim) {
          stop = true;
          break;
        }
        d++;
      }
      if (stop || d >= n) {
        pos = d;
        break;
      }
      if (!to_line_end && d > last_cand) break;  // chain position unchanged
      int64_t e = dfa_walk(c, d);
      if (e >= 0) {
        em.put(em.nkept + added, d, e);
        added++;
        pos = e > d ? e : d + 1;
      } else {
        pos = d + 1;
        if (c.a.filter.kind == F_RUNSTART)
          while (pos < n && in_filter_set(c, byte_at(c, pos))) pos++;
      }
    }
  }
  added = __shfl_sync(0xffffffffu, added, 0);
  pos = __shfl_sync(0xffffffffu, pos, 0);
  if (!DIRECT && em.nkept + added > STG) em.overflow = true;
  em.nkept += added;
  return pos;
}

template <bool DIRECT>
__device__ int64_t process_batch(const Ctx& c, Emitter<DIRECT>& em, int64_t cand, bool valid,
                                 int64_t kept_end) {
  int64_t end = -1;
  if (valid) end = dfa_walk(c, cand);
  const bool ok = end >= 0;
  const unsigned okmask = __ballot_sync(0xffffffffu, ok);
  if (!okmask) return kept_end;
  if (c.a.mode == M_ISMATCH) {
    if (c.lane == 0) c.a.total[1] = 1ull;
    em.nkept += __popc(okmask);
    return kept_end;
  }
  const unsigned lower = okmask & ((1u << c.lane) - 1u);
  const int src = lower ? 31 - __clz(lower) : 0;
  int64_t pe = __shfl_sync(0xffffffffu, end, src);
  if (!lower) pe = kept_end;
  bool bad = ok && cand < pe;
  if (c.a.skip_safe && ok && end < c.a.n) bad |= in_filter_set(c, byte_at(c, end));
  if (__any_sync(0xffffffffu, bad)) {
    // replay from the chain position through the last candidate of this batch
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int hi = 31 - __clz(vmask);
    const int64_t last_cand = __shfl_sync(0xffffffffu, cand, hi);
    return serial_chain<DIRECT>(c, em, kept_end, last_cand, false);
  }
  if (ok) em.put(em.nkept + __popc(lower), cand, end);
  const unsigned add = __popc(okmask);
  if (!DIRECT && em.nkept + add > STG) em.overflow = true;
  em.nkept += add;
  const int top = 31 - __clz(okmask);
  return __shfl_sync(0xffffffffu, end, top);
}

template <bool DIRECT>
__device__ void phase_b(const Ctx& c, Emitter<DIRECT>& em) {
  const int64_t lo = c.sm.ls[c.warp];
  const int64_t hi = c.sm.ls[c.warp + 1];
  if (lo >= hi || lo >= c.a.n) return;
  const int64_t bm_end = c.cbeg + CH + OVER;  // bitmap covers [cbeg, bm_end)
  const int64_t hi_b = hi < bm_end ? hi : bm_end;
  int64_t kept_end = lo;
  uint16_t* q = c.sm.queue[c.warp];
  int qlen = 0;
  const int lo_rel = (int)(lo - c.cbeg);
  const int hi_rel = (int)(hi_b - c.cbeg);
  for (int wbase = lo_rel >> 5; wbase * 32 < hi_rel; wbase += GROUP) {
    const int widx = wbase + c.lane;
    uint32_t word = 0;
    if (c.lane < GROUP && widx < NWORDS && widx * 32 < hi_rel) {
      word = c.sm.cand[widx];
      const int b0 = widx * 32;
      if (b0 < lo_rel) word &= ~0u << (lo_rel - b0);
      if (b0 + 32 > hi_rel) word &= (1u << (hi_rel - b0)) - 1u;
    }
    int total;
    int off = qlen + warp_excl_scan(__popc(word), c.lane, total);
    while (word) {
      const int b = __ffs(word) - 1;
      word &= word - 1;
      q[off++] = (uint16_t)(widx * 32 + b);
    }
    qlen += total;
    __syncwarp();
    int head = 0;
    while (qlen - head >= 32) {
      const int64_t cand = c.cbeg + q[head + c.lane];
      kept_end = process_batch<DIRECT>(c, em, cand, true, kept_end);
      head += 32;
    }
    if (head) {
      const int rem = qlen - head;
      uint16_t tmp = c.lane < rem ? q[head + c.lane] : 0;
      __syncwarp();
      if (c.lane < rem) q[c.lane] = tmp;
      qlen = rem;
      __syncwarp();
    }
  }
  if (qlen) {
    const bool valid = c.lane < qlen;
    const int64_t cand = c.cbeg + (valid ? q[c.lane] : 0);
    kept_end = process_batch<DIRECT>(c, em, cand, valid, kept_end);
  }
  if (hi > bm_end) {
    // the last owned line runs past the classified window: finish it serially
    int64_t from = kept_end > bm_end ? kept_end : bm_end;
    serial_chain<DIRECT>(c, em, from, 0, true);
  }
}

__global__ void __launch_bounds__(THREADS, 2) scan_dfa_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  uint16_t* s_trans = reinterpret_cast<uint16_t*>(smem_raw + sizeof(Smem));
  uint8_t* s_eoi = reinterpret_cast<uint8_t*>(s_trans + (size_t)a.dfa.nstates * 256);
  uint8_t* s_lut = s_eoi + ((a.dfa.nstates + 15) & ~15);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // stage the automaton once per CTA
  for (int i = tid; i < a.dfa.nstates * 128; i += THREADS)
    reinterpret_cast<uint32_t*>(s_trans)[i] = reinterpret_cast<const uint32_t*>(a.dfa.trans)[i];
  for (int i = tid; i < a.dfa.nstates; i += THREADS) s_eoi[i] = a.dfa.eoi[i];
  if (a.filter.kind == F_LUT)
    for (int i = tid; i < 256; i += THREADS) s_lut[i] = a.filter.lut[i];
  if (tid == 0) {
    mbar_init(&sm.mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t parity = 0;
  for (;;) {
    if (tid == 0) sm.chunk = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const int64_t chunk = sm.chunk;
    if (chunk >= a.nchunks) break;
    if (a.mode == M_ISMATCH && *((volatile unsigned long long*)&a.total[1])) break;
    const int64_t cbeg = chunk * CH;
    const int64_t gw = cbeg - PRE;
    const int64_t lo_g = gw < 0 ? 0 : gw;
    const int64_t hi_g = gw + WIN < a.n ? gw + WIN : a.n;
    const uint32_t bytes = (uint32_t)(hi_g - lo_g);
    const uint32_t bulk = bytes & ~15u;
    if (tid == 0) {
      if (bulk) {
        mbar_expect_tx(&sm.mbar, bulk);
        tma_load_1d(sm.win + (lo_g - gw), a.h + lo_g, bulk, &sm.mbar);
      } else {
        mbar_arrive(&sm.mbar);
      }
    }
    // bytes the bulk copy does not cover: before position 0, the <16 B tail, past the end
    for (int i = tid; i < WIN; i += THREADS) {
      const int64_t g = gw + i;
      if (g < 0 || g >= lo_g + bulk) sm.win[i] = g >= 0 && g < a.n ? a.h[g] : a.delim;
    }
    mbar_wait(&sm.mbar, parity);
    parity ^= 1;
    __syncthreads();

    Ctx c{a, sm, s_trans, s_eoi, s_lut, cbeg, gw, lane, warp};
    phase_a(c);
    {
      int64_t s = find_line_start(c, cbeg + (int64_t)warp * SUB);
      if (lane == 0) sm.ls[warp] = s;
      if (warp == WARPS - 1) {
        int64_t e = find_line_start(c, cbeg + CH);
        if (lane == 0) sm.ls[WARPS] = e;
      }
      if (tid == 0) sm.woverflow = 0;
    }
    __syncthreads();

    Emitter<false> em(c, 0);
    phase_b<false>(c, em);
    if (lane == 0) {
      sm.wcount[warp] = em.nkept;
      if (em.overflow) atomicOr(&sm.woverflow, 1u);
    }
    __syncthreads();

    if (a.mode == M_FINDALL) {
      if (warp == 0) {
        unsigned agg = 0;
        for (int w = 0; w < WARPS; w++) agg += sm.wcount[w];
        unsigned long long excl = 0;
        if (chunk == 0) {
          if (lane == 0) st_release(&a.status[0], LB_PREFIX | agg);
        } else {
          if (lane == 0) st_release(&a.status[chunk], LB_AGG | agg);
          int64_t look = chunk - 1;
          for (;;) {
            const int64_t idx = look - lane;
            unsigned long long v = LB_PREFIX;  // lanes before chunk 0 act as a zero prefix
            if (idx >= 0) {
              do {
                v = ld_acquire(&a.status[idx]);
              } while ((v >> 62) == 0);
            }
            const unsigned pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
            // sum values of lanes up to and including the first PREFIX lane
            const int first = pm ? __ffs(pm) - 1 : 32;
            unsigned long long val = lane <= first ? (v & LB_VALUE) : 0ull;
#pragma unroll
            for (int d = 16; d; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
            excl += val;
            if (pm) break;
            look -= 32;
          }
          if (lane == 0) st_release(&a.status[chunk], LB_PREFIX | (excl + agg));
        }
        if (lane == 0) {
          sm.cta_base = excl;
          if (chunk == a.nchunks - 1) a.total[0] = excl + agg;
        }
      }
      __syncthreads();
      unsigned wexcl = 0;
      for (int w = 0; w < warp; w++) wexcl += sm.wcount[w];
      const unsigned long long gbase = sm.cta_base + wexcl;
      if (!sm.woverflow) {
        const unsigned cnt = sm.wcount[warp];
        for (unsigned i = lane; i < cnt; i += 32) {
          const unsigned long long gi = gbase + i;
          if ((int64_t)gi < a.cap) {
            const uint2 m = sm.stage[warp][i];
            longlong2 v = make_longlong2(cbeg + m.x + a.base, cbeg + m.y + a.base);
            *reinterpret_cast<longlong2*>(a.out + 2 * gi) = v;
          }
        }
      } else {
        Emitter<true> em2(c, gbase);
        phase_b<true>(c, em2);
      }
    } else {
      if (tid == 0) {
        unsigned agg = 0;
        for (int w = 0; w < WARPS; w++) agg += sm.wcount[w];
        if (agg) atomicAdd(a.total, (unsigned long long)agg);
      }
    }
    __syncthreads();  // window, bitmap and staging are reused by the next chunk
  }
}

}  // namespace

size_t scan_dfa_smem_bytes(int nstates) {
  return sizeof(Smem) + (size_t)nstates * 512 + ((nstates + 15) & ~15) + 256;
}

int64_t scan_dfa_chunks(int64_t n) { return n <= 0 ? 0 : (n + CH - 1) / CH; }

// Launches the scan on `stream`.  ticket/status/total must be zeroed by the caller.
cudaError_t launch_scan_dfa(const ScanArgs& a, int sm_count, cudaStream_t stream) {
  if (a.nchunks == 0) return cudaSuccess;
  const size_t smem = scan_dfa_smem_bytes(a.dfa.nstates);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(scan_dfa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_dfa_kernel, THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  int64_t grid = (int64_t)sm_count * per_sm;
  if (grid > a.nchunks) grid = a.nchunks;
  scan_dfa_kernel<<<(unsigned)grid, THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace cgx
