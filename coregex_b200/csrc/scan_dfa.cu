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
//            word per lane and byte class).  For flat patterns the class words feed a
//            bit-parallel right-to-left evaluation of the whole pattern (warp-wide carry chains
//            through ballots) that leaves only positions a match can really start at; otherwise
//            the first-byte / run-start set is the candidate set  — the prefilter;
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

constexpr int WARPS = 8;                 // compute warps
constexpr int CTHREADS = WARPS * 32;
constexpr int THREADS = CTHREADS + 32;   // + one writer warp (look-back and ordered output)
// named barriers (0 is __syncthreads)
constexpr int BAR_COMPUTE = 1, BAR_FULL = 2 /*+buf*/, BAR_EMPTY = 4 /*+buf*/, BAR_OVF = 6;
constexpr int CH = 31 * 1024;     // chunk bytes owned by one CTA iteration (CH+OVER = 32 tiles)
constexpr int OVER = 1024;        // bytes after the chunk that are classified too
constexpr int PRE = 16;           // bytes before the chunk kept in the window
constexpr int WIN = PRE + CH + OVER;
constexpr int WINPAD = 16;        // sentinel bytes after the window (always the delimiter)
constexpr int TILE = 1024;        // bytes per warp classification step (32 B per lane)
constexpr int NTILES = (CH + OVER) / TILE;
constexpr int NWORDS = NTILES * 32;
constexpr int SUB = CH / WARPS;   // slice whose line starts a warp owns
constexpr int GROUP = 16;         // bitmap words compacted per step
constexpr int QCAP = GROUP * 32 + 32;
constexpr int STG = 96;           // staged matches per warp and buffer (two buffers: deferred output)
constexpr int64_t INF = (int64_t)1 << 62;
constexpr uint32_t FULL = 0xffffffffu;

struct Smem {
  uint64_t mbar;
  int64_t ls[WARPS + 1];
  unsigned wcount[WARPS];
  unsigned woverflow2[2];
  unsigned chunk2[2];
  unsigned long long cta_base;
  unsigned chunk;
  alignas(16) uint32_t cand[NWORDS];
  uint16_t queue[WARPS][QCAP];
  uint2 stage[2][WARPS][STG];
  unsigned wcount2[2][WARPS];
  int64_t meta_chunk[2], meta_gw[2];   // written by the compute warps, read by the writer warp
  unsigned meta_agg[2], meta_ovf[2];
  alignas(16) uint32_t cmask[WARPS][8][32];  // per-lane class words of the tile a warp is evaluating
  FlatDev flat;                  // copy of the flat program (indexed constant-bank loads are slow)
  alignas(128) uint8_t win[WIN + WINPAD];
  // followed by: uint32_t trans[nstates*256]; uint8_t eoi[nstates]; uint8_t lut[256]
};

struct Ctx {
  const ScanArgs& a;
  Smem& sm;
  const uint32_t* trans;  // shared: (next_state * 1024) | (match_before << 31)
  const uint8_t* eoi;     // shared
  const uint8_t* lut;     // shared (F_LUT)
  int64_t cbeg;           // chunk begin (global)
  int64_t gw;             // global position of win[0]
  int wend;               // number of valid bytes in the window
  int lane, warp;
  int buf;                // staging buffer of this chunk
  // multi-literal engine tables (shared memory copies)
  const uint32_t* t_fp;
  const uint64_t* t_lit8;  // [npat] first 8 bytes, then [npat] byte masks
  const uint16_t* t_fp2;   // bucket masks of the third byte
  const uint8_t* t_bytes;
  const int32_t* t_offs;
  const uint16_t* t_order;
  const uint16_t* t_boff;
  // record engine tables (shared memory copies)
  const uint16_t* l_ut;
  const uint16_t* l_rt;
  const uint8_t* l_ueoi;
  const uint8_t* l_reoi;
};

__device__ __forceinline__ uint8_t byte_at(const Ctx& c, int64_t p) {
  const int64_t i = p - c.gw;
  if (i >= 0 && i < WIN) return c.sm.win[i];
  return __ldg(c.a.h + p);
}

__device__ __forceinline__ bool in_filter_set(const Ctx& c, uint8_t b) {
  if (c.a.filter.kind == F_LUT) return c.lut[b] != 0;
  if (c.a.filter.nranges == 1)
    return (unsigned)(b - c.a.filter.lo[0]) <= (unsigned)(c.a.filter.hi[0] - c.a.filter.lo[0]);
  bool r = false;
  for (int k = 0; k < c.a.filter.nranges; k++) r |= (b >= c.a.filter.lo[k] && b <= c.a.filter.hi[k]);
  return r;
}

__device__ __forceinline__ int start_kind(uint8_t b) {
  if (b == '\n') return 3;
  if (b == '\r') return 4;
  const bool w = (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
  return w ? 1 : 0;
}

// Anchored leftmost-first walk from global position p0 through global memory (any position).
__device__ int64_t dfa_walk_slow(const Ctx& c, int64_t p0) {
  unsigned s = c.a.dfa.start[0];
  if (c.a.dfa.kind_lut_needed) s = c.a.dfa.start[p0 == 0 ? (c.a.base == 0 ? 2 : start_kind(c.a.delim)) : start_kind(byte_at(c, p0 - 1))];
  int64_t last = -1, p = p0;
  const int64_t n = c.a.n;
  while (s) {
    if (p >= n) {
      if (c.eoi[s]) last = n;
      break;
    }
    const uint32_t e = c.trans[(s << 8) + byte_at(c, p)];
    if (e >> 31) last = p;
    s = (e & 0x7fffffffu) >> 10;
    p++;
  }
  return last;
}

// Anchored walk from window index i0 (global p0 = gw + i0).  Returns the match end as a window
// index (may exceed WIN when the slow path took over) or -1.  The fast loop has no bounds check:
// win[wend .. wend+WINPAD) holds the delimiter, which sends every state to DEAD.
__device__ __forceinline__ int dfa_walk(const Ctx& c, int i0) {
  uint32_t sp = (uint32_t)c.a.dfa.start[0] << 10;
  if (c.a.dfa.kind_lut_needed) {
    const int k = (c.gw + i0 == 0) ? (c.a.base == 0 ? 2 : start_kind(c.a.delim)) : start_kind(c.sm.win[i0 - 1]);
    sp = (uint32_t)c.a.dfa.start[k] << 10;
  }
  int last = -1, i = i0;
  const char* tb = reinterpret_cast<const char*>(c.trans);
  // the next byte is fetched one step ahead so that only the transition load is on the
  // loop-carried dependency chain (win[] is readable WINPAD bytes past the last valid byte)
  uint32_t b = c.sm.win[i];
  while (sp) {
    const uint32_t nb = c.sm.win[i + 1];
    const uint32_t e = *reinterpret_cast<const uint32_t*>(tb + sp + b * 4);
    if ((int)e < 0) last = i;
    sp = e & 0x7fffffffu;
    i++;
    b = nb;
  }
  if (i > c.wend) {
    // consumed a sentinel: the walk left the window (long line) or hit end of input
    const int64_t e = dfa_walk_slow(c, c.gw + i0);
    return e < 0 ? -1 : (int)(e - c.gw);
  }
  return last;
}

__device__ __forceinline__ void load32(const uint8_t* p32, uint32_t (&w)[8]);

// ---- multi-literal (Teddy) engine -------------------------------------------------------------
// reference prefilter/teddy.go:491-521 (candidate = AND of per-position nibble masks, here folded
// into one byte table per position) and :532-550 (verifyBucket).
__device__ __forceinline__ uint32_t teddy_mask(const Ctx& c, uint32_t b0, uint32_t b1) {
  return (c.t_fp[b0] & 0xFFFFu) & (c.t_fp[b1] >> 16);
}

// bytes p .. p+7 of the haystack as a little-endian word (bytes at or beyond n read as anything:
// every comparison is guarded by p + len <= n)
__device__ __forceinline__ uint64_t load8(const Ctx& c, int64_t p) {
  const int64_t i = p - c.gw;
  if (i >= 0 && i + 12 <= WIN) {
    // three aligned words of the window, shifted into place
    const uint32_t* w = reinterpret_cast<const uint32_t*>(c.sm.win + (i & ~(int64_t)3));
    const uint32_t sh = ((uint32_t)i & 3u) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
  }
  uint64_t v = 0;
  for (int k = 0; k < 8; k++)
    if (p + k < c.a.n) v |= (uint64_t)byte_at(c, p + k) << (8 * k);
  return v;
}

// does literal `id` stand at p?  hay8 = load8(p): the first eight bytes are one masked comparison
__device__ __forceinline__ bool lit_equal(const Ctx& c, int64_t p, int id, uint64_t hay8, int& len) {
  const int o = c.t_offs[id];
  len = c.t_offs[id + 1] - o;
  if (p + len > c.a.n) return false;
  if ((hay8 ^ c.t_lit8[id]) & c.t_lit8[c.a.teddy.npat + id]) return false;
  for (int k = 8; k < len; k++)
    if (byte_at(c, p + k) != c.t_bytes[o + k]) return false;
  return true;
}

// SIMD-regime verify at global position p: buckets low -> high, insertion order inside a bucket
__device__ int64_t teddy_verify(const Ctx& c, int64_t p, uint32_t mask) {
  const uint64_t hay8 = load8(c, p);
  while (mask) {
    const int b = __ffs(mask) - 1;
    mask &= mask - 1;
    if (b >= c.a.teddy.nbuckets) break;
    for (int k = c.t_boff[b]; k < c.t_boff[b + 1]; k++) {
      int len;
      if (lit_equal(c, p, c.t_order[k], hay8, len)) return p + len;
    }
  }
  return -1;
}
// scalar-regime verify (haystack[start:] shorter than 16 bytes): plain literal-id order
// (reference prefilter/teddy.go:447-458 findMatchScalar)
__device__ int64_t teddy_verify_scalar(const Ctx& c, int64_t p) {
  const uint64_t hay8 = load8(c, p);
  for (int id = 0; id < c.a.teddy.npat; id++) {
    int len;
    if (lit_equal(c, p, id, hay8, len)) return p + len;
  }
  return -1;
}

// batch verify from window index i0; returns end window index or -1
__device__ __forceinline__ int teddy_walk(const Ctx& c, int i0) {
  // The last position of the window has its second byte outside it (win[wend] is the sentinel, not
  // haystack): a long record that the owner follows beyond its window can start a literal exactly
  // there, so that byte comes from the haystack itself.
  uint32_t b1 = c.sm.win[i0 + 1];
  if (i0 + 1 >= c.wend) {
    if (c.gw + i0 + 1 >= c.a.n) return -1;  // a literal has three bytes or more
    b1 = byte_at(c, c.gw + i0 + 1);
  }
  const uint32_t m = teddy_mask(c, c.sm.win[i0], b1);
  if (!m) return -1;
  const int64_t e = teddy_verify(c, c.gw + i0, m);
  return e < 0 ? -1 : (int)(e - c.gw);
}

__device__ void phase_a_teddy(const Ctx& c) {
  for (int t = c.warp; t < NTILES; t += WARPS) {
    const int rel = t * TILE + c.lane * 32;
    const uint8_t* p = c.sm.win + PRE + rel;
    uint32_t w[8];
    load32(p, w);
    uint32_t m = 0;
    uint32_t tprev = c.t_fp[w[0] & 255u];
    // bytes k + 1 and k + 2 of position k: from the registers, the last two from the window
    auto byte = [&](int i) -> uint32_t { return i < 32 ? ((w[i >> 2] >> (8 * (i & 3))) & 255u) : (uint32_t)p[i]; };
    uint32_t tn = c.t_fp[byte(1)];
    // the last piece of the window: the bytes after it are not here (sentinel) — its last two
    // positions pass on what is known, whoever continues beyond the window verifies them
    const bool edge = rel + 32 == CH + OVER;
    const bool use3 = c.a.teddy.use_fp2 != 0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
      uint32_t t2 = use3 ? (uint32_t)c.t_fp2[byte(k + 2)] : 0xFFFFu;
      if (k >= 30 && edge) t2 = 0xFFFFu;
      if (k == 31 && edge) tn = 0xFFFF0000u;
      if ((tprev & 0xFFFFu) & (tn >> 16) & t2) m |= 1u << k;  // three-byte fingerprint
      tprev = tn;
      tn = c.t_fp[byte(k + 2)];
    }
    const int64_t gp = c.cbeg + rel;  // a 3-byte fingerprint needs position + 2 < n
    if (gp + 34 > c.a.n) {
      const int64_t v = c.a.n - 2 - gp;
      m = v <= 0 ? 0u : (v >= 32 ? m : (m & ((1u << v) - 1u)));
    }
    c.sm.cand[t * 32 + c.lane] = m;
  }
}

// ---- phase A: candidate bitmap -------------------------------------------------------------------
__device__ __forceinline__ void load32(const uint8_t* p32, uint32_t (&w)[8]) {
  const uint4 v0 = *reinterpret_cast<const uint4*>(p32);
  const uint4 v1 = *reinterpret_cast<const uint4*>(p32 + 16);
  w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
  w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
}

__device__ __forceinline__ uint32_t classify32(const Ctx& c, const uint8_t* p32) {
  uint32_t w[8];
  load32(p32, w);
  uint32_t m = 0;
  if (c.a.filter.kind == F_LUT) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t x = w[k];
      const uint32_t f = (c.lut[x & 255] ? 1u : 0u) | (c.lut[(x >> 8) & 255] ? 2u : 0u) |
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

__device__ void phase_a_plain(const Ctx& c) {
  for (int t = c.warp; t < NTILES; t += WARPS) {
    const int rel = t * TILE + c.lane * 32;  // relative to cbeg
    const uint8_t* p = c.sm.win + PRE + rel;
    uint32_t m = classify32(c, p);
    if (c.a.filter.kind == F_RUNSTART) {
      const uint32_t prev = in_filter_set(c, p[-1]) ? 1u : 0u;
      m = m & ~((m << 1) | prev);
    }
    const int64_t gp = c.cbeg + rel;  // positions at or beyond n are never candidates
    if (gp + 32 > c.a.n) {
      const int64_t v = c.a.n - gp;
      m = v <= 0 ? 0u : (m & ((1u << v) - 1u));
    }
    c.sm.cand[t * 32 + c.lane] = m;
  }
}

// Class membership of 32 bytes as a BIT-REVERSED word: bit (31-b) <=> byte b is in class `cls`.
__device__ __forceinline__ uint32_t pack_rev(const uint32_t (&fl)[8]) {
  uint32_t acc[4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    acc[a] = __dp4a(fl[2 * a], 0x10204080u, 0u);          // bytes 0..3 of the group -> weights 128..16
    acc[a] = __dp4a(fl[2 * a + 1], 0x01020408u, acc[a]);  // bytes 4..7 -> weights 8..1
  }
  // acc[a] = 128 * (8-bit reversed mask of bytes 8a..8a+7)
  return (acc[0] << 17) | (acc[1] << 9) | (acc[2] << 1) | (acc[3] >> 7);
}

__device__ __noinline__ uint32_t class_mask_rev_generic(const FlatDev& f, int cls, uint32_t w0, uint32_t w1,
                                                        uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5,
                                                        uint32_t w6, uint32_t w7);

__device__ __forceinline__ uint32_t class_mask_rev(const FlatDev& f, int cls, const uint32_t (&w)[8]) {
  const int nr = f.cls_nranges[cls];
  if (nr == 1 && f.cls_mode[cls][0] == 0) {  // one XOR-alignable range: 3 ops per word
    uint32_t fl[8];
    const uint32_t k1 = f.cls_k1[cls][0], k2 = f.cls_k2[cls][0];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t x = w[k];
      const uint32_t z = ((x ^ k1) & 0x7F7F7F7Fu) + k2;  // bit7 set <=> (x^lo)&0x7f > width
      fl[k] = ~(z | x) & 0x80808080u;
    }
    return pack_rev(fl);
  }
  return class_mask_rev_generic(f, cls, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
}

__device__ __noinline__ uint32_t class_mask_rev_generic(const FlatDev& f, int cls, uint32_t w0, uint32_t w1,
                                                        uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5,
                                                        uint32_t w6, uint32_t w7) {
  const uint32_t w[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
  uint32_t fl[8];
  const int nr = f.cls_nranges[cls];
#pragma unroll
  for (int k = 0; k < 8; k++) fl[k] = 0;
  for (int r = 0; r < nr; r++) {
    const uint32_t k1 = f.cls_k1[cls][r], k2 = f.cls_k2[cls][r];
    if (f.cls_mode[cls][r] == 0) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const uint32_t x = w[k];
        const uint32_t z = ((x ^ k1) & 0x7F7F7F7Fu) + k2;
        fl[k] |= ~(z | x) & 0x80808080u;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) fl[k] |= swar_in_range(w[k], k1, k2);
    }
  }
  return pack_rev(fl);
}

// 2048-bit helpers over the warp for the flat evaluation: lane l holds the 64-bit word l of the
// tile's bit-reversed position set; word 0 = the LAST 64 bytes of the 2 KB tile, bit (63-b) of a
// word <=> byte b of its 64-byte piece.
__device__ __forceinline__ uint64_t shl1_64(uint64_t m, int lane) {
  // markers move one position towards lower addresses (higher bits); the bit entering word 0 (a
  // position past the tile) is 1: unknown territory is assumed to allow a match
  const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
  uint32_t dn = __shfl_up_sync(FULL, hi, 1);
  if (lane == 0) dn = FULL;
  return ((uint64_t)__funnelshift_l(lo, hi, 1) << 32) | __funnelshift_l(dn, lo, 1);
}
__device__ __forceinline__ uint64_t add2048(uint64_t s, uint64_t cc, int lane) {
  const uint64_t sum = s + cc;
  const uint32_t G = __ballot_sync(FULL, sum < s);
  const uint32_t P = __ballot_sync(FULL, sum == ~0ull);
  const uint32_t A = G | P;
  const uint32_t carries = A ^ G ^ (A + G);  // bit l = carry into word l
  return sum + ((carries >> lane) & 1u);
}

__device__ __forceinline__ uint64_t class_mask_rev64(const FlatDev& f, int cls, const uint32_t (&w0)[8],
                                                     const uint32_t (&w1)[8]) {
  // bytes 0..31 of the piece land in the high word, bytes 32..63 in the low word
  return ((uint64_t)class_mask_rev(f, cls, w0) << 32) | class_mask_rev(f, cls, w1);
}

constexpr int FTILE = 2048;                       // flat evaluation: 64 bytes per lane
constexpr int NFTILES = (CH + OVER) / FTILE;

static_assert(NFTILES == 2 * WARPS, "each scanning warp evaluates exactly two 2 KB tiles per chunk");

// classes 0 and 1 of one 64-byte piece (bit-reversed), classes 2,3 parked in shared memory
__device__ __forceinline__ void flat_classify(const FlatDev& f, int nclasses, const uint8_t* p, uint64_t* cmw,
                                              uint64_t& cm0, uint64_t& cm1, uint64_t& runset, const Ctx& c) {
  uint32_t w0[8], w1[8];
  load32(p, w0);
  load32(p + 32, w1);
  cm0 = class_mask_rev64(f, 0, w0, w1);
  cm1 = 0;
  if (nclasses > 1) cm1 = class_mask_rev64(f, 1, w0, w1);
  for (int k = 2; k < nclasses; k++) cmw[(k - 2) * 32] = class_mask_rev64(f, k, w0, w1);
  runset = cm0;
  if (c.a.filter.kind == F_RUNSTART && !f.first_is_filter) {
    // run starts are those of the FILTER set (ASCII digits for the reference's DigitPrefilter),
    // which differs from class 0 when the pattern starts with a strict subset such as [0-5]
    uint32_t dh = 0, dl = 0;
    for (int r = 0; r < c.a.filter.nranges; r++) {
      const uint32_t klo = swar_klo(c.a.filter.lo[r]), khi = swar_khi(c.a.filter.hi[r]);
      uint32_t mh = 0, ml = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        mh |= pack4(swar_in_range(w0[k], klo, khi)) << (4 * k);
        ml |= pack4(swar_in_range(w1[k], klo, khi)) << (4 * k);
      }
      dh |= __brev(mh);
      dl |= __brev(ml);
    }
    runset = ((uint64_t)dh << 32) | dl;
  }
}

__device__ __forceinline__ void flat_store(const Ctx& c, int t, int piece, uint64_t R) {
  uint32_t m0 = __brev((uint32_t)(R >> 32));  // bytes 0..31 of the piece
  uint32_t m1 = __brev((uint32_t)R);          // bytes 32..63
  const int64_t gp = c.cbeg + t * FTILE + piece * 64;
  if (gp + 64 > c.a.n) {
    const int64_t v0 = c.a.n - gp, v1 = c.a.n - gp - 32;
    m0 = v0 <= 0 ? 0u : (v0 >= 32 ? m0 : (m0 & ((1u << v0) - 1u)));
    m1 = v1 <= 0 ? 0u : (v1 >= 32 ? m1 : (m1 & ((1u << v1) - 1u)));
  }
  *reinterpret_cast<uint2*>(&c.sm.cand[t * 64 + piece * 2]) = make_uint2(m0, m1);
}

// The warp's two tiles (t = warp and warp + WARPS) are evaluated together: the two marker chains
// are independent, which gives the scheduler two dependency chains per warp to interleave.
__device__ void phase_a_flat(const Ctx& c) {
  const FlatDev& f = c.sm.flat;
  uint64_t* cmwa = reinterpret_cast<uint64_t*>(&c.sm.cmask[c.warp][0][0]) + c.lane;  // classes 2,3 tile a
  uint64_t* cmwb = cmwa + 64;                                                          // ... tile b
  const int nclasses = f.nclasses, rev_nops = f.rev_nops, init_cls = f.rev_init_class;
  const int ta = c.warp, tb = c.warp + WARPS;
  const int piece = 31 - c.lane;  // lane l holds the (31-l)-th 64-byte piece of a tile
  uint64_t a0, a1, ra, b0, b1, rb;
  flat_classify(f, nclasses, c.sm.win + PRE + ta * FTILE + piece * 64, cmwa, a0, a1, ra, c);
  flat_classify(f, nclasses, c.sm.win + PRE + tb * FTILE + piece * 64, cmwb, b0, b1, rb, c);
  // right-to-left evaluation: M = positions from which items k..end can match
  uint64_t Ma = init_cls == 0 ? a0 : init_cls == 1 ? a1 : cmwa[(init_cls - 2) * 32];
  uint64_t Mb = init_cls == 0 ? b0 : init_cls == 1 ? b1 : cmwb[(init_cls - 2) * 32];
  for (int k = 0; k < rev_nops; k++) {
    const uint32_t op = f.rev_ops[k];
    const uint32_t cls = op >> 2, kind = op & 3u;
    const uint64_t Ca = cls == 0 ? a0 : cls == 1 ? a1 : cmwa[(cls - 2) * 32];
    const uint64_t Cb = cls == 0 ? b0 : cls == 1 ? b1 : cmwb[(cls - 2) * 32];
    const uint64_t ua = shl1_64(Ma, c.lane) & Ca;  // byte in class and the rest matches after it
    const uint64_t ub = shl1_64(Mb, c.lane) & Cb;
    if (kind == 0) {
      Ma = ua;
      Mb = ub;
    } else if (kind == 3) {
      Ma |= ua;
      Mb |= ub;
    } else {
      // extend leftwards through the run; a second marker inside one run survives the carry of
      // the first as a 1 in the sum, so the markers themselves are OR-ed back
      const uint64_t pa = (~add2048(ua, Ca, c.lane) & Ca) | ua;
      const uint64_t pb = (~add2048(ub, Cb, c.lane) & Cb) | ub;
      Ma = kind == 1 ? pa : (Ma | pa);
      Mb = kind == 1 ? pb : (Mb | pb);
    }
  }
  if (c.a.filter.kind == F_RUNSTART) {
    uint32_t upa = __shfl_down_sync(FULL, (uint32_t)ra, 1) & 1u;  // lowest-address bit of the next piece
    uint32_t upb = __shfl_down_sync(FULL, (uint32_t)rb, 1) & 1u;
    if (c.lane == 31) {
      upa = in_filter_set(c, c.sm.win[PRE + ta * FTILE - 1]) ? 1u : 0u;
      upb = in_filter_set(c, c.sm.win[PRE + tb * FTILE - 1]) ? 1u : 0u;
    }
    Ma &= ra & ~((ra >> 1) | ((uint64_t)upa << 63));
    Mb &= rb & ~((rb >> 1) | ((uint64_t)upb << 63));
  }
  flat_store(c, ta, piece, Ma);
  flat_store(c, tb, piece, Mb);
}

// first position q >= from with q == 0 or byte(q-1) == delim, searched inside the window only:
// 128 bytes per step, one word per lane, zero-byte test on (word ^ delimiter splat)
__device__ int64_t find_line_start(const Ctx& c, int64_t from) {
  if (from <= 0) return 0;
  const int i0 = (int)(from - 1 - c.gw);
  const uint32_t splat = (uint32_t)c.a.delim * 0x01010101u;
  for (int base = i0 & ~3; base < c.wend; base += 128) {
    const int idx = base + 4 * c.lane;
    uint32_t z = 0;
    if (idx < c.wend) {  // the word may run into win[wend..]: the sentinel is filtered below
      const uint32_t t = *reinterpret_cast<const uint32_t*>(c.sm.win + idx) ^ splat;
      z = (t - 0x01010101u) & ~t & 0x80808080u;  // lowest set flag = first delimiter byte (exact)
      if (idx < i0) z &= 0xFFFFFFFFu << (8 * (i0 - idx));
    }
    const unsigned b = __ballot_sync(FULL, z != 0);
    if (b) {
      const int l = __ffs(b) - 1;
      const uint32_t zl = __shfl_sync(FULL, z, l);
      const int pos = base + 4 * l + ((__ffs(zl) - 1) >> 3);
      return pos < c.wend ? c.gw + pos + 1 : INF;
    }
  }
  return INF;
}

// ---- phase B: verify + chain ---------------------------------------------------------------------
// Positions inside phase B are window indices (int, relative to gw); a record must be < 2 GiB.
template <bool DIRECT>
struct Emitter {
  const Ctx& c;
  unsigned nkept = 0;        // warp-uniform
  bool overflow = false;     // warp-uniform (STAGE mode)
  unsigned long long goff;   // DIRECT mode: global index of this warp's first match
  __device__ Emitter(const Ctx& cc, unsigned long long g) : c(cc), goff(g) {}

  __device__ __forceinline__ void put(unsigned idx, int s, int e) {
    if (c.a.mode != M_FINDALL) return;
    if (DIRECT) {
      const unsigned long long gi = goff + idx;
      if ((int64_t)gi < c.a.cap) {
        const int64_t b = c.gw + c.a.base;
        *reinterpret_cast<longlong2*>(c.a.out + 2 * gi) = make_longlong2(b + s, b + e);
      }
    } else {
      if (idx < STG) c.sm.stage[c.buf][c.warp][idx] = make_uint2((unsigned)s, (unsigned)e);
    }
  }
};

// Multi-literal engine: one lane replays reference meta/findall.go:176-290 over
// Teddy.FindMatch (prefilter/teddy.go:391-445): the regime (scalar vs SIMD verify order) is fixed
// per call by the distance from the call's start position to the end of the haystack.
template <bool DIRECT>
__device__ int serial_chain_teddy(const Ctx& c, Emitter<DIRECT>& em, int pos_i, int last_cand, bool to_line_end) {
  unsigned added = 0;
  if (c.lane == 0) {
    const int64_t n = c.a.n;
    int64_t pos = c.gw + pos_i;
    const int64_t lastc = c.gw + last_cand;
    bool more = true;
    while (more && pos < n) {
      const bool scalar = n + c.a.after - pos < 16;
      more = false;
      for (int64_t p = pos; p + 2 <= n; p++) {
        const uint8_t b0 = byte_at(c, p);
        if (to_line_end && b0 == c.a.delim) break;
        if (!to_line_end && p > lastc) break;
        const uint32_t m = teddy_mask(c, b0, byte_at(c, p + 1));
        if (!m) continue;
        const int64_t e = scalar ? teddy_verify_scalar(c, p) : teddy_verify(c, p, m);
        if (e >= 0) {
          em.put(em.nkept + added, (int)(p - c.gw), (int)(e - c.gw));
          added++;
          pos = e;
          more = true;
          break;
        }
      }
    }
    pos_i = (int)(pos - c.gw);
  }
  added = __shfl_sync(FULL, added, 0);
  pos_i = __shfl_sync(FULL, pos_i, 0);
  if (!DIRECT && em.nkept + added > STG) em.overflow = true;
  em.nkept += added;
  return pos_i;
}

// One lane replays the reference loop from window index `pos` while the next candidate is
// <= last_cand (or, with to_line_end, until the current line ends).  Returns the new chain position.
template <bool DIRECT>
__device__ int serial_chain(const Ctx& c, Emitter<DIRECT>& em, int pos_i, int last_cand, bool to_line_end) {
  if (c.a.engine == SEL_TEDDY) return serial_chain_teddy<DIRECT>(c, em, pos_i, last_cand, to_line_end);
  unsigned added = 0;
  if (c.lane == 0) {
    const int64_t n = c.a.n;
    int64_t pos = c.gw + pos_i;
    const int64_t lastc = c.gw + last_cand;
    bool after_match = false;
    while (pos < n) {
      int64_t d = pos;  // reference: digitPos = prefilter.Find(haystack, pos)
      bool stop = false;
      while (d < n) {
        const uint8_t b = byte_at(c, d);
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
      // Later candidates belong to later batches — except a resume position in the middle of a
      // run, which the run-start filter never lists: it is tried here, right after its match.
      const bool mid_run = after_match && d == pos && c.a.filter.kind == F_RUNSTART && d > 0 &&
                           in_filter_set(c, byte_at(c, d - 1));
      if (!to_line_end && d > lastc && !mid_run) break;  // chain position unchanged
      after_match = false;
      const int64_t e = dfa_walk_slow(c, d);
      if (e >= 0) {
        em.put(em.nkept + added, (int)(d - c.gw), (int)(e - c.gw));
        added++;
        pos = e > d ? e : d + 1;
        after_match = true;
      } else {
        pos = d + 1;
        if (c.a.filter.kind == F_RUNSTART)
          while (pos < n && in_filter_set(c, byte_at(c, pos))) pos++;
      }
    }
    pos_i = (int)(pos - c.gw);
  }
  added = __shfl_sync(FULL, added, 0);
  pos_i = __shfl_sync(FULL, pos_i, 0);
  if (!DIRECT && em.nkept + added > STG) em.overflow = true;
  em.nkept += added;
  return pos_i;
}

// ---- record engine: one lane per record ----------------------------------------------------------
// The reference's UseDFA path (meta/find_indices.go:686-705, dfa/lazy/lazy.go:1102-1315 forward,
// :1769-1920 reverse): the unanchored forward DFA yields the end of the leftmost-first match, the
// reverse DFA (no break at match, last flag wins) its start; the search resumes at the end.
// Positions are window indices; bytes past the window come from global memory.
__device__ __forceinline__ unsigned line_byte(const Ctx& c, int i) {
  return i < WIN ? (unsigned)c.sm.win[i] : (unsigned)__ldg(c.a.h + c.gw + i);
}
__device__ __forceinline__ int kind_before(const Ctx& c, int i) {
  return (c.gw + i == 0) ? (c.a.base == 0 ? 2 : start_kind(c.a.delim)) : start_kind((uint8_t)line_byte(c, i - 1));
}
// start of the match that ends at e, not before lo
__device__ int line_reverse(const Ctx& c, int e, int lo, int nrel) {
  unsigned st = c.a.line.rstart[0];
  if (c.a.line.rkinds) st = c.a.line.rstart[e >= nrel ? (c.a.after == 0 ? 2 : start_kind(c.a.delim)) : start_kind((uint8_t)line_byte(c, e))];
  int last = lo, q = e;
  while (st) {
    if (q == lo) {
      // lower bound reached: only the pending flag counts (look-behind needs the byte before lo)
      if (c.gw + q == 0 && c.a.base == 0) {
        if (c.l_reoi[st]) last = q;
      } else {
        const unsigned b = c.gw + q == 0 ? (unsigned)c.a.delim : line_byte(c, q - 1);
        if (c.l_rt[(st << 8) + b] & 0x8000u) last = q;
      }
      break;
    }
    const unsigned t = c.l_rt[(st << 8) + line_byte(c, q - 1)];
    if (t & 0x8000u) last = q;
    st = t & 0x7FFFu;
    q--;
  }
  return last;
}
// all matches of the record that starts at window index i0.  EMIT=false: count them and remember
// the first two ends; EMIT=true: write (start,end) to consecutive output slots from `idx`.
template <bool EMIT, bool DIRECT>
__device__ void line_scan(const Ctx& c, Emitter<DIRECT>& em, int i0, unsigned idx, int& cnt, int& e0, int& e1) {
  const int64_t nr64 = c.a.n - c.gw;
  const int nrel = nr64 > 0x7fffffff ? 0x7fffffff : (int)nr64;
  int pos = i0;
  for (;;) {
    unsigned st = c.a.line.ustart[0];
    if (c.a.line.ukinds) st = c.a.line.ustart[kind_before(c, pos)];
    int last = -1, i = pos;
    while (st) {
      if (i >= nrel) {
        if (c.l_ueoi[st]) last = i;
        break;
      }
      const unsigned b = line_byte(c, i);
      const unsigned t = c.l_ut[(st << 8) + b];
      if (t & 0x8000u) last = i;
      st = t & 0x7FFFu;
      if (b == c.a.delim) break;
      i++;
    }
    if (last < 0) break;
    if (EMIT) {
      em.put(idx + cnt, line_reverse(c, last, pos, nrel), last);
    } else {
      if (cnt == 0) e0 = last;
      if (cnt == 1) e1 = last;
    }
    cnt++;
    if (c.a.mode == M_ISMATCH) break;
    pos = last;  // matches are non-empty, so this advances
  }
}

template <bool DIRECT>
__device__ void process_batch_lines(const Ctx& c, Emitter<DIRECT>& em, int cand, bool valid) {
  int cnt = 0, e0 = -1, e1 = -1;
  if (valid) line_scan<false, DIRECT>(c, em, cand, 0, cnt, e0, e1);
  int total;
  const int excl = warp_excl_scan(cnt, c.lane, total);
  if (!total) return;
  if (c.a.mode == M_ISMATCH) {
    if (c.lane == 0) c.a.total[1] = 1ull;
    em.nkept += total;
    return;
  }
  if (c.a.mode == M_FINDALL && cnt) {
    const int64_t nr64 = c.a.n - c.gw;
    const int nrel = nr64 > 0x7fffffff ? 0x7fffffff : (int)nr64;
    const unsigned idx = em.nkept + excl;
    if (cnt <= 2) {
      em.put(idx, line_reverse(c, e0, cand, nrel), e0);
      if (cnt == 2) em.put(idx + 1, line_reverse(c, e1, e0, nrel), e1);
    } else {
      int k = 0, x0, x1;
      line_scan<true, DIRECT>(c, em, cand, idx, k, x0, x1);
    }
  }
  if (!DIRECT && em.nkept + total > STG) em.overflow = true;
  em.nkept += total;
}

template <bool DIRECT>
__device__ __forceinline__ int process_batch(const Ctx& c, Emitter<DIRECT>& em, int cand, bool valid,
                                             int kept_end) {
  int end = -1;
  if (valid) end = c.a.engine == SEL_TEDDY ? teddy_walk(c, cand) : dfa_walk(c, cand);
  const bool ok = end >= 0;
  const unsigned okmask = __ballot_sync(FULL, ok);
  if (!okmask) return kept_end;
  if (c.a.mode == M_ISMATCH) {
    if (c.lane == 0) c.a.total[1] = 1ull;
    em.nkept += __popc(okmask);
    return kept_end;
  }
  const unsigned lower = okmask & ((1u << c.lane) - 1u);
  int pe = __shfl_sync(FULL, end, lower ? 31 - __clz(lower) : 0);
  if (!lower) pe = kept_end;
  bool bad = ok && cand < pe;
  if (c.a.skip_safe && ok && c.gw + end < c.a.n) bad |= in_filter_set(c, byte_at(c, c.gw + end));
  // literal candidates inside the last 16 bytes of the haystack may fall under the reference's
  // scalar verify order: let the exact replay decide
  bool replay = false;
  if (c.a.engine == SEL_TEDDY) replay = valid && c.gw + cand > c.a.n + c.a.after - 16;
  if (c.a.filter.kind == F_RUNSTART) replay |= bad;  // a resume position inside a run is no candidate
  replay = __any_sync(FULL, replay);
  if (!replay && __any_sync(FULL, bad)) {
    // Candidates overlap kept matches.  The filter is exhaustive (every possible match start is
    // a candidate), so the reference loop (meta/findall.go:176-290: search from the previous end,
    // take the first start that matches) is a walk over "first matching candidate at or after my
    // end" links.  Links by binary search over the (ascending) candidates, then follow the chain.
    const int cv = valid ? cand : 0x7fffffff;
    auto first_ok_at_or_after = [&](int t) {
      int lo = 0;
#pragma unroll
      for (int s = 16; s; s >>= 1) {
        const int probe = __shfl_sync(FULL, cv, lo + s - 1);
        if (probe < t) lo += s;
      }
      const int last = __shfl_sync(FULL, cv, lo);
      const unsigned m = last < t ? 0u : okmask & (~0u << lo);
      return m ? __ffs(m) - 1 : 32;
    };
    const int nxt = first_ok_at_or_after(ok ? end : 0x7fffffff);
    int cur = first_ok_at_or_after(kept_end);
    unsigned keptmask = 0;
    while (cur < 32) {
      keptmask |= 1u << cur;
      cur = __shfl_sync(FULL, nxt, cur);
    }
    if (!keptmask) return kept_end;
    if (keptmask >> c.lane & 1u) em.put(em.nkept + __popc(keptmask & ((1u << c.lane) - 1u)), cand, end);
    const unsigned add = __popc(keptmask);
    if (!DIRECT && em.nkept + add > STG) em.overflow = true;
    em.nkept += add;
    return __shfl_sync(FULL, end, 31 - __clz(keptmask));
  }
  if (replay) {
    // replay from the chain position through the last candidate of this batch
    const unsigned vmask = __ballot_sync(FULL, valid);
    const int last_cand = __shfl_sync(FULL, cand, 31 - __clz(vmask));
    return serial_chain<DIRECT>(c, em, kept_end, last_cand, false);
  }
  if (ok) em.put(em.nkept + __popc(lower), cand, end);
  const unsigned add = __popc(okmask);
  if (!DIRECT && em.nkept + add > STG) em.overflow = true;
  em.nkept += add;
  return __shfl_sync(FULL, end, 31 - __clz(okmask));
}

template <bool DIRECT>
__device__ void phase_b(const Ctx& c, Emitter<DIRECT>& em) {
  const int64_t lo = c.sm.ls[c.warp];
  const int64_t hi = c.sm.ls[c.warp + 1];
  if (lo >= hi || lo >= c.a.n) return;
  const int64_t bm_end = c.cbeg + CH + OVER;  // bitmap covers [cbeg, bm_end)
  const int64_t hi_b = hi < bm_end ? hi : bm_end;
  int kept_end = (int)(lo - c.gw);
  uint16_t* q = c.sm.queue[c.warp];
  int qlen = 0;
  const bool lines = c.a.engine == SEL_LINE;
  const int lo_rel = (int)(lo - c.cbeg);
  int hi_rel = (int)(hi_b - c.cbeg);
  if (lines) {
    // the bitmap marks record delimiters: the owned records are the one starting at `lo` plus
    // one after every delimiter in [lo, hi-1) (the delimiter at hi-1 opens the next owner's record)
    if (hi <= bm_end) hi_rel -= 1;
    if (c.lane == 0) q[0] = (uint16_t)(lo - c.gw);
    qlen = 1;
    __syncwarp();
  }
  // Four bitmap words per lane and step (4 KB of input): count, warp-scan, then every lane appends
  // its own candidates to the queue in position order.  A step that holds more candidates than the
  // queue can take (dense data) is split into passes of four lanes each.
  const int w_lo = lo_rel >> 5, w_hi = (hi_rel + 31) >> 5;  // bitmap words [w_lo, w_hi)
  auto drain = [&]() {
    __syncwarp();
    int head = 0;
    while (qlen - head >= 32) {
      if (lines) process_batch_lines<DIRECT>(c, em, q[head + c.lane], true);
      else kept_end = process_batch<DIRECT>(c, em, q[head + c.lane], true, kept_end);
      head += 32;
    }
    if (head) {
      const int rem = qlen - head;
      const uint16_t tmp = c.lane < rem ? q[head + c.lane] : 0;
      __syncwarp();
      if (c.lane < rem) q[c.lane] = tmp;
      qlen = rem;
      __syncwarp();
    }
  };
  for (int wbase = w_lo & ~3; wbase < w_hi; wbase += 128) {
    const int w0 = wbase + 4 * c.lane;
    uint4 ld = make_uint4(0u, 0u, 0u, 0u);
    if (w0 < w_hi && w0 < NWORDS) ld = *reinterpret_cast<const uint4*>(&c.sm.cand[w0]);
    uint32_t wd[4] = {ld.x, ld.y, ld.z, ld.w};
    if (w0 * 32 < lo_rel || (w0 + 4) * 32 > hi_rel) {  // a lane at either end of the owned range
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int lo_sh = lo_rel - (w0 + k) * 32, hi_sh = hi_rel - (w0 + k) * 32;
        if (lo_sh > 0) wd[k] = lo_sh >= 32 ? 0u : wd[k] & (~0u << lo_sh);
        if (hi_sh < 32) wd[k] = hi_sh <= 0 ? 0u : wd[k] & ((1u << hi_sh) - 1u);
      }
    }
    const int cnt = __popc(wd[0]) + __popc(wd[1]) + __popc(wd[2]) + __popc(wd[3]);
    int total;
    const int excl = warp_excl_scan(cnt, c.lane, total);
    if (!total) continue;
    auto append = [&](int off) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t v = wd[k];
        const int base = (w0 + k) * 32 + PRE + (lines ? 1 : 0);  // window index of bit 0 (record: byte after)
        while (v) {
          q[off++] = (uint16_t)(base + __ffs(v) - 1);
          v &= v - 1;
        }
      }
    };
    if (qlen + total <= QCAP) {
      append(qlen + excl);
      qlen += total;
      drain();
    } else {
      // dense: four lanes (512 bytes, at most 512 new entries on top of a remainder < 32) per pass
      for (int pass = 0; pass < 8; pass++) {
        const bool mine = (c.lane >> 2) == pass;
        int tot2;
        const int off = qlen + warp_excl_scan(mine ? cnt : 0, c.lane, tot2);
        if (!tot2) continue;
        if (mine) append(off);
        qlen += tot2;
        drain();
      }
    }
  }
  if (qlen) {
    const bool valid = c.lane < qlen;
    if (lines) process_batch_lines<DIRECT>(c, em, valid ? q[c.lane] : 0, valid);
    else kept_end = process_batch<DIRECT>(c, em, valid ? q[c.lane] : 0, valid, kept_end);
  }
  if (hi > bm_end && !lines) {
    // the last owned line runs past the classified window: finish it serially
    const int bme = (int)(bm_end - c.gw);
    serial_chain<DIRECT>(c, em, kept_end > bme ? kept_end : bme, 0, true);
  }
}

// Decoupled look-back for `chunk` with local aggregate `agg` (warp 0 only): publishes the
// inclusive prefix and returns the exclusive one.
__device__ unsigned long long look_back(const ScanArgs& a, int64_t chunk, unsigned agg, int lane) {
  unsigned long long excl = 0;
  if (chunk == 0) {
    if (lane == 0) st_status(&a.status[0], LB_PREFIX | agg);
    return 0;
  }
  int64_t look = chunk - 1;
  for (;;) {
    const int64_t idx = look - lane;
    unsigned long long v = LB_PREFIX;  // lanes before chunk 0 act as a zero prefix
    if (idx >= 0) {
      do {
        v = ld_status(&a.status[idx]);
      } while ((v >> 62) == 0);
    }
    const unsigned pm = __ballot_sync(FULL, (v >> 62) == 2);
    const int first = pm ? __ffs(pm) - 1 : 32;  // nearest chunk that already has a prefix
    unsigned long long val = lane <= first ? (v & LB_VALUE) : 0ull;
#pragma unroll
    for (int d = 16; d; d >>= 1) val += __shfl_xor_sync(FULL, val, d);
    excl += val;
    if (pm) break;
    look -= 32;
  }
  if (lane == 0) st_status(&a.status[chunk], LB_PREFIX | (excl + agg));
  return excl;
}

// all staged matches of one chunk, written by the writer warp in match order
__device__ __forceinline__ void write_staged(const ScanArgs& a, Smem& sm, int buf, int lane, int64_t gw,
                                             unsigned long long gbase) {
  const int64_t b = gw + a.base;
  for (int w = 0; w < WARPS; w++) {
    const unsigned cnt = sm.wcount2[buf][w];
    for (unsigned i = lane; i < cnt; i += 32) {
      const unsigned long long gi = gbase + i;
      if ((int64_t)gi < a.cap) {
        const uint2 m = sm.stage[buf][w][i];
        *reinterpret_cast<longlong2*>(a.out + 2 * gi) = make_longlong2(b + m.x, b + m.y);
      }
    }
    gbase += cnt;
  }
}

#ifdef CGX_TIMING
#define TSTAMP(var) const long long var = clock64()
#define TACC(slot, t0, t1) \
  if (lane == 0) atomicAdd(&a.total[slot], (unsigned long long)((t1) - (t0)))
#else
#define TSTAMP(var)
#define TACC(slot, t0, t1)
#endif

// Roles: warps 0..7 scan (ticket -> TMA load -> phase A -> phase B into a staging buffer),
// warp 8 turns staged chunks into ordered global output (decoupled look-back + int64 stores).
// Two staging buffers decouple them: a scanning warp only waits if the writer is two chunks behind.
__global__ void __launch_bounds__(THREADS, 3) scan_dfa_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  uint32_t* s_trans = reinterpret_cast<uint32_t*>(smem_raw + sizeof(Smem));
  uint8_t* s_eoi = reinterpret_cast<uint8_t*>(s_trans + (size_t)a.dfa.nstates * 256);
  uint8_t* s_lut = s_eoi + ((a.dfa.nstates + 15) & ~15);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // stage the automaton once per CTA, widening u16 (state | match<<15) to premultiplied u32
  for (int i = tid; i < a.dfa.nstates * 256; i += THREADS) {
    const uint32_t e = a.dfa.trans[i];
    s_trans[i] = ((e & 0x7FFFu) << 10) | ((e & 0x8000u) << 16);
  }
  for (int i = tid; i < a.dfa.nstates; i += THREADS) s_eoi[i] = a.dfa.eoi[i];
  // ... or the literal tables (one blob, same layout as in global memory)
  const unsigned char* t_blob = smem_raw + sizeof(Smem);
  if (a.engine == SEL_TEDDY)
    for (int i = tid; i < a.teddy.blob_bytes / 4; i += THREADS)
      reinterpret_cast<uint32_t*>(smem_raw + sizeof(Smem))[i] = a.teddy.fp[i];
  if (a.engine == SEL_LINE)
    for (int i = tid; i < a.line.blob_bytes / 4; i += THREADS)
      reinterpret_cast<uint32_t*>(smem_raw + sizeof(Smem))[i] = reinterpret_cast<const uint32_t*>(a.line.blob)[i];
  const uint16_t* l_ut = reinterpret_cast<const uint16_t*>(smem_raw + sizeof(Smem));
  const uint16_t* l_rt = l_ut + (size_t)a.line.un * 256;
  const uint8_t* l_ueoi = reinterpret_cast<const uint8_t*>(l_rt + (size_t)a.line.rn * 256);
  const unsigned char* g_blob = reinterpret_cast<const unsigned char*>(a.teddy.fp);
  if (a.filter.kind == F_LUT)
    for (int i = tid; i < 256; i += THREADS) s_lut[i] = a.filter.lut[i];
  if (tid < WINPAD) sm.win[WIN + tid] = a.delim;
  for (int i = tid; i < (int)(sizeof(FlatDev) / 4); i += THREADS)
    reinterpret_cast<uint32_t*>(&sm.flat)[i] = reinterpret_cast<const uint32_t*>(&a.flat)[i];
  if (tid == 0) {
    mbar_init(&sm.mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == WARPS) {
    // ---------------- writer warp ----------------
    if (a.mode != M_FINDALL) return;
    for (int b = 0;; b ^= 1) {
      bar_sync(BAR_FULL + b, THREADS);
      const int64_t chunk = sm.meta_chunk[b];
      if (chunk < 0) break;
      const unsigned agg = sm.meta_agg[b];
      const unsigned long long excl = look_back(a, chunk, agg, lane);
      if (lane == 0 && chunk == a.nchunks - 1) a.total[0] = excl + agg;
      if (sm.meta_ovf[b]) {
        if (lane == 0) sm.cta_base = excl;
        __syncwarp();
        bar_arrive(BAR_OVF, THREADS);  // the scanning warps replay phase B writing directly
      } else {
        write_staged(a, sm, b, lane, sm.meta_gw[b], excl);
      }
      bar_arrive(BAR_EMPTY + b, THREADS);
    }
    return;
  }

  // ---------------- scanning warps ----------------
  uint32_t parity = 0;
  int buf = 0;
  unsigned used = 0;  // FINDALL chunks staged so far
  unsigned it = 0;
  // one thread decides for the whole CTA, so the loop exit is uniform
  auto next_ticket = [&]() -> unsigned {
    if (a.mode == M_ISMATCH && *((volatile unsigned long long*)&a.total[1])) return 0xFFFFFFFFu;
    return atomicAdd(a.ticket, 1u);
  };
  if (tid == 0) sm.chunk2[0] = next_ticket();
  bar_sync(BAR_COMPUTE, CTHREADS);
  for (;; it++) {
    // the ticket of this iteration was fetched before the previous iteration's last barrier
    const int64_t chunk = sm.chunk2[it & 1];
    if (chunk >= a.nchunks) break;  // also how an is-match early exit arrives (see next_ticket)
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
      // the next ticket is drawn while this chunk's window is in flight (its latency hides behind
      // the wait below), and the next window starts moving towards L2 a whole chunk ahead
      const unsigned nxt = next_ticket();
      sm.chunk2[(it + 1) & 1] = nxt;
      if ((int64_t)nxt < a.nchunks) {
        const int64_t nlo = (int64_t)nxt * CH - PRE;  // nxt > 0 here: chunk 0 is never a "next"
        const int64_t nhi = nlo + WIN < a.n ? nlo + WIN : a.n;
        const uint32_t nb = (uint32_t)(nhi - nlo) & ~15u;
        if (nb) tma_prefetch_l2(a.h + nlo, nb);
      }
    }
    // bytes the bulk copy does not cover: before position 0, the <16 B tail, past the end
    if (gw < 0 || bulk != (uint32_t)WIN) {
      for (int i = tid; i < WIN; i += CTHREADS) {
        const int64_t g = gw + i;
        if (g < 0 || g >= lo_g + bulk) sm.win[i] = g >= 0 && g < a.n ? a.h[g] : a.delim;
      }
    }
    TSTAMP(t_0);
    mbar_wait(&sm.mbar, parity);  // every scanning thread observes the TMA completion itself
    parity ^= 1;
    TSTAMP(t_1);
    TACC(2, t_0, t_1);
    if (gw < 0 || bulk != (uint32_t)WIN) bar_sync(BAR_COMPUTE, CTHREADS);  // generic fill stores

    Ctx c{a, sm, s_trans, s_eoi, s_lut, cbeg, gw, (int)(hi_g - gw), lane, warp, buf,
          reinterpret_cast<const uint32_t*>(t_blob),
          reinterpret_cast<const uint64_t*>(t_blob + ((const unsigned char*)a.teddy.lit8 - g_blob)),
          reinterpret_cast<const uint16_t*>(t_blob + ((const unsigned char*)a.teddy.fp2 - g_blob)),
          t_blob + (a.teddy.bytes - g_blob),
          reinterpret_cast<const int32_t*>(t_blob + ((const unsigned char*)a.teddy.offs - g_blob)),
          reinterpret_cast<const uint16_t*>(t_blob + ((const unsigned char*)a.teddy.order - g_blob)),
          reinterpret_cast<const uint16_t*>(t_blob + ((const unsigned char*)a.teddy.bucket_off - g_blob)),
          l_ut, l_rt, l_ueoi, l_ueoi + a.line.un};
    if (a.engine == SEL_TEDDY) phase_a_teddy(c);
    else if (a.flat.nops) phase_a_flat(c);
    else phase_a_plain(c);
    TSTAMP(t_1a);
    TACC(3, t_1, t_1a);
    {
      const int64_t s = find_line_start(c, cbeg + (int64_t)warp * SUB);
      if (lane == 0) sm.ls[warp] = s;
      if (warp == WARPS - 1) {
        const int64_t e = find_line_start(c, cbeg + CH);
        if (lane == 0) sm.ls[WARPS] = e;
      }
      if (tid == 0) {
        sm.woverflow2[it & 1] = 0;
      }
    }
    TSTAMP(t_2);
    TACC(5, t_1a, t_2);
    if (a.mode == M_FINDALL && used >= 2) bar_sync(BAR_EMPTY + buf, THREADS);  // buffer released?
    TSTAMP(t_2b);
    TACC(6, t_2, t_2b);
    bar_sync(BAR_COMPUTE, CTHREADS);
    TSTAMP(t_3);
    TACC(7, t_2b, t_3);

    Emitter<false> em(c, 0);
    phase_b<false>(c, em);
    if (lane == 0) {
      sm.wcount2[buf][warp] = em.nkept;
      if (em.overflow) atomicOr(&sm.woverflow2[it & 1], 1u);
    }
    bar_sync(BAR_COMPUTE, CTHREADS);

    unsigned agg = 0;
    for (int w = 0; w < WARPS; w++) agg += sm.wcount2[buf][w];
    if (a.mode == M_FINDALL) {
      const bool ovf = sm.woverflow2[it & 1] != 0;
      if (tid == 0) {
        // the aggregate is visible to other CTAs immediately; the prefix follows from the writer
        if (chunk != 0) st_status(&a.status[chunk], LB_AGG | agg);
        sm.meta_chunk[buf] = chunk;
        sm.meta_gw[buf] = gw;
        sm.meta_agg[buf] = agg;
        sm.meta_ovf[buf] = ovf ? 1u : 0u;
      }
      __syncwarp();
      bar_arrive(BAR_FULL + buf, THREADS);
      if (ovf) {
        // more matches than the staging buffer holds: wait for the offset, then replay phase B
        // writing straight to global memory (the window is still resident)
        bar_sync(BAR_OVF, THREADS);
        unsigned wexcl = 0;
        for (int w = 0; w < warp; w++) wexcl += sm.wcount2[buf][w];
        Emitter<true> em2(c, sm.cta_base + wexcl);
        phase_b<true>(c, em2);
        bar_sync(BAR_COMPUTE, CTHREADS);  // the replay still reads the window
      }
      used++;
      buf ^= 1;
    } else {
      if (tid == 0 && agg) {
        atomicAdd(a.total, (unsigned long long)agg);
        a.total[1] = 1ull;  // is-match flag (also reached through the serial replay paths)
      }
    }
    // no barrier here: everything the next iteration overwrites (window, bitmap, class scratch)
    // was last read before the barrier that followed phase B
  }
  if (a.mode == M_FINDALL) {
    // tell the writer to finish (after it released the buffer we are about to mark)
    if (used >= 2) bar_sync(BAR_EMPTY + buf, THREADS);
    if (tid == 0) sm.meta_chunk[buf] = -1;
    __syncwarp();
    bar_arrive(BAR_FULL + buf, THREADS);
  }
}

}  // namespace

size_t scan_dfa_smem_bytes(int nstates, int blob_bytes) {
  if (blob_bytes) return sizeof(Smem) + (size_t)blob_bytes + 16;
  return sizeof(Smem) + (size_t)nstates * 1024 + ((nstates + 15) & ~15) + 256;
}

int64_t scan_dfa_chunks(int64_t n) { return n <= 0 ? 0 : (n + CH - 1) / CH; }

// Launches the scan on `stream`.  ticket/status/total must be zeroed by the caller.
cudaError_t launch_scan_dfa(const ScanArgs& a, int sm_count, cudaStream_t stream) {
  if (a.nchunks == 0) return cudaSuccess;
  const size_t smem = scan_dfa_smem_bytes(a.dfa.nstates, a.engine == SEL_TEDDY ? a.teddy.blob_bytes
                                                               : a.engine == SEL_LINE ? a.line.blob_bytes : 0);
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
