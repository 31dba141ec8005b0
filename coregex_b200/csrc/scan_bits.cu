// scan_bits.cu — the bitstream engine, second generation: sm_100a kernel for flat deterministic
// patterns (`\d+\.\d+\.\d+\.\d+`, `\w+@\w+\.\w+`, `[a-z]+=\d+`, `\d+`, `\w+` ...), the north-star path.
//
// Replaces, for a whole corpus at once:
//   reference meta/findall.go:176-290        findAllIndicesLoop (pos = end chaining)
//   reference meta/find_indices.go:1050-1088 DigitPrefilter loop (candidate -> SearchAtAnchored)
//   reference simd/memchr_digit_amd64.s:26   memchrDigitAVX2
//   reference dfa/lazy/lazy.go:219-324       SearchAtAnchored (per-byte class + table walk)
//   reference nfa/charclass_searcher.go      FindAllIndices (lone `C+` patterns)
//
// Match STARTS come from a right-to-left marker pass over per-class position bitmaps, match ENDS
// from a left-to-right pass over the same bitmaps (forced greedy == leftmost-first because the
// host proved the pattern deterministic, host/engine.cpp DecideBitstream).  The first generation
// (round 1) ran both passes across the warp — lane l held one 64-bit word of a 2 KB tile and every
// step of every pass paid a shuffle for the carry, a ballot for ownership and a rank scan per tile;
// it was bound by instruction issue at 0.31 warp-instructions per byte.  Here the passes are
// LANE-SERIAL:
//
//   phase A  a warp draws a chunk of TPC tiles (ticket counter); per tile one TMA bulk copy of 2 KB
//            into the warp's window ring, lane l classifies the 64-byte piece l (4 x LDS.128, SWAR
//            range tests, dp4a bit packing) and stores one 64-bit word per class into the chunk's
//            class-bitmap array in shared memory — the transposition from "lane = piece of a tile"
//            to "lane = contiguous run of K pieces" costs one STS and one LDS per word;
//   phase B  lane l owns the K consecutive words [K l, K l + K) of the chunk plus the first word of
//            its neighbour (overlap).  Sweep 1 walks them right to left, sweep 2 left to right; a
//            carry or a shifted-out bit travels from one word to the next in a register.  No
//            shuffle, no ballot, no scan inside the passes;
//   ownership a byte of no class ("sync byte") can be in no match, so matches never cross one.  A
//            lane owns the starts after the first sync byte at or after its first word up to the
//            first sync byte at or after its neighbour's first word: every start has exactly one
//            owner, and whatever enters a lane's window from outside dies at a sync byte before it
//            reaches an owned start;
//   output   the passes leave two bitmaps (starts, ends) per chunk in shared memory — any number of
//            matches fits, there is no staging overflow.  The chunk's count is published at once;
//            one resolver warp per CTA performs the two-level decoupled look-back and hands the
//            chunk's global offset back; the scanning warp then turns its bitmaps into int64
//            (start,end) pairs in global match order (rank = lane prefix + position in the lane).
//   exact    starts and ends must alternate; a lane whose words fail the check (overlapping
//            candidates, `1.2.3.4.5`) or whose last segment has no sync byte inside the window
//            replays the reference loop for exactly its own range (serial, rare).
// Every corpus byte crosses HBM once; output is 16 B per match.  A FindAll launch needs no memset:
// look-back words carry the launch's epoch, group accumulators and the ticket clean themselves.
#include "scan_common.cuh"
#include "scan_params.h"

// CGX_TEDDY (csrc/scan_teddy.cu includes this file with it): the same skeleton — gangs of chunks,
// per-warp TMA window ring, staged matches, one look-back per gang — with the multi-literal engine
// in place of the class tests and sweeps.  Replaces reference prefilter/teddy.go:391-445 FindMatch
// (teddy_ssse3_amd64.s:273, teddy_avx2_amd64.s:43) + :532-550 verifyBucket under the FindAll loop.
//   phase A  per byte two table loads (bucket masks of fingerprint bytes 0 and 1) and one AND with
//            predicate: a candidate bitmap per chunk;
//   phase B  every lane runs the reference loop (candidate -> verify -> continue at the match end)
//            over its own K words.  It enters the chain at a SAFE POINT: it first verifies the
//            candidates of the word before its region without chaining; a position of that word's
//            second half that lies strictly inside none of those spans cannot lie inside a kept match
//            either (literals are at most 32 bytes long), so the chain state there is known.  No
//            lane depends on another one, chunks need one word of context before their first
//            owned word and one after their last.
#ifdef CGX_TEDDY
#define scan_flat_chunks scan_teddy_chunks
#define scan_flat_smem_bytes scan_teddy_smem_bytes
#define scan_flat_threads scan_teddy_threads
#define scan_flat_warps scan_teddy_warps
#define launch_scan_flat launch_scan_teddy
#define sim_launch_scan_flat sim_launch_scan_teddy
#endif

namespace cgx {

namespace {

constexpr uint32_t FULL = 0xffffffffu;
#ifndef CGX_K
#define CGX_K 4            // words (64-byte pieces) per lane and chunk == tiles per chunk
#endif
#ifndef CGX_CTAS
#define CGX_CTAS 1         // resident CTAs per SM the kernel is built for
#endif
#ifndef CGX_NB
#define CGX_NB 2           // window ring: tiles in flight per warp
#endif
#ifndef CGX_UNROLL
#define CGX_UNROLL (CGX_K + 1)  // sweep loops fully unrolled: the per-step carries of a sweep stay in predicate registers
#endif
constexpr int K = CGX_K;
constexpr int FW_CTAS = CGX_CTAS;
constexpr int NB = CGX_NB;
constexpr int UNROLL = CGX_UNROLL;
#ifndef CGX_UNROLL_A
#define CGX_UNROLL_A 1     // trips around the window ring unrolled in the tile loop
#endif
constexpr int UNROLL_A = CGX_UNROLL_A;
#ifndef CGX_FAST_UNROLL
#define CGX_FAST_UNROLL 1  // the same in the fast form of the tile loop (fully unrolled it ran 9 % slower: instruction cache)
#endif
constexpr int FAST_UNROLL = CGX_FAST_UNROLL;
constexpr int TILE = 2048;                 // one bulk copy, one classification round (64 B per lane)
constexpr int TPC = K;                     // tiles per chunk
constexpr int NWORDS = 32 * K;             // words of a chunk's window
constexpr int WINDOW = NWORDS * 64;        // bytes of a chunk's window
#ifdef CGX_TEDDY
constexpr int OVLW = 2;                    // a word of context before the owned words, one after
#else
constexpr int OVLW = 1;
#endif
constexpr int CHUNKB = (NWORDS - OVLW) * 64;  // bytes between chunk origins: chunks overlap by OVLW words
constexpr int NSLOTS = 32 * (K + 1);       // word w lives in slot w + w / K (one pad slot per lane region)
// slot NSLOTS stays all zero: lane 31 has no neighbour word, it reads this one instead (no class
// byte, no marker: a word that leaves every piece of per-lane state as it is)
constexpr int NSLOTS1 = NSLOTS + 1;
// staged matches per chunk (the IP corpus has ~80 per 8 KB).  CGX_PARK (dense patterns, see below):
// every chunk parks its bitmaps; ONE staging buffer turns them into coalesced stores, CAP pairs a round
#ifndef CGX_CAP
#define CGX_CAP_DEFAULT 1
#endif
static_assert(K == 4 || K == 8 || K == 16, "words per lane: a power of two that divides 32");

// ---- pipe-aware primitives (the kernel is bound by the integer ALU pipe) -------
//  * a LOP3 takes one immediate at most; as an explicit lop3.b32 `(w ^ k) & m` is one instruction;
//  * `x + k` as x * one + k with a multiplier the compiler cannot see through is an IMAD: same
//    result, issued to the FMA pipe, which is otherwise idle.
#if defined(CGX_CPU_SIM) || !defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t xor_and(uint32_t w, uint32_t k, uint32_t m) { return (w ^ k) & m; }
__device__ __forceinline__ uint32_t nor_and(uint32_t z, uint32_t w, uint32_t m) { return ~(z | w) & m; }
__device__ __forceinline__ uint32_t mad_fma(uint32_t x, uint32_t one, uint32_t k) { return x * one + k; }
#else
__device__ __forceinline__ uint32_t xor_and(uint32_t w, uint32_t k, uint32_t m) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(d) : "r"(w), "r"(k), "r"(m));  // (a ^ b) & c
  return d;
}
__device__ __forceinline__ uint32_t nor_and(uint32_t z, uint32_t w, uint32_t m) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(d) : "r"(z), "r"(w), "r"(m));  // ~(a | b) & c
  return d;
}
__device__ __forceinline__ uint32_t mad_fma(uint32_t x, uint32_t one, uint32_t k) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(k));
  return d;
}
#endif

#ifdef CGX_JIT
#include "cgx_jit_prog.h"  // generated per pattern (host/engine.cpp JitHeader), compiled by csrc/jit.cu
#endif

// The pattern-dependent parts exist twice: as an interpreter over ScanArgs::flat (this translation
// unit as nvcc builds it, and the CPU emulator build), and — when the host JIT-compiles this file
// with NVRTC for one pattern (csrc/jit.cu, -DCGX_JIT + a generated cgx_jit_prog.h) — as
// straight-line code: no program loads, no dispatch, class constants as immediates, one search mode.
#ifdef CGX_JIT
#define P_NCLASSES CGX_JIT_NCLASSES
#define P_RUNSTART CGX_JIT_RUNSTART
#define P_MIDRUN CGX_JIT_MIDRUN
#define P_MODE CGX_JIT_MODE
#else
#define P_NCLASSES f.nclasses
#define P_RUNSTART f.bs_runstart
#define P_MIDRUN f.bs_midrun_check
#define P_MODE a.mode
#endif
#ifdef CGX_JIT
constexpr int NC = CGX_JIT_NCLASSES;       // class words per slot
constexpr int NSTATE = CGX_JIT_NSTATE;     // per-step carry / shift state of a sweep
#elif defined(CGX_TEDDY)
constexpr int NC = 2;                      // slot = (candidates -> starts, ends)
constexpr int NSTATE = 1;
#else
constexpr int NC = 4;
constexpr int NSTATE = 24;
#endif
constexpr int NPAIR = (NC + 1) / 2;        // 16-byte slot arrays: classes (0,1) and (2,3)
// CGX_PARK: patterns whose matches can be a byte or two long (`\d+`, `\w+`: a chunk holds far more
// matches than the staging buffer) get a second set of slot arrays: a chunk that does not fit the
// staging buffer leaves its (starts, ends) bitmaps where they are and the next chunk is scanned in
// the other set, instead of waiting for the gang's offset on the spot.
#if defined(CGX_JIT) && defined(CGX_JIT_PARK) && !defined(CGX_PARK)
#define CGX_PARK CGX_JIT_PARK
#endif
#ifndef CGX_PARK
#define CGX_PARK 0
#endif
constexpr int NBUF = CGX_PARK ? 2 : 1;
#ifdef CGX_CAP_DEFAULT
#define CGX_CAP (CGX_PARK ? 1024 : 160)
#endif
constexpr int CAP = CGX_CAP;
constexpr int NST = CGX_PARK ? 1 : 2;      // staging buffers
// scanning warps per CTA: as many as the shared memory of one SM holds (one or two slot arrays per warp)
// The register file is split over the four schedulers: with the resolver the CTA has 24 warps (6 per
// scheduler, 80 registers a thread) when one slot array per warp fits the shared memory of an SM 23
// times, 20 warps (96 registers) with two slot arrays.  Measured on B200 (IP regex, 16 GiB):
// K=8 x 15 warps 3225 GB/s, K=4 x 19 warps 3203 GB/s, K=4 x 23 warps 3380 GB/s.
#ifndef CGX_WARPS
#define CGX_WARPS (NPAIR == 1 ? (CGX_PARK ? 15 : 23) : (CGX_PARK ? 11 : 19))
#endif
constexpr int FW_WARPS = CGX_WARPS;
constexpr int FW_THREADS = (FW_WARPS + 1) * 32;  // + one resolver warp
static_assert(FW_WARPS <= 31, "one CTA holds at most 31 scanning warps and the resolver");

// global position (relative to h) of a chunk's window: the multi-literal engine keeps one word of
// context before the owned words (chunk 0 starts the haystack and has none)
__device__ __forceinline__ int64_t chunk_origin(int64_t chunk) {
#ifdef CGX_TEDDY
  return chunk * (int64_t)CHUNKB - (chunk > 0 ? 64 : 0);
#else
  return chunk * (int64_t)CHUNKB;
#endif
}

template <bool B>
struct BoolC {
  static constexpr bool value = B;
};

struct alignas(16) Slot {
  uint64_t a, b;
};
struct WarpSmem {
  alignas(128) uint8_t win[NB][TILE];
  // class bitmaps of the chunk being scanned; after sweep 2 array 0 holds (starts, ends) instead
  Slot cls[NBUF][NPAIR][NSLOTS1];
  uint64_t mk[NSLOTS1];       // sweep 1 -> sweep 2: "a match can start here", forward orientation
  // A chunk's matches wait here (chunk-relative u16 offsets, in match order) for the chunk's global
  // offset while the next chunk is scanned: two buffers.  A chunk with more than CAP matches keeps
  // its bitmaps instead and waits for its offset on the spot.
  alignas(8) uint16_t st[NST][2][CAP];  // [buffer][starts | ends]; also the scratch of replay_bits
  uint64_t mbar[NB];
  // matches of a serially replayed segment that end beyond the chunk's bitmap (per staging buffer):
  // found again, and stored, when the chunk's offset is known
  int64_t far_from[2], far_stop[2];
  uint32_t far_cnt[2];
  uint32_t bits_cnt[2];       // matches recorded in the bitmaps / staged
};
// a scanned chunk handed from a scanning warp to the CTA's resolver warp (one slot per buffer)
struct Mail {
  volatile int state;  // 0 = buffer free, 1 = chunk waits for its offset, 2 = offset known
  unsigned cnt;
  int64_t chunk;
  unsigned long long excl;  // state 2: global index of the chunk's first match
};
struct CtaSmem {
  WarpSmem w[FW_WARPS];
  Mail mail[FW_WARPS][2];
  // FindAll: the CTA's warps scan GANGS of FW_WARPS consecutive chunks (warp w takes chunk
  // gang * FW_WARPS + w).  The resolver warp draws the gang tickets, two gangs ahead, into this ring:
  // (sequence number + 1) << 32 | ticket.
  volatile unsigned long long tk[4];
  // arrivals << 40 | matches of the gang being scanned (per parity of its sequence number): the
  // last warp to arrive publishes the gang's count in the look-back words
  unsigned long long gang_acc[2];
#ifdef CGX_TEDDY
  // bucket masks per byte value of fingerprint byte 0 / byte 1 (reference prefilter/teddy.go:271-311
  // buildMasks, the two nibble tables of a position folded into one byte table)
  uint16_t tfa[256], tfb[256];
  // the same for the third byte (every literal has three bytes or more, teddy.go:NewTeddy): cuts the
  // candidates that reach verification by an order of magnitude
  uint16_t tfc[256];
  // the literals as the verify wants them (at most 64: Fat Teddy's limit): first 8 bytes, byte mask,
  // length, bucket-major order — a verify is one round trip to the haystack and shared-memory reads
  unsigned long long tlit8[64], tmsk8[64];
  uint16_t tlen[64], torder[64], tboff[17];
#endif
};

__device__ __forceinline__ uint64_t mk64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ uint32_t lo32(uint64_t x) { return (uint32_t)x; }
__device__ __forceinline__ uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }

// 64-bit x + y + cin (cin is 0 or 1) with the carry out as a 0/1 register: the carry enters and
// leaves through the adder's flag (four IADD3) — no 64-bit compares.
__device__ __forceinline__ uint64_t adc64(uint64_t x, uint64_t y, uint32_t& carry) {
#if defined(CGX_CPU_SIM) || !defined(__CUDA_ARCH__)
  const uint64_t s1 = x + y;
  const uint64_t s2 = s1 + carry;
  carry = (s1 < x || s2 < s1) ? 1u : 0u;
  return s2;
#else
  uint32_t lo, hi, c = carry;
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %3, 0xffffffff;\n\taddc.cc.u32 %0, %4, %6;\n\taddc.cc.u32 %1, %5, %7;\n\t"
      "addc.u32 %2, 0, 0;\n\t}"
      : "=r"(lo), "=r"(hi), "=r"(carry)
      : "r"(c), "r"(lo32(x)), "r"(hi32(x)), "r"(lo32(y)), "r"(hi32(y)));
  return mk64(hi, lo);
#endif
}
// markers move s positions (1 or 2) towards higher bit indices; `below` is the high half of the word
// processed before this one (its top bits enter at the bottom)
template <int S>
__device__ __forceinline__ uint64_t shl_in(uint64_t m, uint32_t below) {
  return mk64(__funnelshift_l(lo32(m), hi32(m), S), __funnelshift_l(below, lo32(m), S));
}
__device__ __forceinline__ uint64_t brev64(uint64_t x) { return mk64(__brev(lo32(x)), __brev(hi32(x))); }

// ---- classification ----------------------------------------------------------------------------
// 8 flag words (bit 7 of a byte set <=> byte in class) -> bit-reversed 32-bit mask (bit 31-b <=> byte b).
// Each dp4a pair gathers 8 flags into bits 7..14; the groups are chained through the accumulator
// input (shifted by 8 each time: IMAD.SHL, FMA pipe) and the last one is joined by a multiply-add,
// so the whole pack costs one ALU-pipe instruction (the final right shift).
__device__ __forceinline__ uint32_t pack_rev(const uint32_t* fl, uint32_t one) {
  uint32_t acc = __dp4a(fl[0], 0x10204080u, 0u);
  acc = __dp4a(fl[1], 0x01020408u, acc);
  acc = __dp4a(fl[2], 0x10204080u, acc << 8);
  acc = __dp4a(fl[3], 0x01020408u, acc);
  acc = __dp4a(fl[4], 0x10204080u, acc << 8);
  acc = __dp4a(fl[5], 0x01020408u, acc);  // 24 flags in bits 7..30
  uint32_t last = __dp4a(fl[6], 0x10204080u, 0u);
  last = __dp4a(fl[7], 0x01020408u, last);
  return mad_fma(acc, one + one, last >> 7);
}

// class C of the 64 bytes in w: reversed orientation (bit 63-b <=> byte b)
template <int C>
__device__ __forceinline__ uint64_t class_rev64(const FlatDev& f, const uint32_t (&w)[16], uint32_t one) {
  uint32_t fl[16];
#ifdef CGX_JIT
#pragma unroll
  for (int k = 0; k < 16; k++) fl[k] = cgx_jit_flags<C>(w[k], one);  // generated: ranges as constants
#else
#pragma unroll
  for (int k = 0; k < 16; k++) fl[k] = 0;
  const int nr = f.cls_nranges[C];
  for (int r = 0; r < nr; r++) {
    const uint32_t k1 = f.cls_k1[C][r], k2 = f.cls_k2[C][r];
    if (f.cls_mode[C][r] == 0) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const uint32_t z = mad_fma(xor_and(w[k], k1, 0x7F7F7F7Fu), one, k2);  // bit7 set <=> (x^lo)&0x7f > width
        fl[k] |= nor_and(z, w[k], 0x80808080u);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; k++) fl[k] |= swar_in_range(w[k], k1, k2);
    }
  }
#endif
  return mk64(pack_rev(fl, one), pack_rev(fl + 8, one));  // bytes 0..31 in the high word
}

// Class bitmaps (reversed orientation) of the lane's 64-byte piece of the tile at `win`: 4 x LDS.128.
// Pieces lie 64 bytes apart, so if every lane read quarter j of its piece with load j the eight lanes
// of a quarter-warp would meet in two groups of four banks (a 4-way conflict: 16 wavefronts per load
// instead of 4).  Lane l therefore reads quarter j ^ r with load j, r = (l >> 1) & 3 (CGX_ROT): the
// four lanes that share a bank group take four different quarters.  The flags are packed as if load j
// were quarter j; two byte permutes per class word (per-lane selectors) put the 16-bit fields back.
struct LaneRot {
  uint32_t o0;               // lane * 64 + (r << 4): offset of the quarter load 0 reads
  uint32_t sel_lo, sel_hi;   // __byte_perm selectors that swap the 16-bit fields q <-> q ^ r
#ifdef CGX_TEDDY
  uint32_t fsel_lo, fsel_hi; // the same for a word in forward orientation (field q = bytes 2q, 2q + 1)
#endif
};
#ifndef CGX_ROT
#define CGX_ROT 1
#endif
__device__ __forceinline__ LaneRot lane_rot(int lane) {
  LaneRot lr;
  const uint32_t r = CGX_ROT ? ((uint32_t)lane >> 1) & 3u : 0u;
  lr.o0 = (uint32_t)lane * 64u + (r << 4);
  // field q (bytes 16 q .. 16 q + 15 of the piece) occupies bytes 7 - 2 q and 6 - 2 q of the word
  uint32_t sl = 0u, sh = 0u;
#pragma unroll
  for (uint32_t i = 0; i < 8u; i++) {
    const uint32_t q = (7u - i) >> 1, sub = (7u - i) & 1u;
    const uint32_t src = 7u - (2u * (q ^ r) + sub);
    if (i < 4u) sl |= src << (4u * i);
    else sh |= src << (4u * (i - 4u));
  }
  lr.sel_lo = sl;
  lr.sel_hi = sh;
#ifdef CGX_TEDDY
  uint32_t fl = 0u, fh = 0u;
#pragma unroll
  for (uint32_t i = 0; i < 8u; i++) {
    const uint32_t src = 2u * ((i >> 1) ^ r) + (i & 1u);
    if (i < 4u) fl |= src << (4u * i);
    else fh |= src << (4u * (i - 4u));
  }
  lr.fsel_lo = fl;
  lr.fsel_hi = fh;
#endif
  return lr;
}
__device__ __forceinline__ void classify_piece(const FlatDev& f, const uint8_t* win, const LaneRot& lr, uint32_t one,
                                               uint64_t (&cm)[4]) {
  uint32_t w[16];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint4 v = *reinterpret_cast<const uint4*>(win + (lr.o0 ^ (uint32_t)(j << 4)));
    w[4 * j] = v.x;
    w[4 * j + 1] = v.y;
    w[4 * j + 2] = v.z;
    w[4 * j + 3] = v.w;
  }
  cm[0] = class_rev64<0>(f, w, one);
  cm[1] = P_NCLASSES > 1 ? class_rev64<1>(f, w, one) : 0ull;
  cm[2] = P_NCLASSES > 2 ? class_rev64<2>(f, w, one) : 0ull;
  cm[3] = P_NCLASSES > 3 ? class_rev64<3>(f, w, one) : 0ull;
#if CGX_ROT
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const uint32_t lo = lo32(cm[c]), hi = hi32(cm[c]);
    cm[c] = mk64(__byte_perm(lo, hi, lr.sel_hi), __byte_perm(lo, hi, lr.sel_lo));
  }
#endif
}

// ---- marker passes, one word at a time -----------------------------------------------------------
// Per step of a pass the lane carries two registers from the word it processed before: the adder's
// carry and the high half of the step's input word (its top bits are what a shift brings in).
struct PassState {
  uint32_t hi[NSTATE];
  uint32_t cy[NSTATE];
};
__device__ __forceinline__ void pass_reset(PassState& s) {
#pragma unroll
  for (int i = 0; i < NSTATE; i++) s.hi[i] = s.cy[i] = 0u;
}

// Right to left (reversed orientation): M = positions from which items k..end can match.
// kind: 0 = one byte of the class, 1 = class+, 2 = class*, 3 = class?
template <int C>
__device__ __forceinline__ void rev_step(uint32_t kind, const uint64_t (&c)[4], uint64_t& M, uint32_t& shi, uint32_t& scy) {
  const uint64_t Cw = c[C];
  const uint64_t u = shl_in<1>(M, shi) & Cw;
  shi = hi32(M);
  if (kind == 0) {
    M = u;
  } else if (kind == 3) {
    M |= u;
  } else {
    // extend through the run towards lower addresses; a second marker inside one run survives the
    // carry of the first as a 1 in the sum, so the markers themselves are OR-ed back
    const uint64_t p = (~adc64(u, Cw, scy) & Cw) | u;
    M = kind == 1 ? p : (M | p);
  }
}
// `C+ a` seen from the right: one byte of class A, then a run of class C.  q = shl1(A) & C marks the
// run bytes that directly follow (in marker direction) a byte of A, so shl1(shl1(M) & A) & C ==
// shl2(M) & q: the two steps cost one shift.
template <int C>
__device__ __forceinline__ void rev_fused(const uint64_t (&c)[4], uint64_t q, uint64_t& M, uint32_t& shi, uint32_t& scy) {
  const uint64_t u = shl_in<2>(M, shi) & q;
  shi = hi32(M);
  M = (~adc64(u, c[C], scy) & c[C]) | u;
}
// Left to right (forward orientation): T = positions a marker stands at before item k; the forced
// greedy choice (take the whole run / take the optional byte whenever it is there).
template <int C>
__device__ __forceinline__ void fwd_step(uint32_t kind, const uint64_t (&c)[4], uint64_t& T, uint32_t& shi, uint32_t& scy) {
  const uint64_t Cw = c[C];
  const uint64_t i = T & Cw;  // markers that can take a byte
  if (kind == 0) {
    T = shl_in<1>(i, shi);
    shi = hi32(i);
  } else if (kind == 3) {
    T = (T & ~Cw) | shl_in<1>(i, shi);
    shi = hi32(i);
  } else {
    // a marker inside a run of ones carries out to the first zero after the run
    const uint64_t e = adc64(i, Cw, scy) & ~Cw;
    T = kind == 1 ? e : ((T & ~Cw) | e);
  }
}

#define CGX_CLASS_SWITCH(cls, CALL)  \
  switch (cls) {                     \
    case 0: { constexpr int C = 0; CALL; } break; \
    case 1: { constexpr int C = 1; CALL; } break; \
    case 2: { constexpr int C = 2; CALL; } break; \
    default: { constexpr int C = 3; CALL; } break; \
  }

// the right-to-left steps for one word (c = class words, reversed); returns M
__device__ __forceinline__ uint64_t rev_word(const FlatDev& f, const uint64_t (&c)[4], PassState& st) {
  uint64_t M;
#ifdef CGX_JIT
  M = c[CGX_JIT_REV_INIT];
#define CGX_QDEF(A, B, I)                                          \
  const uint64_t q##A##_##B = shl_in<1>(c[A], st.hi[I]) & c[B];    \
  st.hi[I] = hi32(c[A]);
  CGX_JIT_REV_FUSED(CGX_QDEF)
#define CGX_STEP(kind, cls, I) rev_step<cls>(kind, c, M, st.hi[I], st.cy[I]);
#define CGX_FUSE(A, B, I) rev_fused<B>(c, q##A##_##B, M, st.hi[I], st.cy[I]);
  CGX_JIT_REV_PASS(CGX_STEP, CGX_FUSE)
#undef CGX_STEP
#undef CGX_FUSE
#undef CGX_QDEF
#else
  const int init_cls = f.rev_init_class;
  M = init_cls == 0 ? c[0] : init_cls == 1 ? c[1] : init_cls == 2 ? c[2] : c[3];
  const int rev_nops = f.rev_nops;
  for (int k = 0; k < rev_nops; k++) {
    const uint32_t op = f.rev_ops[k];
    const uint32_t kind = op & 3u;
    CGX_CLASS_SWITCH(op >> 2, (rev_step<C>(kind, c, M, st.hi[k], st.cy[k])));
  }
#endif
  return M;
}
// the left-to-right steps for one word (c = class words, forward); T enters as the starts, returns the ends
__device__ __forceinline__ uint64_t fwd_word(const FlatDev& f, const uint64_t (&c)[4], uint64_t T, PassState& st) {
#ifdef CGX_JIT
#define CGX_STEP(kind, cls, I) fwd_step<cls>(kind, c, T, st.hi[I], st.cy[I]);
  CGX_JIT_FWD_PASS(CGX_STEP)
#undef CGX_STEP
#else
  const int fwd_nops = f.fwd_nops;
  for (int k = 0; k < fwd_nops; k++) {
    const uint32_t op = f.fwd_ops[k];
    const uint32_t kind = op & 3u;
    CGX_CLASS_SWITCH(op >> 2, (fwd_step<C>(kind, c, T, st.hi[k], st.cy[k])));
  }
#endif
  return T;
}

// Do the starts S and ends E of one word alternate (start, end, start, end ...; an end may share
// its position with the next start)?  With in = 1 when a match is open at the word's first bit,
// the word is consistent iff D = E - S - in, the span mask, has its edges exactly at S ^ E:
// D ^ (D << 1 | in) == S ^ E (tests/test_sim_flat.py checks the criterion exhaustively on short
// words).  The span mask's top bit says whether a match is still open after the word.
// Returns the word's discrepancy bits (0 = consistent): a sweep ORs them together and looks once.
__device__ __forceinline__ uint64_t word_misordered(uint64_t S, uint64_t E, uint32_t& in) {
  const uint64_t D = E - S - in;
  const uint64_t x = (D ^ ((D << 1) | in)) ^ (S ^ E);
  in = (uint32_t)(D >> 63);
  return x;
}

// ---- serial replay (cold) -------------------------------------------------------------------------
__device__ __forceinline__ bool in_filter(const ScanArgs& a, uint32_t b) {
  if (a.filter.kind == F_LUT) return __ldg(a.filter.lut + b) != 0;
  bool r = false;
  for (int k = 0; k < a.filter.nranges; k++) r |= (b >= a.filter.lo[k] && b <= a.filter.hi[k]);
  return r;
}
__device__ __forceinline__ bool is_sync(const ScanArgs& a, uint32_t b) {
  return (a.flat.sync_lut[b >> 5] >> (b & 31)) & 1u;
}
// anchored leftmost-first walk through global memory (reference dfa/lazy/lazy.go:219-324)
__device__ __noinline__ int64_t dfa_walk_global(const ScanArgs& a, int64_t p0) {
  unsigned s = a.dfa.start[0];
  int64_t last = -1, p = p0;
  while (s) {
    if (p >= a.n) {
      if (__ldg(a.dfa.eoi + s)) last = a.n;
      break;
    }
    const uint32_t e = __ldg(a.dfa.trans + (s << 8) + __ldg(a.h + p));
    if (e & 0x8000u) last = p;
    s = e & 0x7FFFu;
    p++;
  }
  return last;
}
// Cold: region-relative position of the first (or last) sync byte among the `len` bytes from global
// position rb on; bytes at or beyond the end of input count as sync bytes; -1 when the region
// starts the haystack (first) or holds none.
__device__ __noinline__ int region_sync_cold(const ScanArgs& a, int64_t rb, int len, bool last, bool hay_start) {
  if (!last) {
    if (hay_start) return -1;
    for (int i = 0; i < len; i++)
      if (rb + i >= a.n || is_sync(a, __ldg(a.h + rb + i))) return i;
    return len - 1;  // (not reached for a lane that owns something)
  }
  for (int i = len - 1; i >= 0; i--)
    if (rb + i >= a.n || is_sync(a, __ldg(a.h + rb + i))) return i;
  return -1;
}
__device__ __forceinline__ void smem_or64(uint64_t* w, int bit) {
#if defined(CGX_CPU_SIM)
  *w |= 1ull << bit;
#else
  atomicOr(reinterpret_cast<unsigned int*>(w) + (bit >> 5), 1u << (bit & 31));
#endif
}
// The calling lane replays the reference loop (meta/findall.go:176-290 over
// meta/find_indices.go:1050-1088) from global position `from` until the candidate search meets a
// sync byte at or after `stop_min` (or the end of input).  A match whose end lies inside the chunk's
// bitmap is recorded there (res != nullptr) — the others ("far", only the chunk's last segment can
// produce them) are counted, or, when `far_out` is given, stored from index far_idx on.
__device__ __noinline__ unsigned replay_cold(const ScanArgs& a, int64_t cb, Slot* res, int64_t from, int64_t stop_min,
                                             int64_t* far_out, unsigned long long far_idx, unsigned* nbits) {
  unsigned far = 0, bits = 0;
  const int64_t n = a.n;
  int64_t pos = from;
  while (pos < n) {
    int64_t d = pos;
    bool stop = false;
    while (d < n) {
      const uint32_t b = __ldg(a.h + d);
      if (in_filter(a, b)) break;
      if (d >= stop_min && is_sync(a, b)) {
        stop = true;
        break;
      }
      d++;
    }
    if (stop || d >= n) break;
    const int64_t e = dfa_walk_global(a, d);
    if (e >= 0) {
      const int64_t re = e - cb, rd = d - cb;
      if (re < (int64_t)WINDOW) {
        if (res) {
          const int ws = (int)(rd >> 6), we = (int)(re >> 6);
          smem_or64(&res[ws + ws / K].a, (int)(rd & 63));
          smem_or64(&res[we + we / K].b, (int)(re & 63));
        }
        bits++;
      } else {
        if (far_out && (int64_t)(far_idx + far) < a.cap) {
          far_out[2 * (far_idx + far)] = d + a.base;
          far_out[2 * (far_idx + far) + 1] = e + a.base;
        }
        far++;
      }
      pos = e > d ? e : d + 1;
    } else {
      pos = d + 1;
      // a failed run start fails for the whole run (digitRunSkipSafe, meta/strategy.go:525-560)
      if (a.flat.bs_runstart)
        while (pos < n && in_filter(a, __ldg(a.h + pos))) pos++;
    }
  }
  if (nbits) *nbits = bits;
  return far;
}

// The bit-parallel form of the replay for a lane whose last segment is merely OPEN (no sync byte in
// its neighbour's first word, nothing misordered): the calling lane classifies the words from the one
// that holds `from` (= the position after the lane's last sync byte) up to the first word with a sync
// byte at or after `stop_min`, straight from global memory, and runs both sweeps over them on its own
// — some hundred instructions per word instead of a dependent table walk per byte.  Matches whose end
// lies inside the chunk's bitmap are recorded there, the others are counted (`far`, as replay_cold
// does).  Returns false when the stretch is longer than MAXW words or turns out misordered: the
// caller then takes the reference loop (replay_cold).
constexpr int RB_MAXW = 8;
static_assert(2 * CAP * 2 >= RB_MAXW * 7 * 8, "replay_bits keeps its words in a staging buffer");
__device__ __noinline__ bool replay_bits(const ScanArgs& a, const FlatDev& f, int64_t cb, Slot* res, int64_t from,
                                         int64_t stop_min, uint32_t one, uint64_t* scratch, unsigned* far_out) {
  // (shared-memory scratch, not local arrays: a stack frame in this cold function cost the hot loop
  // 4 % on B200)
  uint64_t(*cr)[4] = reinterpret_cast<uint64_t(*)[4]>(scratch);  // class words, reversed orientation
  uint64_t* mk = scratch + RB_MAXW * 4;
  const int64_t w0 = (from - cb) >> 6;
  int nw = 0;
  int last_sync = -1;  // bit (forward orientation) of the closing sync byte in word nw - 1
  for (;;) {
    if (nw == RB_MAXW) return false;
    const int64_t wp = cb + ((w0 + nw) << 6);  // global position of the word
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (wp + 16 * j < a.n) v = __ldg(reinterpret_cast<const uint4*>(a.h + wp) + j);  // (whole 16-byte blocks, as the bulk copies)
      w[4 * j] = v.x;
      w[4 * j + 1] = v.y;
      w[4 * j + 2] = v.z;
      w[4 * j + 3] = v.w;
    }
    cr[nw][0] = class_rev64<0>(f, w, one);
    cr[nw][1] = P_NCLASSES > 1 ? class_rev64<1>(f, w, one) : 0ull;
    cr[nw][2] = P_NCLASSES > 2 ? class_rev64<2>(f, w, one) : 0ull;
    cr[nw][3] = P_NCLASSES > 3 ? class_rev64<3>(f, w, one) : 0ull;
    const int64_t valid = a.n - wp;  // bytes of this word inside the input
    if (valid < 64) {
      const uint64_t m = valid <= 0 ? 0ull : (~0ull << (64 - valid));
#pragma unroll
      for (int c = 0; c < 4; c++) cr[nw][c] &= m;
    }
    // first sync byte of this word at or after stop_min (and, in the first word, at or after `from`)
    uint64_t nz = ~brev64(cr[nw][0] | cr[nw][1] | cr[nw][2] | cr[nw][3]);
    int64_t lo = (from > stop_min ? from : stop_min) - wp;
    if (lo > 0) nz = lo >= 64 ? 0ull : (nz & (~0ull << lo));
    nw++;
    if (nz) {
      last_sync = __ffsll((long long)nz) - 1;
      break;
    }
  }
  PassState st;
  pass_reset(st);
  for (int i = nw - 1; i >= 0; i--) mk[i] = brev64(rev_word(f, cr[i], st));
  pass_reset(st);
  uint32_t in = 0u, c0hi = 0u;
  uint64_t badbits = 0ull;
  uint64_t* Sw = scratch + RB_MAXW * 5;
  uint64_t* Ew = scratch + RB_MAXW * 6;
  for (int i = 0; i < nw; i++) {
    uint64_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) c[k] = brev64(cr[i][k]);
    uint64_t own = ~0ull;
    if (i == 0) {
      const int lo = (int)(from - (cb + (w0 << 6)));
      own = lo >= 64 ? 0ull : (~0ull << lo);
    }
    if (i == nw - 1) own &= last_sync >= 63 ? ~0ull : ((2ull << last_sync) - 1ull);  // up to and including the sync byte
    uint64_t M = mk[i];
    const uint64_t c0prev = shl_in<1>(c[0], c0hi);
    c0hi = hi32(c[0]);
    if (P_RUNSTART) M &= c[0] & ~c0prev;
    const uint64_t S = M & own;
    const uint64_t E = fwd_word(f, c, S, st) & own;
    badbits |= word_misordered(S, E, in);
    if (P_MIDRUN) badbits |= E & c[0] & c0prev;
    Sw[i] = S;
    Ew[i] = E;
  }
  if (badbits != 0ull || in != 0u) return false;
  // starts and ends alternate (an end may stand where the next start does): pair them in order; a
  // match is recorded when its end is inside the chunk's bitmap, counted otherwise
  unsigned far = 0u;
  int64_t open_start = -1;
  for (int i = 0; i < nw; i++) {
    uint64_t S = Sw[i], E = Ew[i];
    const int64_t wrel = (w0 + i) << 6;
    for (;;) {
      if (open_start < 0) {
        if (!S) break;
        open_start = wrel + (__ffsll((long long)S) - 1);
        S &= S - 1ull;
      } else {
        if (!E) break;
        const int64_t e = wrel + (__ffsll((long long)E) - 1);
        E &= E - 1ull;
        if (e < (int64_t)WINDOW) {
          const int ws = (int)(open_start >> 6), we = (int)(e >> 6);
          smem_or64(&res[ws + ws / K].a, (int)(open_start & 63));
          smem_or64(&res[we + we / K].b, (int)(e & 63));
        } else {
          far++;
        }
        open_start = -1;
      }
    }
  }
  *far_out = far;
  return true;
}

#ifdef CGX_TEDDY
// ---- multi-literal engine ---------------------------------------------------------------------------
// Candidate bitmap (FORWARD orientation: bit b <=> byte b) of the lane's 64-byte piece: position p is
// a candidate when some bucket holds a literal whose first two bytes are h[p], h[p+1]
// (tfa[h[p]] & tfb[h[p+1]] != 0).  The four quarter loads are rotated like classify_piece's; the
// quarters are evaluated one by one (the byte after a quarter comes from shared memory), packed as
// if load j were quarter j and put in place by two byte permutes.
__device__ __forceinline__ uint64_t teddy_piece(const uint8_t* win, const LaneRot& lr, const uint16_t* tfa,
                                                const uint16_t* tfb, const uint16_t* tfc) {
  uint32_t f16[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint32_t qoff = lr.o0 ^ (uint32_t)(j << 4);  // offset of the quarter load j reads
    const uint4 v = *reinterpret_cast<const uint4*>(win + qoff);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    // the two bytes after the quarter; those after the tile's last byte are not here: every bucket passes
    const bool more = qoff + 16u < (uint32_t)TILE;
    const uint32_t n1 = more ? (uint32_t)win[qoff + 16u] : 0u, n2 = more ? (uint32_t)win[qoff + 17u] : 0u;
    uint32_t bb[18];
#pragma unroll
    for (int k = 0; k < 16; k++) bb[k] = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
    bb[16] = n1;
    bb[17] = n2;
    uint32_t m = 0u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const uint32_t ta = tfa[bb[k]];
      const uint32_t tb = (k + 1 < 16 || more) ? (uint32_t)tfb[bb[k + 1]] : 0xFFFFu;
      const uint32_t tc = (k + 2 < 16 || more) ? (uint32_t)tfc[bb[k + 2]] : 0xFFFFu;
      if (ta & tb & tc) m |= 1u << k;
    }
    f16[j] = m;
  }
  const uint32_t lo = f16[0] | (f16[1] << 16), hi = f16[2] | (f16[3] << 16);
#if CGX_ROT
  return mk64(__byte_perm(lo, hi, lr.fsel_hi), __byte_perm(lo, hi, lr.fsel_lo));
#else
  return mk64(hi, lo);
#endif
}

// bytes p .. p+7 of the haystack as a little-endian word (global memory; bytes at or beyond n read
// as anything: every comparison is guarded by p + len <= n)
__device__ __forceinline__ uint64_t teddy_load8(const ScanArgs& a, int64_t p) {
  const int64_t q = p & ~(int64_t)7;
  if (q + 16 <= ((a.n + 15) & ~(int64_t)15)) {  // both words inside the readable 16-byte blocks
    const uint64_t w0 = __ldg(reinterpret_cast<const unsigned long long*>(a.h + q));
    const uint64_t w1 = __ldg(reinterpret_cast<const unsigned long long*>(a.h + q) + 1);
    const unsigned sh = (unsigned)(p & 7) * 8u;
    return sh ? (w0 >> sh) | (w1 << (64u - sh)) : w0;
  }
  uint64_t v = 0;
  for (int k = 0; k < 8; k++)
    if (p + k < a.n) v |= (uint64_t)__ldg(a.h + p + k) << (8 * k);
  return v;
}
// the engine's shared-memory tables (CtaSmem)
struct TeddyTab {
  const uint16_t *tfa, *tfb, *tfc;
  const unsigned long long *lit8, *msk8;
  const uint16_t *len, *order, *boff;
};
__device__ __forceinline__ bool teddy_lit_equal(const ScanArgs& a, const TeddyTab& tt, int64_t p, int id, uint64_t hay8,
                                                int& len) {
  len = tt.len[id];
  if ((hay8 ^ tt.lit8[id]) & tt.msk8[id]) return false;
  if (p + len > a.n) return false;
  if (len > 8) {
    const int o = __ldg(a.teddy.offs + id);
    for (int k = 8; k < len; k++)
      if (__ldg(a.h + p + k) != __ldg(a.teddy.bytes + o + k)) return false;
  }
  return true;
}
// Which literal stands at p?  Returns the match end or -1.  SIMD regime: buckets low to high,
// insertion order inside a bucket (reference prefilter/teddy.go:415-428, :532-550); scalar regime
// (fewer than 16 bytes left from the search start, :447-458): plain literal order.  One round trip
// to the haystack (8 bytes); masks and literals come from shared memory.
__device__ __noinline__ int64_t teddy_verify_g(const ScanArgs& a, const TeddyTab& tt, int64_t p, bool scalar) {
  if (p + 3 > a.n) return -1;  // (every literal has three bytes or more)
  const uint64_t hay8 = teddy_load8(a, p);
  uint32_t mask = (uint32_t)tt.tfa[hay8 & 0xFFu] & tt.tfb[(hay8 >> 8) & 0xFFu] & tt.tfc[(hay8 >> 16) & 0xFFu];
  if (!mask) return -1;
  int len;
  if (scalar) {
    for (int id = 0; id < a.teddy.npat; id++)
      if (teddy_lit_equal(a, tt, p, id, hay8, len)) return p + len;
    return -1;
  }
  while (mask) {
    const int b = __ffs((int)mask) - 1;
    mask &= mask - 1u;
    if (b >= a.teddy.nbuckets) break;
    const int k1 = tt.boff[b + 1];
    for (int k = tt.boff[b]; k < k1; k++)
      if (teddy_lit_equal(a, tt, p, (int)tt.order[k], hay8, len)) return p + len;
  }
  return -1;
}
// records the match [s, e) in the chunk's bitmaps (positions relative to the window origin cb)
__device__ __forceinline__ void teddy_record(Slot* cls0, uint64_t* sbits, int64_t cb, int64_t s, int64_t e) {
  const int ws = (int)((s - cb) >> 6), we = (int)((e - cb) >> 6);
  smem_or64(&sbits[ws + ws / K], (int)((s - cb) & 63));
  smem_or64(&cls0[we + we / K].b, (int)((e - cb) & 63));
}
// Cold: the reference loop for the starts in [x0, x1) through global memory, with the verify order
// that the distance from each search start to the end of the haystack asks for.  Used near the end
// of the haystack and when the word before a lane's region offers no safe point.  It finds its own
// safe point by walking back 32 bytes at a time: a position q is safe when no span that starts in
// [q - 64, q) crosses it (spans that start earlier end before it: literals are at most 32 bytes
// long); otherwise any position of [q - 32, q) strictly inside none of those spans is.
__device__ __forceinline__ bool teddy_is_candidate(const ScanArgs& a, int64_t p) {
  return p + 2 <= a.n &&
         ((__ldg(a.teddy.fp + __ldg(a.h + p)) & 0xFFFFu) & (__ldg(a.teddy.fp + __ldg(a.h + p + 1)) >> 16)) != 0u;
}
__device__ __noinline__ void teddy_lane_cold(const ScanArgs& a, const TeddyTab& tt, int64_t cb, Slot* cls0,
                                             uint64_t* sbits, int64_t x0, int64_t x1) {
  int64_t pos = 0;
  for (int64_t q = x0; q > 0; q -= 32) {
    const int64_t lo = q >= 64 ? q - 64 : 0;
    uint64_t covered = 0ull;  // bit i <=> position lo + i lies strictly inside a span
    bool cross = false;
    for (int64_t p = lo; p < q; p++) {
      if (!teddy_is_candidate(a, p)) continue;
      const int64_t e = teddy_verify_g(a, tt, p, false);
      if (e < 0) continue;
      for (int64_t i = p + 1; i < e && i < q; i++) covered |= 1ull << (i - lo);
      if (e > q) cross = true;
    }
    if (!cross) {
      pos = q;
      break;
    }
    int64_t u = -1;
    for (int64_t i = q - 1; i >= q - 32 && i >= lo; i--)
      if (!((covered >> (i - lo)) & 1ull)) {
        u = i;
        break;
      }
    if (u >= 0) {
      pos = u;
      break;
    }
  }
  for (int64_t p = pos; p < x1 && p + 2 <= a.n; p++) {
    if (!teddy_is_candidate(a, p)) continue;
    const int64_t e = teddy_verify_g(a, tt, p, a.n + a.after - pos < 16);
    if (e < 0) continue;
    if (p >= x0) teddy_record(cls0, sbits, cb, p, e);
    pos = e;
    p = e - 1;
  }
}
// Phase B of a lane: the words [own_lo, own_hi) of the window at cb are its own, word own_lo - 1
// (if any) is where it looks for its safe point.  cand: the chunk's candidate bitmap (slot .a).
__device__ __forceinline__ void teddy_lane(const ScanArgs& a, const TeddyTab& tt, int64_t cb, Slot* cls0,
                                           uint64_t* sbits, int own_lo, int own_hi) {
  if (own_hi <= own_lo) return;
  const int64_t x0 = cb + (int64_t)own_lo * 64, x1 = cb + (int64_t)own_hi * 64;
  if (x0 >= a.n) return;
  // near the end of the haystack the verify order depends on the search start: exact replay
  if (a.n + a.after - (x0 - 64) < 16 + 64 + (int64_t)(own_hi - own_lo) * 64 + 64) {
    teddy_lane_cold(a, tt, cb, cls0, sbits, x0, x1);
    return;
  }
  int64_t pos = x0;
  int w = own_lo;
  if (own_lo > 0) {
    // spans of the word before the region, unchained
    const int pw = own_lo - 1;
    const int64_t wp = cb + (int64_t)pw * 64;
    uint64_t c = cls0[pw + pw / K].a, covered = 0ull;
    bool cross = false;
    while (c) {
      const int b = __ffsll((long long)c) - 1;
      c &= c - 1ull;
      const int64_t e = teddy_verify_g(a, tt, wp + b, false);
      if (e < 0) continue;
      const int len = (int)(e - (wp + b));
      // positions strictly inside the span: b + 1 .. b + len - 1
      if (len > 1) {
        const uint64_t from = b + 1 >= 64 ? 0ull : (~0ull << (b + 1));
        const uint64_t upto = b + len - 1 >= 63 ? ~0ull : ((2ull << (b + len - 1)) - 1ull);
        covered |= from & upto;
      }
      if (b + len > 64) cross = true;  // x0 lies strictly inside
    }
    if (cross) {
      const uint64_t safe = ~covered & 0xFFFFFFFF00000000ull;  // (a literal is at most 32 bytes long)
      if (!safe) {
        teddy_lane_cold(a, tt, cb, cls0, sbits, x0, x1);
        return;
      }
      pos = wp + (63 - __clzll((long long)safe));
      w = pw;
    }
  }
  for (; w < own_hi; w++) {
    const int64_t wp = cb + (int64_t)w * 64;
    uint64_t c = cls0[w + w / K].a;
    if (wp < pos) c = pos - wp >= 64 ? 0ull : (c & (~0ull << (pos - wp)));
    while (c) {
      const int b = __ffsll((long long)c) - 1;
      c &= c - 1ull;
      const int64_t s = wp + b;
      if (s < pos) continue;  // inside the match before
      const int64_t e = teddy_verify_g(a, tt, s, false);
      if (e < 0) continue;
      if (s >= x0) teddy_record(cls0, sbits, cb, s, e);
      pos = e;
    }
  }
}
#endif  // CGX_TEDDY

// ---- two-level look-back ------------------------------------------------------------------------------
// Chunk c publishes its match count in status[c] (flag LB_AGG).  Chunks form groups of 32; the
// warp that finishes a group's last chunk publishes the group's sum in gstatus[g] (LB_AGG), and
// whoever learns the prefix at a group's start publishes it as gstatus[g-1] (LB_PREFIX, the
// inclusive prefix through group g-1).  The exclusive prefix of chunk c is therefore: counts of the
// earlier chunks of its own group (one 32-wide load) + a decoupled look-back over GROUP words.
// Words carry the launch's epoch (bits 42..61): a word of another epoch reads as empty, so nothing
// has to be cleared between launches.
constexpr int EP_SHIFT = 42;
constexpr unsigned long long EP_MASK = 0xFFFFFull << EP_SHIFT;
constexpr unsigned long long VAL_MASK = (1ull << EP_SHIFT) - 1;
__device__ __forceinline__ unsigned long long ep_word(const ScanArgs& a, unsigned long long flag, unsigned long long v) {
  return flag | ((unsigned long long)a.epoch << EP_SHIFT) | v;
}
// flag of a word as this launch sees it (0 = empty / stale)
__device__ __forceinline__ unsigned ep_flag(const ScanArgs& a, unsigned long long v) {
  return ((v & EP_MASK) >> EP_SHIFT) == (unsigned long long)a.epoch ? (unsigned)(v >> 62) : 0u;
}
__device__ unsigned long long lb_resolve(const ScanArgs& a, int64_t chunk, int lane) {
  const int64_t g = chunk >> 5;
  const int64_t idx1 = (g << 5) + lane;
  unsigned long long excl = 0, gpre = 0;
  bool have1 = false, have2 = g == 0;
  int64_t look = g - 1;
  for (;;) {
    // both levels are requested before either is looked at: one L2 round trip per attempt
    unsigned f1 = 1u, f2 = 2u;          // lanes beyond the chunk contribute nothing; before group 0: zero prefix
    unsigned long long v1 = 0, v = 0;
    if (!have1 && idx1 < chunk) {
      v1 = ld_status(&a.status[idx1]);
      f1 = ep_flag(a, v1);
    }
    const int64_t idx2 = look - lane;
    if (!have2 && idx2 >= 0) {
      v = ld_status(&a.gstatus[idx2]);
      f2 = ep_flag(a, v);
    }
    if (!have1 && __all_sync(FULL, f1 != 0u)) {
      excl = __reduce_add_sync(FULL, (unsigned)(v1 & 0xFFFFFFFFull));
      have1 = true;
    }
    if (!have2) {
      const uint32_t empty = __ballot_sync(FULL, f2 == 0u);
      const uint32_t pm = __ballot_sync(FULL, f2 == 2u);
      const int fe = empty ? __ffs((int)empty) - 1 : 32;
      const int fp = pm ? __ffs((int)pm) - 1 : 32;
      if (fp < fe) {
        // sums of the groups before the prefix, then the prefix itself
        const unsigned part = __reduce_add_sync(FULL, lane < fp ? (unsigned)(v & 0xFFFFFFFFull) : 0u);
        gpre += part + __shfl_sync(FULL, v & VAL_MASK, fp);
        have2 = true;
        // the inclusive prefix through the previous group, for everybody behind us
        if (lane == 0 && (look != g - 1 || fp != 0)) st_status(&a.gstatus[g - 1], ep_word(a, LB_PREFIX, gpre));
      } else if (fe > 0) {
        gpre += __reduce_add_sync(FULL, lane < fe ? (unsigned)(v & 0xFFFFFFFFull) : 0u);
        look -= fe;
        if (fe == 32) continue;  // a full window of sums: keep walking without a pause
      }
    }
    if (have1 && have2) return excl + gpre;
    cgx_backoff();
  }
}

// publishes the chunk's count; the warp that completes a group publishes the group's sum.  One
// 64-bit atomic carries both the arrival count (bits 40..) and the running sum (bits 0..39), so the
// last arrival knows the group's total without re-reading anything — and puts the accumulator back
// to zero for the next launch.
__device__ __forceinline__ void publish_count(const ScanArgs& a, int64_t chunk, unsigned cnt, int64_t nunits) {
  {
    const int64_t g = chunk >> 5;
    const int64_t members = nunits - (g << 5) < 32 ? nunits - (g << 5) : 32;
    st_status(&a.status[chunk], ep_word(a, LB_AGG, cnt));
    const unsigned long long old = atomicAdd(&a.gacc[g], (1ull << 40) | cnt);
    if ((int64_t)(old >> 40) == members - 1) {
      st_status(&a.gstatus[g], ep_word(a, LB_AGG, (old & ((1ull << 40) - 1)) + cnt));
      a.gacc[g] = 0ull;
    }
  }
}

// The resolver warp (FindAll).  The unit of the look-back is the GANG: FW_WARPS consecutive chunks
// scanned by the warps of one CTA.  Per gang the resolver waits until every warp has handed its
// count over, turns the counts into offsets within the gang (one warp scan), resolves the gang's
// global offset by the two-level look-back — one L2 round trip per gang, not per chunk — and hands
// every warp the global index of its chunk's first match; the scanning warps store their staged
// matches themselves when they next need the buffer.  It also draws the CTA's gang tickets — one
// gang per ticket: handing a CTA RUNS of consecutive gangs (so that only a run's first gang looks
// back) serialises the CTAs behind each other's runs with two buffers per warp (measured: 44 GB/s).
__device__ void resolver_warp(const ScanArgs& a, CtaSmem& cs, int lane, unsigned ngangs) {
  auto fetch = [&](unsigned seq) -> unsigned {
    unsigned t = 0;
    if (lane == 0) {
      t = atomicAdd(a.ticket, 1u);
      // every CTA draws exactly one ticket past the last gang; the very last draw of the launch
      // puts the counter back to zero
      if (t == ngangs + gridDim.x - 1u) *a.ticket = 0u;
      cs.tk[seq & 3u] = ((unsigned long long)(seq + 1u) << 32) | t;
    }
    return __shfl_sync(FULL, t, 0);
  };
  unsigned g = fetch(0u);
  unsigned gn = g < ngangs ? fetch(1u) : 0xFFFFFFFFu;
  for (unsigned i = 0; g < ngangs; i++) {
    const unsigned gnn = gn < ngangs ? fetch(i + 2u) : 0xFFFFFFFFu;
    const int b = (int)(i & 1u);
    for (;;) {
      int st = 1;
      if (lane < FW_WARPS) st = cs.mail[lane][b].state;
      if (__all_sync(FULL, st == 1)) break;
      cgx_backoff();
    }
    cgx_fence_block();
    const unsigned cnt = lane < FW_WARPS ? cs.mail[lane][b].cnt : 0u;
    unsigned x = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned y = __shfl_up_sync(FULL, x, d);
      if (lane >= d) x += y;
    }
#ifdef CGX_EXP_NOLB
    // experiment only (wrong output order): what the kernel would do if offsets cost nothing
    const unsigned long long excl = (unsigned long long)g * FW_WARPS * 100ull;
#else
    const unsigned long long excl = lb_resolve(a, (int64_t)g, lane);
#endif
    if (lane < FW_WARPS) {
      Mail& m = cs.mail[lane][b];
      m.excl = excl + (x - cnt);
      cgx_fence_block();
      m.state = 2;
    }
    g = gn;
    gn = gnn;
  }
}

// ---- output: bitmaps -> int64 pairs in global match order ---------------------------------------------
// Lane l turns the bits of its K words into positions: the i-th start of the chunk goes to
// out[2 (excl + i)], the i-th end to out[2 (excl + i) + 1].  Starts and ends are independent
// streams (a match that starts in one lane's words may end in the next lane's).
// One 32-bit half of a bitmap word: a store per set bit.  CHECK: the output may be too small.
template <bool CHECK>
__device__ __forceinline__ void emit_half(uint32_t w, int64_t pos0, int64_t*& o, const int64_t* end) {
  while (w) {
    const int b = __ffs((int)w) - 1;
    w &= w - 1u;
    if (!CHECK || o < end) *o = pos0 + b;
    o += 2;
  }
}
__device__ __forceinline__ void stage_half(uint32_t w, int pos0, uint16_t*& o) {
  while (w) {
    const int b = __ffs((int)w) - 1;
    w &= w - 1u;
    *o++ = (uint16_t)(pos0 + b);
  }
}
// bitmaps -> global memory (a chunk that did not fit the staging buffer; its offset is known)
template <bool CHECK>
__device__ __forceinline__ void extract_words(const ScanArgs& a, const Slot* res, int64_t wb, int64_t* os, int64_t* oe) {
  const int64_t* end = a.out + 2 * a.cap;
#pragma unroll 1
  for (int j = 0; j < K; j++) {
    const Slot v = res[j];
    const int64_t pb = wb + j * 64;
    emit_half<CHECK>(lo32(v.a), pb, os, end);
    emit_half<CHECK>(hi32(v.a), pb + 32, os, end);
    emit_half<CHECK>(lo32(v.b), pb, oe, end);
    emit_half<CHECK>(hi32(v.b), pb + 32, oe, end);
  }
}
// rk / rke: starts / ends before this lane's words
__device__ __forceinline__ void extract_direct(const ScanArgs& a, const Slot* arr, int64_t chunk,
                                               unsigned long long excl, unsigned cnt, uint32_t rk, uint32_t rke, int lane) {
  const Slot* res = arr + lane * (K + 1);
  int64_t* os = a.out + 2 * (excl + rk);
  int64_t* oe = a.out + 2 * (excl + rke) + 1;
  const int64_t wb = chunk_origin(chunk) + a.base + (int64_t)lane * (K * 64);
  if ((int64_t)(excl + cnt) <= a.cap) extract_words<false>(a, res, wb, os, oe);  // the usual case: everything fits
  else extract_words<true>(a, res, wb, os, oe);
}
// bitmaps -> staging buffer sb: chunk-relative offsets in match order.  Starts: every lane walks the
// set bits of its own K words (rk = starts before them).  Ends: starts and ends alternate, so the end
// of the i-th match is the first end bit after the i-th start — one lane per MATCH, which spreads
// the work evenly over the warp however the matches are distributed over the lanes' words.
// Stages the matches with ranks [r0, r0 + n) of the chunk (r0 = 0 and n = all of them unless CGX_PARK
// takes a dense chunk in rounds).
__device__ __forceinline__ void extract_staged(WarpSmem& ws, const Slot* all, int sb, uint32_t rk, unsigned r0, unsigned n,
                                               int lane) {
  const Slot* res = all + lane * (K + 1);
  const int wb = lane * (K * 64);
  uint16_t* os = ws.st[sb][0];
  if (r0 == 0u && !CGX_PARK) {
    os += rk;
#pragma unroll 1
    for (int j = 0; j < K; j++) {
      const uint64_t v = res[j].a;
      const int pb = wb + j * 64;
      stage_half(lo32(v), pb, os);
      stage_half(hi32(v), pb + 32, os);
    }
  } else {
    // only the starts whose rank falls into the round's window
    unsigned i = rk;
#pragma unroll 1
    for (int j = 0; j < K && i < r0 + n; j++) {
      uint64_t v = res[j].a;
      const unsigned c = (unsigned)__popcll(v);
      if (i + c > r0) {
        const int pb = wb + j * 64;
        while (v) {
          const uint32_t lo = lo32(v);
          const int b = lo ? __ffs((int)lo) - 1 : 31 + __ffs((int)hi32(v));
          v &= v - 1ull;
          if (i >= r0 && i < r0 + n) os[i - r0] = (uint16_t)(pb + b);
          i++;
        }
      } else {
        i += c;
      }
    }
  }
  __syncwarp();
  const uint16_t* ss = ws.st[sb][0];
  uint16_t* se = ws.st[sb][1];
#pragma unroll 1
  for (unsigned i = lane; i < n; i += 32) {
    const unsigned p = (unsigned)ss[i] + 1u;  // a match is not empty: its end lies after its start
    unsigned w = p >> 6;
    uint64_t bits = all[w + w / K].b & (~0ull << (p & 63u));
    while (bits == 0ull && w < (unsigned)NWORDS - 1u) {
      w++;
      bits = all[w + w / K].b;
    }
    const uint32_t lo = lo32(bits), hi = hi32(bits);
    se[i] = (uint16_t)(w * 64u + (lo ? (unsigned)__ffs((int)lo) - 1u : 31u + (unsigned)__ffs((int)hi)));
  }
}
// staged matches of a resolved chunk -> int64 pairs in global match order (coalesced 16-byte stores)
__device__ __forceinline__ void write_out(const ScanArgs& a, const WarpSmem& ws, int sb, int64_t chunk, unsigned n,
                                          unsigned long long excl, int lane) {
  const int64_t b = chunk_origin(chunk) + a.base;
  const uint16_t* ss = ws.st[sb][0];
  const uint16_t* se = ws.st[sb][1];
  if ((int64_t)(excl + n) <= a.cap) {  // the usual case: the whole chunk fits the output
    longlong2* o = reinterpret_cast<longlong2*>(a.out) + excl;
    for (unsigned i = lane; i < n; i += 32) o[i] = make_longlong2(b + ss[i], b + se[i]);
  } else {
    for (unsigned i = lane; i < n; i += 32) {
      const unsigned long long gi = excl + i;
      if ((int64_t)gi < a.cap) *reinterpret_cast<longlong2*>(a.out + 2 * gi) = make_longlong2(b + ss[i], b + se[i]);
    }
  }
}

}  // namespace

#ifdef CGX_JIT
#ifndef CGX_CPU_SIM
// the loader (jit.cu) reads the launch shape from the module it just built
extern "C" __device__ const int cgx_flat_jit_info[5] = {(int)sizeof(CtaSmem), FW_THREADS, FW_WARPS, FW_CTAS, CHUNKB};
#endif
extern "C" __global__ void __launch_bounds__(FW_THREADS, FW_CTAS) cgx_flat_jit(const __grid_constant__ ScanArgs a) {
#else
namespace {
__global__ void __launch_bounds__(FW_THREADS, FW_CTAS) scan_flat_kernel(const __grid_constant__ ScanArgs a) {
#endif
  CGX_DYN_SMEM(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  CtaSmem& cs = *reinterpret_cast<CtaSmem*>(smem_raw);
  const FlatDev& f = a.flat;
  if (tid < 2 * FW_WARPS) cs.mail[tid >> 1][tid & 1].state = 0;
  if (tid < 4) cs.tk[tid] = 0ull;
  if (tid < 2) cs.gang_acc[tid] = 0ull;
#ifdef CGX_TEDDY
  for (int i = tid; i < 256; i += FW_THREADS) {
    const uint32_t t = a.teddy.fp[i];
    cs.tfa[i] = (uint16_t)(t & 0xFFFFu);
    cs.tfb[i] = (uint16_t)(t >> 16);
    cs.tfc[i] = a.teddy.fp2[i];
  }
  for (int i = tid; i < a.teddy.npat && i < 64; i += FW_THREADS) {
    cs.tlit8[i] = reinterpret_cast<const unsigned long long*>(a.teddy.lit8)[i];
    cs.tmsk8[i] = reinterpret_cast<const unsigned long long*>(a.teddy.lit8)[a.teddy.npat + i];
    cs.tlen[i] = (uint16_t)(a.teddy.offs[i + 1] - a.teddy.offs[i]);
    cs.torder[i] = a.teddy.order[i];
  }
  for (int i = tid; i <= a.teddy.nbuckets && i < 17; i += FW_THREADS) cs.tboff[i] = a.teddy.bucket_off[i];
#endif
  static_assert(2 * FW_WARPS <= FW_THREADS, "one thread per mail slot at start-up");
  if (warp < FW_WARPS && lane == 0) {
#pragma unroll
    for (int b = 0; b < NB; b++) mbar_init(&cs.w[warp].mbar[b], 1);
    fence_mbar_init();
    for (int b = 0; b < NBUF; b++)
      for (int q = 0; q < NPAIR; q++) cs.w[warp].cls[b][q][NSLOTS] = Slot{0ull, 0ull};
    cs.w[warp].mk[NSLOTS] = 0ull;
  }
  cgx_syncthreads();
  const unsigned nch = (unsigned)a.nchunks;
  const unsigned ngangs = (nch + FW_WARPS - 1u) / FW_WARPS;
  if (warp == FW_WARPS) {
    if (P_MODE == M_FINDALL) resolver_warp(a, cs, lane, ngangs);
    return;
  }
  WarpSmem& ws = cs.w[warp];
  const unsigned nwarps_total = gridDim.x * FW_WARPS;
  // FindAll: tickets are gangs (drawn by the resolver), warp w scans chunk gang * FW_WARPS + w — a
  // chunk past the end is empty; the other modes need no order: every warp draws chunks on its own
  const bool gangs = P_MODE == M_FINDALL;
  const unsigned nunits = gangs ? ngangs : nch;
  // a 1 the compiler cannot fold (a launch has at least one chunk): multiplier of mad_fma
  const uint32_t one = (uint32_t)(a.nchunks > 0);

  // tickets: every scanning warp draws until it gets one past the last chunk; the warp that draws
  // the very last ticket of the launch puts the counter back to zero
  unsigned seq = 0;  // FindAll: the next ticket is the CTA's seq-th gang
  auto take_ticket = [&]() -> unsigned {
    unsigned t = 0;
    if (gangs) {
      if (lane == 0) {
        unsigned long long v;
        for (;;) {
          v = cs.tk[seq & 3u];
          if ((unsigned)(v >> 32) == seq + 1u) break;
          cgx_backoff();
        }
        t = (unsigned)v;
      }
      seq++;
      return __shfl_sync(FULL, t, 0);
    }
    if (lane == 0) {
      if (P_MODE == M_ISMATCH && *((volatile unsigned long long*)&a.total[1])) {
        t = 0xFFFFFFFFu;
      } else {
        t = atomicAdd(a.ticket, 1u);
        if (P_MODE != M_ISMATCH && t == nch + nwarps_total - 1u) *a.ticket = 0u;
      }
    }
    return __shfl_sync(FULL, t, 0);
  };
  // ---- window ring: tile number `seq` (running count per warp) lives in buffer seq % NB ----
  // The prefetcher runs NB tiles ahead of the classifier, across chunk borders: pf_src walks through
  // the chunk being prefetched (lane 0's copy is the one used), pf_left counts its tiles still to
  // be requested, pf_whole says that all of its windows lie inside the input (no bounds to look at).
  // (a chunk is a whole number of trips around the ring, so tile t of every chunk lives in buffer
  // t % NB — the buffer index is a compile-time constant in the unrolled tile loop — and a request
  // always goes to the buffer whose tile has just been classified)
  static_assert(TPC % NB == 0, "a chunk is a whole number of trips around the window ring");
  uint32_t rpar = 0u;       // parity of the phase the reader waits for
  const uint8_t* pf_src = a.h;
  int pf_left = 0;
  bool pf_whole = false;
  auto prefetch_chunk = [&](unsigned chunk) {
    const int64_t cbeg = chunk_origin(chunk);
    pf_src = a.h + cbeg;
    pf_left = TPC;
    pf_whole = cbeg + WINDOW <= a.n;
  };
  auto issue_next = [&](const int b) {
    if (lane == 0) {
      if (pf_whole) {  // the common case: a whole tile
        mbar_expect_tx(&ws.mbar[b], (uint32_t)TILE);
        tma_load_1d(ws.win[b], pf_src, (uint32_t)TILE, &ws.mbar[b]);
      } else {
        const int64_t left = a.n - (pf_src - a.h);
        if (left > 0) {
          // whole 16-byte blocks: the last block may run up to 15 bytes past n, inside the caller's
          // 16-byte aligned allocation granule; those bytes are masked out below (nv)
          const uint32_t bulk = left >= TILE ? (uint32_t)TILE : (uint32_t)((left + 15) & ~(int64_t)15);
          mbar_expect_tx(&ws.mbar[b], bulk);
          tma_load_1d(ws.win[b], pf_src, bulk, &ws.mbar[b]);
        } else {
          mbar_arrive(&ws.mbar[b]);
        }
      }
    }
    pf_src += TILE;
    pf_left--;
  };
  // the last chunk's owner reports the totals
  auto report_total = [&](int64_t chunk, unsigned long long total) {
    if (lane == 0 && chunk == (int64_t)nch - 1) {
      a.total[0] = total;
      a.total[1] = total ? 1ull : 0ull;
      if (a.result) {
        a.result[0] = total;
        a.result[1] = total ? 1ull : 0ull;
      }
    }
  };
  // Stores the matches staged in buffer b (if any) once the resolver has supplied the chunk's
  // offset; returns whether the buffer is free afterwards.  Blocking, or one look.
  auto flush = [&](int b, bool block) -> bool {
    Mail& m = cs.mail[warp][b];
    for (;;) {
      int st = 0;
      if (lane == 0) st = m.state;
      st = __shfl_sync(FULL, st, 0);
      if (st == 0) return true;
      if (st == 2) break;
      if (!block) return false;
      cgx_backoff();
    }
    cgx_fence_block();
    const int64_t chunk = m.chunk;
    const unsigned long long excl = m.excl;
    report_total(chunk, excl + m.cnt);
    unsigned nbits = ws.bits_cnt[b];
    if (CGX_PARK) {
      // the chunk's bitmaps waited in slot array b: rounds of CAP matches through the staging buffer
      nbits = m.cnt - ws.far_cnt[b];
      const Slot* arr = ws.cls[b][0];
      uint32_t x = 0u;
      for (int j = 0; j < K; j++) x += __popcll(arr[lane * (K + 1) + j].a);
      const uint32_t mine = x;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
      }
      x -= mine;
      for (unsigned r0 = 0; r0 < nbits; r0 += (unsigned)CAP) {
        const unsigned nr = nbits - r0 < (unsigned)CAP ? nbits - r0 : (unsigned)CAP;
        extract_staged(ws, arr, 0, x, r0, nr, lane);
        __syncwarp();
        write_out(a, ws, 0, chunk, nr, excl + r0, lane);
        __syncwarp();
      }
    } else {
      write_out(a, ws, b, chunk, nbits, excl, lane);
    }
    // matches of a replayed segment that end beyond the bitmap: found again, stored after the others
    if (lane == 0 && ws.far_cnt[b])
      replay_cold(a, chunk_origin(chunk), nullptr, ws.far_from[b], ws.far_stop[b], a.out, excl + nbits, nullptr);
    __syncwarp();
    if (lane == 0) m.state = 0;
    __syncwarp();
    return true;
  };
  // ticket -> chunk number (>= nch: nothing to scan — no ticket, or the empty end of the last gang)
  auto chunk_of = [&](unsigned t) -> unsigned {
    if (!gangs) return t;
    return t < nunits ? t * (unsigned)FW_WARPS + (unsigned)warp : 0xFFFFFFFEu;
  };
  unsigned tcur = take_ticket();
  unsigned cur = chunk_of(tcur);
  unsigned nxt = 0xFFFFFFFEu;  // none
  static_assert(NB <= TPC, "the ring never holds more than one chunk");
  if (cur < nch) {
    prefetch_chunk(cur);
#pragma unroll
    for (int t = 0; t < NB; t++) issue_next(t);
  }
  const LaneRot lrot = lane_rot(lane);
  int sb = 0;
  // FindAll: the warp's count joins its gang's; the last arrival publishes the gang's count
  auto post_count = [&](unsigned gang, unsigned cnt) {
    if (lane == 0) {
      const unsigned long long old = atomicAdd(&cs.gang_acc[sb], (1ull << 40) | cnt);
      if ((unsigned)(old >> 40) == (unsigned)FW_WARPS - 1u) {
        cs.gang_acc[sb] = 0ull;
        publish_count(a, (int64_t)gang, (unsigned)(old & ((1ull << 40) - 1ull)) + cnt, (int64_t)ngangs);
      }
    }
  };
  while (tcur < nunits) {
    const unsigned tnxt = take_ticket();
    nxt = chunk_of(tnxt);
    if (cur >= nch) {
      // (FindAll) the empty end of the last gang: the warp still takes part in the gang's hand-over
      post_count(tcur, 0u);
      flush(sb, true);
      if (lane == 0) {
        ws.far_cnt[sb] = 0u;
        ws.bits_cnt[sb] = 0u;
        Mail& m = cs.mail[warp][sb];
        m.chunk = cur;
        m.cnt = 0u;
        cgx_fence_block();
        m.state = 1;
      }
      __syncwarp();
      sb ^= 1;
      tcur = tnxt;
      cur = nxt;
      continue;
    }
    const int64_t cb = chunk_origin(cur);
    // the next chunk is known one chunk ahead: its bytes are asked into L2 now (no shared memory
    // needed for that distance), the window ring then only has to cover the L2 latency
    if (nxt < nch && lane == 0) {
      const int64_t nb = chunk_origin(nxt);
      const int64_t left = a.n - nb;
      const uint32_t bytes = left >= WINDOW ? (uint32_t)WINDOW : (uint32_t)(left & ~(int64_t)15);
      if (bytes) tma_prefetch_l2(a.h + nb, bytes);
    }
    if (P_MODE == M_FINDALL) {
      // CGX_PARK: slot array sb may still hold the bitmaps of the chunk before last
      if (CGX_PARK) flush(sb, true);
      flush(sb ^ 1, false);  // the older staged chunk, if its offset has arrived
    }
    Slot(*cls)[NSLOTS1] = ws.cls[CGX_PARK ? sb : 0];

    // ================= phase A: classify the chunk's tiles into class bitmaps ====================
    const bool whole = cb + WINDOW <= a.n;
    // word t * 32 + lane lives in slot (t * 32 + lane) + (t * 32 + lane) / K: 32 + 32 / K slots further per tile
    Slot* dst = &cls[0][lane + lane / K];
    // The tile loop exists twice: FAST for a chunk whose window — and that of the chunk the ring
    // moves on to — lies wholly inside the input (no end-of-input masks, no partial copies, the
    // ring's bookkeeping folded at compile time), and the general form.
    auto tile_loop = [&](auto fast) {
      constexpr bool FAST = decltype(fast)::value;
#pragma unroll(FAST ? FAST_UNROLL : UNROLL_A)
      for (int t0 = 0; t0 < TPC; t0 += NB) {
#pragma unroll
        for (int b = 0; b < NB; b++) {
          const int t = t0 + b;
          mbar_wait(&ws.mbar[b], rpar);
          uint64_t cm[4];
#ifdef CGX_TEDDY
          cm[0] = teddy_piece(ws.win[b], lrot, cs.tfa, cs.tfb, cs.tfc);
          cm[1] = 0ull;  // (the ends bitmap starts empty)
#else
          classify_piece(f, ws.win[b], lrot, one, cm);
#endif
          // every lane holds its piece in registers: the buffer can take the tile NB ahead
          __syncwarp();
          if (FAST) {
            // (the ring is exactly NB tiles ahead: tiles of this chunk until t + NB reaches its end,
            // then the next chunk's)
            if (t + NB == TPC && nxt < nch) prefetch_chunk(nxt);
            if (t + NB < TPC || nxt < nch) {
              if (lane == 0) {
                mbar_expect_tx(&ws.mbar[b], (uint32_t)TILE);
                tma_load_1d(ws.win[b], pf_src, (uint32_t)TILE, &ws.mbar[b]);
              }
              pf_src += TILE;
              pf_left--;
            }
          } else {
            if (pf_left == 0 && t + NB == TPC && nxt < nch) prefetch_chunk(nxt);  // the ring moves on to the next chunk
            if (pf_left) issue_next(b);
            if (!whole) {
              // bytes at or beyond the end of input belong to no class.  Reversed bit r <=> byte 63 - r
              const int64_t v = a.n - (cb + (int64_t)t * TILE + lane * 64);  // valid bytes of this piece
#ifdef CGX_TEDDY
              cm[0] &= v >= 64 ? ~0ull : (v <= 0 ? 0ull : ((1ull << v) - 1ull));  // (forward orientation)
#else
              const uint64_t m = v >= 64 ? ~0ull : (v <= 0 ? 0ull : (~0ull << (64 - v)));
#pragma unroll
              for (int c = 0; c < 4; c++) cm[c] &= m;
#endif
            }
          }
          dst[0] = Slot{cm[0], cm[1]};
          if (NPAIR > 1) dst[NSLOTS1] = Slot{cm[2], cm[3]};
#ifdef CGX_TEDDY
          ws.mk[dst - &cls[0][0]] = 0ull;  // the starts bitmap of the chunk
#endif
          dst += 32 + 32 / K;
        }
        rpar ^= 1u;
      }
    };
#ifndef CGX_FASTA
#define CGX_FASTA 1
#endif
    if (CGX_FASTA && whole && pf_left == TPC - NB && (nxt >= nch || chunk_origin(nxt) + WINDOW <= a.n)) tile_loop(BoolC<true>{});
    else tile_loop(BoolC<false>{});
    __syncwarp();

#ifdef CGX_TEDDY
    // ================= phase B: every lane runs the reference loop over its own words =================
    const int s0 = lane * (K + 1);
    unsigned cS = 0u, far = 0u;
    int64_t rp_from = 0, rp_stop = 0;
    {
      const int o = cur == 0u ? 0 : 1;  // first owned word of the window
      const int own_lo = K * lane + o;
      int own_hi = own_lo + K;
      if (own_hi > o + NWORDS - 2) own_hi = o + NWORDS - 2;
      const TeddyTab tt{cs.tfa, cs.tfb, cs.tfc, cs.tlit8, cs.tmsk8, cs.tlen, cs.torder, cs.tboff};
      teddy_lane(a, tt, cb, cls[0], ws.mk, own_lo, own_hi);
      __syncwarp();
      // the starts move next to the ends: slot = (starts, ends), as the output stage expects
      for (int j = 0; j < K; j++) {
        const uint64_t sj = ws.mk[s0 + j];
        cls[0][s0 + j].a = sj;
        cS += __popcll(sj);
      }
      __syncwarp();
    }
#else
    // ================= phase B: lane-serial marker sweeps over K words + the neighbour's first ======
    // lane 31's neighbour word would lie outside the window: its own last word is the overlap, and
    // where the others read their neighbour's word it reads the all-zero slot NSLOTS
    const int ovl = lane == 31 ? K - 1 : K;
    // slot of the lane's word 0; its words follow, then the region's pad slot, then the neighbour's word 0
    const int s0 = lane * (K + 1);
    auto slot_of = [&](int j) { return s0 + j + (j == K ? 1 : 0); };
    PassState st;
    // ---- sweep 1, right to left: where can a match start (reversed orientation) ----
    pass_reset(st);
#pragma unroll UNROLL
    for (int j = K; j >= 0; j--) {
      uint64_t c[4];
      {
        const Slot v = cls[0][slot_of(j)];
        c[0] = v.a;
        c[1] = v.b;
        c[2] = c[3] = 0ull;
        if (NPAIR > 1) {
          const Slot v2 = cls[NPAIR - 1][slot_of(j)];
          c[2] = v2.a;
          c[3] = v2.b;
        }
      }
      const uint64_t M = rev_word(f, c, st);
      // the neighbour's word 0 has been read by everybody before its owner rewrites it (at j == 0)
      if (j == K) __syncwarp();
      if (j < K) {
        // what the lane computes for its own words is what everybody uses: forward orientation
        ws.mk[s0 + j] = brev64(M);
        cls[0][s0 + j] = Slot{brev64(c[0]), brev64(c[1])};
        if (NPAIR > 1) cls[NPAIR - 1][s0 + j] = Slot{brev64(c[2]), brev64(c[3])};
      }
    }
    __syncwarp();

    // ---- sweep 2, left to right: owned starts, their ends, alternation check ----
    // Results replace the class words in place, except at the two ends of the region, whose class
    // words the neighbours still read: word 0's result waits in the region's pad slot, the overlap
    // word's result (it belongs to the neighbour's word 0) in two dead words of `mk`.
    pass_reset(st);
    uint32_t seen = (cur == 0u && lane == 0) ? FULL : 0u;  // the start of the haystack counts as a sync byte before it
    uint32_t in = 0u, c0hi = 0u;
    bool open = false;
    uint64_t badbits = 0ull;
    unsigned cS = 0u;
    static_assert(K >= 3, "the overlap word's result is parked in mk[1..2] of the lane's region");
#pragma unroll UNROLL
    for (int j = 0; j <= K; j++) {
      uint64_t c[4];
      {
        const Slot v = cls[0][slot_of(j)];
        c[0] = v.a;
        c[1] = v.b;
        c[2] = c[3] = 0ull;
        if (NPAIR > 1) {
          const Slot v2 = cls[NPAIR - 1][slot_of(j)];
          c[2] = v2.a;
          c[3] = v2.b;
        }
      }
      uint64_t M = ws.mk[slot_of(j)];
      uint64_t own = ~0ull;
      if (seen == 0u || j == ovl) {
        uint64_t U = c[0] | c[1];
        if (NC > 2) U |= c[2] | c[3];
        const uint64_t nz = ~U;                  // sync bytes of this word
        const uint64_t upto = nz ^ (nz - 1ull);  // bits up to and including the first sync byte (all ones if none)
        // owned: after the lane's first sync byte ...
        own = seen ? ~0ull : ~upto;
        if (nz && j <= ovl) seen = FULL;  // (lane 31's pass over the all-zero slot is not part of its region)
        // ... up to the first sync byte of the neighbour's first word
        if (j == ovl) {
          own &= upto;
          open = nz == 0ull;
        }
      }
      const uint64_t c0prev = shl_in<1>(c[0], c0hi);
      c0hi = hi32(c[0]);
      // only the first byte of a run of class 0 (pattern opens with C+)
      if (P_RUNSTART) M &= c[0] & ~c0prev;
      const uint64_t S = M & own;
      const uint64_t E = fwd_word(f, c, S, st) & own;
      badbits |= word_misordered(S, E, in);
      // (an end in the middle of a class-0 run: the reference resumes there, which is no run start)
      if (P_MIDRUN) badbits |= E & c[0] & c0prev;
      if (j == K) {
        ws.mk[s0 + 1] = S;
        ws.mk[s0 + 2] = E;
      } else {
        cls[0][s0 + (j == 0 ? K : j)] = Slot{S, E};
        cS += __popcll(S);
      }
    }
    // (a match still open after the last word is what an open last segment looks like: that segment
    // is replayed anyway, nothing is misordered)
    const bool bad = badbits != 0ull || (in != 0u && !open);
    // ---- lanes that need the reference loop: clear what the sweeps left in the affected range ----
    const bool replay = bad || (open && seen);
#ifdef CGX_DEBUG_PRINT
    if (replay) printf("chunk %u lane %d bad %d open %d seen %u in %u\n", cur, lane, (int)bad, (int)open, seen, in);
#endif
    int64_t rp_from = 0, rp_stop = 0;
    if (replay) {
      // bad: everything the lane owns; open only: the segment after the lane's last sync byte
      const int64_t rb = cb + (int64_t)lane * (K * 64);
      const int keep = region_sync_cold(a, rb, ovl * 64, /*last=*/!bad, cur == 0u && lane == 0);  // bits <= keep stay
      rp_from = rb + keep + 1;
      rp_stop = bad ? rb + (int64_t)ovl * 64 : rp_from;
      if (rp_stop < rp_from) rp_stop = rp_from;
      cS = 0u;
      for (int j = 0; j <= ovl; j++) {
        const int lo = keep + 1 - j * 64;  // first bit of word j to clear
        const uint64_t m = lo <= 0 ? 0ull : (lo >= 64 ? ~0ull : ((1ull << lo) - 1ull));
        if (j == K) {
          ws.mk[s0 + 1] &= m;
          ws.mk[s0 + 2] &= m;
        } else {
          Slot v = cls[0][s0 + (j == 0 ? K : j)];
          v.a &= m;
          v.b &= m;
          cls[0][s0 + (j == 0 ? K : j)] = v;
          cS += __popcll(v.a);
        }
      }
    }
    __syncwarp();
    // ---- word 0 = own result (pad slot) | what the left neighbour found in it (its overlap word) ----
    {
      Slot v = cls[0][s0 + K];
      if (lane > 0) {
        const uint64_t rs = ws.mk[s0 - (K + 1) + 1], re = ws.mk[s0 - (K + 1) + 2];
        v.a |= rs;
        v.b |= re;
        cS += __popcll(rs);
      }
      cls[0][s0] = v;
    }
    __syncwarp();
    unsigned far = 0u;
    if (__any_sync(FULL, replay)) {
      if (lane == 0) atomicAdd(&a.total[2], 1ull);  // diagnostics: chunks with a serial replay (cgx_debug_scratch)
#ifndef CGX_RBITS
#define CGX_RBITS 1
#endif
      // open segments: bit-parallel replay, one lane at a time (the staging buffer is its scratch)
      bool done_bits = false;
      unsigned todo = CGX_RBITS ? __ballot_sync(FULL, replay && !bad) : 0u;
      if (todo) {
        if (P_MODE == M_FINDALL) flush(sb, true);  // (the buffer's previous chunk has to be out)
        uint64_t* scratch = reinterpret_cast<uint64_t*>(&ws.st[CGX_PARK ? 0 : sb][0][0]);
        while (todo) {
          const int l = __ffs((int)todo) - 1;
          todo &= todo - 1u;
          if (lane == l && replay_bits(a, f, cb, cls[0], rp_from, rp_stop, one, scratch, &far)) {
            done_bits = true;
            atomicAdd(&a.total[3], 1ull);  // diagnostics: open segments replayed bit-parallel
          }
          __syncwarp();
        }
      }
      if (replay && !done_bits) {
        unsigned nb = 0;
        far = replay_cold(a, cb, cls[0], rp_from, rp_stop, nullptr, 0ull, &nb);
      }
      __syncwarp();
      // bits may have landed in other lanes' words: count again
      cS = 0u;
      for (int j = 0; j < K; j++) cS += __popcll(cls[0][s0 + j].a);
    }
#endif  // !CGX_TEDDY
    // ---- counts and ranks ----
    uint32_t x = cS;
    const uint32_t mine = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(FULL, x, d);
      if (lane >= d) x += y;
    }
    const uint32_t tot = __shfl_sync(FULL, x, 31);
    const unsigned nbits = tot;
    const unsigned nfar = __reduce_add_sync(FULL, far);
    const unsigned cnt = nbits + nfar;
    if (P_MODE == M_FINDALL) {
      const uint32_t rk = x - mine;      // starts before this lane's words
      post_count(tcur, cnt);             // everybody behind the gang can go on once it is complete
      flush(sb, true);                   // (the staging buffer's previous chunk is two tickets old)
      if (lane == 0) {
        ws.far_cnt[sb] = nfar;
        ws.bits_cnt[sb] = nbits;
      }
      if (far) {  // (one lane at most: the owner of the chunk's last segment)
        ws.far_from[sb] = rp_from;
        ws.far_stop[sb] = rp_stop;
      }
      const bool staged = !CGX_PARK && nbits <= (unsigned)CAP;
      if (staged) {
        extract_staged(ws, cls[0], sb, rk, 0u, nbits, lane);
      } else if (lane == 0) {
        // nothing staged: the bitmaps are turned into pairs when the offset is known — below, or
        // (CGX_PARK) when the slot array is needed again
        ws.bits_cnt[sb] = 0u;
      }
      __syncwarp();
      if (lane == 0) {
        Mail& m = cs.mail[warp][sb];
        m.chunk = cur;
        m.cnt = cnt;
        cgx_fence_block();
        m.state = 1;
      }
      if (!staged && !CGX_PARK) {
        // more matches than the staging buffer holds: wait for the offset here and store straight
        // from the bitmaps
        Mail& m = cs.mail[warp][sb];
        for (;;) {
          int st = 0;
          if (lane == 0) st = m.state;
          st = __shfl_sync(FULL, st, 0);
          if (st == 2) break;
          cgx_backoff();
        }
        cgx_fence_block();
        const unsigned long long excl = m.excl;
        report_total(cur, excl + cnt);
        // ends before this lane's words (the staged path finds ends per match and needs no rank)
        uint32_t xe = 0u;
        for (int j = 0; j < K; j++) xe += __popcll(cls[0][s0 + j].b);
        const uint32_t mine_e = xe;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(FULL, xe, d);
          if (lane >= d) xe += y;
        }
        extract_direct(a, cls[0], cur, excl, cnt, rk, xe - mine_e, lane);
        if (lane == 0 && nfar)
          replay_cold(a, cb, nullptr, ws.far_from[sb], ws.far_stop[sb], a.out, excl + nbits, nullptr);
        __syncwarp();
        if (lane == 0) m.state = 0;
        __syncwarp();
      }
      sb ^= 1;
    } else if (cnt && lane == 0) {
      atomicAdd(a.total, (unsigned long long)cnt);
      a.total[1] = 1ull;
    }
    tcur = tnxt;
    cur = nxt;
  }
  if (P_MODE == M_FINDALL) {
    flush(sb, true);
    flush(sb ^ 1, true);
  }
}
#ifndef CGX_JIT
}  // namespace
#endif

#if !defined(CGX_JIT) || defined(CGX_CPU_SIM)
int64_t scan_flat_chunks(int64_t n) {
  if (n <= 0) return 0;
  return (n + CHUNKB - 1) / CHUNKB;
}

size_t scan_flat_smem_bytes() { return sizeof(CtaSmem); }
int scan_flat_threads() { return FW_THREADS; }
int scan_flat_warps() { return FW_WARPS; }

#ifdef CGX_CPU_SIM
void sim_launch_scan_flat(const ScanArgs& a, unsigned grid) {
#ifdef CGX_JIT
  sim::launch<ScanArgs>(cgx_flat_jit, grid, FW_THREADS, sizeof(CtaSmem), a);
#else
  sim::launch<ScanArgs>(scan_flat_kernel, grid, FW_THREADS, sizeof(CtaSmem), a);
#endif
}
#else
// Launches the scan on `stream`.  See capi.cu for what must be zero before the launch.
cudaError_t launch_scan_flat(const ScanArgs& a, int sm_count, cudaStream_t stream, int* grid_out) {
  if (a.nchunks == 0) return cudaSuccess;
  const size_t smem = sizeof(CtaSmem);
  static bool configured = false;
  static int per_sm = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(scan_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_flat_kernel, FW_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    configured = true;
  }
  int64_t grid = (int64_t)sm_count * per_sm;
  const int64_t need = (a.nchunks + FW_WARPS - 1) / FW_WARPS;
  if (grid > need) grid = need;
  if (grid_out) *grid_out = (int)grid;
  scan_flat_kernel<<<(unsigned)grid, FW_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}
#endif

#endif  // host side

}  // namespace cgx
