// scan_flat.cu — the bitstream engine: sm_100a kernel for flat deterministic patterns
// (`\d+\.\d+\.\d+\.\d+`, `\w+@\w+\.\w+`, `[a-z]+=\d+` ...), the north-star path.
//
// Replaces, for a whole corpus at once:
//   reference meta/findall.go:176-290        findAllIndicesLoop (pos = end chaining)
//   reference meta/find_indices.go:1050-1088 DigitPrefilter loop (candidate -> SearchAtAnchored)
//   reference simd/memchr_digit_amd64.s:26   memchrDigitAVX2
//   reference dfa/lazy/lazy.go:219-324       SearchAtAnchored (per-byte class + table walk)
//
// Where scan_dfa.cu finds candidate starts bit-parallel and then walks the DFA one candidate per
// lane, this kernel never walks: match STARTS come from a right-to-left marker pass over per-class
// position bitmaps, match ENDS from a left-to-right pass over the same bitmaps (forced greedy ==
// leftmost-first because the host proved the pattern deterministic, host/engine.cpp
// DecideBitstream).  Every scanning warp is autonomous — no CTA barrier anywhere:
//
//   * a warp draws 23.8 KB chunks (12 tiles) from a ticket counter; per iteration it takes one TMA
//     bulk copy of one 2 KB tile into its own double-buffered window (CGX_TILES=2: two overlapping
//     tiles evaluated jointly);
//   * lane l owns a 64-byte piece of the tile: 4 x LDS.128, SWAR range tests (2 LOP3 + 1 IMAD per
//     word and class), dp4a bit packing -> one 64-bit word per class;
//   * tiles overlap by one piece (64 B).  A byte that belongs to no class of the pattern can never
//     be inside a match ("sync byte"), so a tile owns the starts from the first sync byte of its
//     first piece up to the first sync byte of its last piece: chains of overlapping candidates
//     (`pos = end` in the reference loop) never cross tiles;
//   * the passes add 2048-bit numbers; a carry crosses a lane boundary by one shuffle (a tile with
//     a piece made of class bytes only takes the exact two-ballot carry resolution instead);
//   * starts and ends are checked to alternate (per word, from the rank scan: word_misordered).
//     If they do not (candidates that overlap, `1.2.3.4.5`), or a span has no sync byte before the
//     window ends, one lane replays the reference loop over global memory for that region (exact,
//     rare);
//   * matches are staged per chunk in shared memory (u16 offsets) and the chunk's count is
//     published at once.  One resolver warp per CTA performs the decoupled look-back for the
//     CTA's chunks, oldest first, and hands each offset back; the scanning warp stores its staged
//     pairs as int64 in global match order when it next needs that staging buffer.
// Every corpus byte crosses HBM once (+3 % tile overlap served by L2); output is 16 B per match.
// The kernel is bound by the integer ALU pipe (ncu: alu ~78 %, issue ~81 %), so its inner parts are
// written for ALU-pipe instruction count; see "pipe-aware primitives" below and DESIGN.md §5.0.
#include "scan_common.cuh"
#include "scan_params.h"

namespace cgx {

namespace {

#ifndef CGX_TILES
#define CGX_TILES 2
#endif
// CGX_PAIR=1 (experiment, one-tile build only): one bulk copy brings TWO tiles (4032 B) into the
// warp's single 4 KB window; the tiles are evaluated one after the other and the window is refilled
// as soon as the second one has been classified.  Halves the per-tile cost of TMA issue and
// mbarrier wait (~85 of ~620 instructions per tile); the copy's latency is hidden by the second
// tile's passes and by the other warps instead of by a second buffer.  Same chunk geometry as the
// default, so it can be A/B-ed with CGX_JIT_DEFS="-DCGX_PAIR=1" (tools/micro/exp.sh).
#ifndef CGX_PAIR
#define CGX_PAIR 0
#endif
#ifndef CGX_IT_UNROLL
#define CGX_IT_UNROLL 1  // 2: both window buffers in one loop body (constant buffer index, twice the hot code)
#endif
constexpr int IT_UNROLL = CGX_IT_UNROLL;
constexpr uint32_t FULL = 0xffffffffu;
// Launch shape (overridable for experiments: CGX_JIT_DEFS="-DCGX_WARPS=31 -DCGX_CTAS=1").
// Few large CTAs: every CTA spends one warp on the resolver, and the scan is bound by how many
// scanning warps share an SM's integer pipe (measured, 16 GiB IP scan: 4 x (7+1) warps 2064 GB/s,
// 2 x (15+1) warps 2464 GB/s at the same 64 registers per thread).
#ifndef CGX_WARPS
#if CGX_TILES == 1
#define CGX_WARPS 15
#else
#define CGX_WARPS 11
#endif
#endif
#ifndef CGX_CTAS
#define CGX_CTAS 2
#endif
constexpr int FW_WARPS = CGX_WARPS;              // scanning warps per CTA
constexpr int FW_THREADS = (FW_WARPS + 1) * 32;  // + one resolver warp (look-back and ordered output)
constexpr int FW_CTAS = CGX_CTAS;                // resident CTAs per SM the kernel is built for
static_assert(FW_WARPS <= 31, "one CTA holds at most 31 scanning warps and the resolver");
constexpr int TILE = 2048;              // window bytes of one tile (64 per lane)
constexpr int STRIDE = 1984;            // bytes between tile origins (31 pieces)
constexpr int TPC = 12;                 // tiles per chunk (per-chunk costs — ticket, count, look-back, hand-over — are paid once per 23.8 KB)
// tiles evaluated jointly per iteration: 2 gives every warp two independent dependency chains
// (fewer, fatter warps), 1 halves the hot loop's code and registers (more warps)
constexpr int NT = CGX_TILES;
constexpr int ITERS = TPC / NT;         // iterations per chunk
constexpr int CHUNKB = TPC * STRIDE;    // 15872 bytes owned per chunk
constexpr int SUPER = (NT - 1) * STRIDE + TILE;  // bytes loaded per iteration (4032 / 2048)
constexpr int CAP = 384;                // staged matches per chunk

// The pattern-dependent parts exist twice: as an interpreter over ScanArgs::flat (this translation
// unit as nvcc builds it, and the CPU emulator build), and — when the host JIT-compiles this file
// with NVRTC for one pattern (csrc/jit.cpp, -DCGX_JIT + a generated cgx_jit_prog.h) — as
// straight-line code: no program loads, no dispatch, class constants as immediates.
#ifdef CGX_JIT
#define P_NCLASSES CGX_JIT_NCLASSES
#define P_RUNSTART CGX_JIT_RUNSTART
#define P_MIDRUN CGX_JIT_MIDRUN
#else
#define P_NCLASSES f.nclasses
#define P_RUNSTART f.bs_runstart
#define P_MIDRUN f.bs_midrun_check
#endif

struct WarpSmem {
  alignas(128) uint8_t win[2][SUPER];
  uint16_t stS[2][CAP];
  uint16_t stE[2][CAP];
  uint64_t mbar[2];
  // what only the cold paths need of the current chunk (kept out of registers: the hot loop is at
  // its register limit and was recomputing loop invariants every iteration)
  int64_t cb;               // global position of the chunk's first byte
  unsigned long long goff;  // direct: global index of the chunk's first match
  const uint8_t* csrc;      // the chunk's first byte in global memory
  int whole;                // every window of the chunk lies inside the input
  int chunk0;               // the chunk starts the haystack
};
// a scanned chunk handed from a scanning warp to the CTA's resolver warp (one slot per staging buffer)
struct Mail {
  volatile int state;  // 0 = slot and staging buffer free, 1 = chunk waits for its offset, 2 = offset known
  unsigned cnt;
  int64_t chunk;
  unsigned long long excl;  // state 2: global index of the chunk's first match
};
struct CtaSmem {
  WarpSmem w[FW_WARPS];
  Mail mail[FW_WARPS][2];
  volatile int done[FW_WARPS];
};

__device__ __forceinline__ uint64_t mk64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// ---- pipe-aware primitives ----------------------------------------------------------------------
// The kernel is bound by the integer ALU pipe (LOP3/SHF/ISETP/IADD3: one warp instruction per two
// cycles per sub-partition; ncu: alu 73 %, fma 15 %), so the per-byte work is written to need as few
// ALU-pipe instructions as possible:
//  * a LOP3 takes one immediate at most.  With both constants of `(w ^ k) & m` as immediates the
//    compiler emits two LOP3s; as an explicit lop3.b32 one of them is kept in a register: one LOP3.
//  * `x + k` as x * one + k with a multiplier the compiler cannot see through (ScanArgs-derived 1)
//    is an IMAD: same result, issued to the FMA pipe, which is otherwise idle.
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

// one step of an inclusive warp scan: x += (value of lane - d), lanes below d keep x.  The shuffle's
// own "source lane in range" predicate guards the add: no lane compare, no select.
__device__ __forceinline__ uint32_t scan_step(uint32_t x, int d, int lane) {
#if defined(CGX_CPU_SIM) || !defined(__CUDA_ARCH__)
  const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
  return lane >= d ? x + y : x;
#else
  uint32_t r;
  asm volatile(
      "{\n"
      ".reg .u32 t;\n"
      ".reg .pred p;\n"
      "shfl.sync.up.b32 t|p, %1, %2, 0, 0xffffffff;\n"
      "mov.u32 %0, %1;\n"
      "@p add.u32 %0, %1, t;\n"
      "}\n"
      : "=&r"(r)
      : "r"(x), "r"(d));
  return r;
#endif
}

// ---- 2048-bit vectors across the warp: lane l holds word l -------------------------------------
// markers move one position towards higher bit indices; `in` enters bit 0 of lane 0
__device__ __forceinline__ uint64_t shl1(uint64_t m, int lane, uint32_t in) {
  const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
  uint32_t dn = __shfl_up_sync(FULL, hi, 1);
  if (lane == 0) dn = in;
  return mk64(__funnelshift_l(lo, hi, 1), __funnelshift_l(dn, lo, 1));
}
// 2048-bit addition s + cc.  EXACT resolves generate AND propagate words (two ballots).  The fast
// form assumes that no word propagates: with s a subset of cc (a marker can only stand on a byte
// of the class) a word's sum is all ones only if the class word itself is all ones — 64 class
// bytes in one lane's piece — which process_tiles tests once per tile; then the carry into a word
// is just the carry out of its neighbour: no ballot at all, 8 ALU-pipe instructions fewer.
// The same shift for the marker passes, where what enters bit 0 of lane 0 does not matter (lane 0
// keeps the top bit of its own word): in the right-to-left pass that bit stands for the byte past
// the window, and nothing can travel from there to an owned start without crossing the sync byte
// that ends the owned range; in the left-to-right pass a stray marker at the window's first byte
// dies at the first sync byte at the latest, before the owned range begins, and the ends are masked
// with the ownership mask (process_tiles).  Saves one SEL per step on the ALU pipe.
__device__ __forceinline__ uint64_t shl1x(uint64_t m) {
  const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
  const uint32_t dn = __shfl_up_sync(FULL, hi, 1);
  return mk64(__funnelshift_l(lo, hi, 1), __funnelshift_l(dn, lo, 1));
}
// 64-bit a + b and its carry out as a 0/1 register: the carry comes straight from the adder's flag
// (IADD3 / IADD3.X / IADD3.X) instead of a 64-bit compare (two ISETP).
__device__ __forceinline__ uint64_t add_carry(uint64_t x, uint64_t y, uint32_t& carry) {
#if defined(CGX_CPU_SIM) || !defined(__CUDA_ARCH__)
  const uint64_t sum = x + y;
  carry = sum < x ? 1u : 0u;
  return sum;
#else
  uint32_t lo, hi;
  asm("add.cc.u32 %0, %3, %5;\n\taddc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, 0, 0;"
      : "=r"(lo), "=r"(hi), "=r"(carry)
      : "r"((uint32_t)x), "r"((uint32_t)(x >> 32)), "r"((uint32_t)y), "r"((uint32_t)(y >> 32)));
  return mk64(hi, lo);
#endif
}
__device__ __forceinline__ uint64_t shl2x(uint64_t m) {  // shl1x twice
  const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
  const uint32_t dn = __shfl_up_sync(FULL, hi, 1);
  return mk64(__funnelshift_l(lo, hi, 2), __funnelshift_l(dn, lo, 2));
}
template <bool EXACT>
__device__ __forceinline__ uint64_t add2048(uint64_t s, uint64_t cc, int lane, uint32_t prevbit) {
  if (!EXACT) {
    // the carry into a word is the carry out of the word below: one shuffle.  Lane 0 receives its
    // own carry — a stray marker at the first bit, harmless for the reasons given at shl1x.
    uint32_t g;
    const uint64_t sum = add_carry(s, cc, g);
    return sum + __shfl_up_sync(FULL, g, 1);
  }
  const uint64_t sum = s + cc;
  const uint32_t G = __ballot_sync(FULL, sum < s);
  const uint32_t P = __ballot_sync(FULL, sum == ~0ull);
  const uint32_t A = G | P;
  const uint32_t carries = A ^ G ^ (A + G);  // bit l = carry into word l
  return sum + ((carries >> lane) & 1u);
}
// orientation flip: reversed (lane l = piece 31-l, bit 63-b = byte b) <-> forward (lane l = piece l,
// bit b = byte b)
__device__ __forceinline__ uint64_t flip(uint64_t x) {
  const uint32_t lo = __brev((uint32_t)(x >> 32)), hi = __brev((uint32_t)x);
  return mk64(__shfl_xor_sync(FULL, hi, 31), __shfl_xor_sync(FULL, lo, 31));
}

// ---- classification ----------------------------------------------------------------------------
#ifdef CGX_JIT
#include "cgx_jit_prog.h"  // generated per pattern (host/engine.cpp JitHeader), compiled by csrc/jit.cpp
#endif
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

template <int C>
__device__ __forceinline__ uint64_t class_rev64(const FlatDev& f, const uint32_t (&w)[16], uint32_t one) {
  uint32_t fl[16];
#ifdef CGX_JIT
#pragma unroll
  for (int k = 0; k < 16; k++) fl[k] = cgx_jit_flags<C>(w[k], one);  // generated: ranges as constants
#else
  const int nr = f.cls_nranges[C];
  if (nr == 1 && f.cls_mode[C][0] == 0) {  // one XOR-alignable range: 2 ALU + 1 FMA op per word
    const uint32_t k1 = f.cls_k1[C][0], k2 = f.cls_k2[C][0];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const uint32_t z = mad_fma(xor_and(w[k], k1, 0x7F7F7F7Fu), one, k2);  // bit7 set <=> (x^lo)&0x7f > width
      fl[k] = nor_and(z, w[k], 0x80808080u);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; k++) fl[k] = 0;
    for (int r = 0; r < nr; r++) {
      const uint32_t k1 = f.cls_k1[C][r], k2 = f.cls_k2[C][r];
      if (f.cls_mode[C][r] == 0) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const uint32_t z = mad_fma(xor_and(w[k], k1, 0x7F7F7F7Fu), one, k2);
          fl[k] |= nor_and(z, w[k], 0x80808080u);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; k++) fl[k] |= swar_in_range(w[k], k1, k2);
      }
    }
  }
#endif
  return mk64(pack_rev(fl, one), pack_rev(fl + 8, one));  // bytes 0..31 in the high word
}

// Class bitmaps (reversed orientation) of the 64-byte piece at `p`: 4 x LDS.128.  Pieces are 64 B
// apart, so with a straight quarter order one LDS.128 of the warp touches 8 of the 32 banks (4x the
// wavefronts).  CGX_ROT=1 reads the quarters in the order (j + rot) & 3, which covers all banks
// evenly, and undoes the resulting rotation of each bitmap with two byte permutes per class.  That
// trades ~10 ALU-pipe instructions per tile for shared-memory wavefronts; the kernel is bound by
// the ALU pipe and the LSU pipe is at 17 %, so the straight order is the default (measured:
// 2584 -> 2689 GB/s).
#ifndef CGX_ROT
#define CGX_ROT 0
#endif
__device__ __forceinline__ void classify_piece(const FlatDev& f, const uint8_t* p, int lane, uint32_t one,
                                               uint64_t (&cm)[4]) {
  uint32_t w[16];
#if CGX_ROT
  const int rot = ((31 - lane) >> 1) & 3;
#else
  const int rot = 0;
#endif
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint4 v = *reinterpret_cast<const uint4*>(p + (((j + rot) & 3) << 4));
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
  // rotate right by 16*rot bits = 2*rot bytes: selectors are 16-bit windows of one constant
  const uint32_t sel_lo = (uint32_t)(0x1076765454323210ull >> (16 * rot)) & 0xFFFFu;
  const uint32_t sel_hi = (uint32_t)(0x1076765454323210ull >> (16 * ((rot + 2) & 3))) & 0xFFFFu;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    if (c < P_NCLASSES) {
      const uint32_t lo = (uint32_t)cm[c], hi = (uint32_t)(cm[c] >> 32);
      cm[c] = mk64(__byte_perm(lo, hi, sel_hi), __byte_perm(lo, hi, sel_lo));
    }
  }
#endif
}

// ---- marker passes -----------------------------------------------------------------------------
// Right to left (reversed orientation): M = positions from which items k..end can match.
// (tile B's operands are touched only when two tiles are evaluated jointly)
template <int C, bool EXACT>
__device__ __forceinline__ void rev_step(uint32_t kind, const uint64_t (&ca)[4], const uint64_t (&cb)[4],
                                         uint64_t& Ma, uint64_t& Mb, int lane, uint32_t prevbit) {
  const uint64_t Ca = ca[C], Cb = cb[C];
  // unknown territory past the window is assumed to allow a match (bit entering lane 0 is 1)
  const uint64_t ua = shl1x(Ma) & Ca;
  uint64_t ub = 0;
  if (NT == 2) ub = shl1x(Mb) & Cb;
  if (kind == 0) {
    Ma = ua;
    Mb = ub;
  } else if (kind == 3) {
    Ma |= ua;
    Mb |= ub;
  } else {
    // extend through the run towards lower addresses; a second marker inside one run survives the
    // carry of the first as a 1 in the sum, so the markers themselves are OR-ed back
    const uint64_t pa = (~add2048<EXACT>(ua, Ca, lane, prevbit) & Ca) | ua;
    Ma = kind == 1 ? pa : (Ma | pa);
    if (NT == 2) {
      const uint64_t pb = (~add2048<EXACT>(ub, Cb, lane, prevbit) & Cb) | ub;
      Mb = kind == 1 ? pb : (Mb | pb);
    }
  }
}
// `C+ a` seen from the right: one byte of class A, then a run of class C.  q = shl1x(A) & C marks the
// run bytes that directly follow (in marker direction) a byte of A, so shl1x(shl1x(M) & A) & C ==
// shl2x(M) & q: the two steps cost one shift.
template <int C, bool EXACT>
__device__ __forceinline__ void rev_fused(const uint64_t (&ca)[4], const uint64_t (&cb)[4], uint64_t qa, uint64_t qb,
                                          uint64_t& Ma, uint64_t& Mb, int lane, uint32_t prevbit) {
  const uint64_t ua = shl2x(Ma) & qa;
  Ma = (~add2048<EXACT>(ua, ca[C], lane, prevbit) & ca[C]) | ua;
  if (NT == 2) {
    const uint64_t ub = shl2x(Mb) & qb;
    Mb = (~add2048<EXACT>(ub, cb[C], lane, prevbit) & cb[C]) | ub;
  }
}
// Left to right (forward orientation): T = positions a marker stands at before item k; the forced
// greedy choice (take the whole run / take the optional byte whenever it is there).
template <int C, bool EXACT>
__device__ __forceinline__ void fwd_step(uint32_t kind, const uint64_t (&ca)[4], const uint64_t (&cb)[4],
                                         uint64_t& Ta, uint64_t& Tb, int lane, uint32_t prevbit) {
  const uint64_t Ca = ca[C], Cb = cb[C];
  const uint64_t ia = Ta & Ca, ib = Tb & Cb;  // markers that can take a byte
  if (kind == 0) {
    Ta = shl1x(ia);
    if (NT == 2) Tb = shl1x(ib);
  } else if (kind == 3) {
    Ta = (Ta & ~Ca) | shl1x(ia);
    if (NT == 2) Tb = (Tb & ~Cb) | shl1x(ib);
  } else {
    // a marker inside a run of ones carries out to the first zero after the run
    const uint64_t ea = add2048<EXACT>(ia, Ca, lane, prevbit) & ~Ca;
    Ta = kind == 1 ? ea : ((Ta & ~Ca) | ea);
    if (NT == 2) {
      const uint64_t eb = add2048<EXACT>(ib, Cb, lane, prevbit) & ~Cb;
      Tb = kind == 1 ? eb : ((Tb & ~Cb) | eb);
    }
  }
}

#define CGX_CLASS_SWITCH(cls, CALL)  \
  switch (cls) {                     \
    case 0: { constexpr int C = 0; CALL; } break; \
    case 1: { constexpr int C = 1; CALL; } break; \
    case 2: { constexpr int C = 2; CALL; } break; \
    default: { constexpr int C = 3; CALL; } break; \
  }

// The two passes of one iteration.  EXACT: see add2048.
template <bool EXACT>
__device__ __forceinline__ void rev_pass(const FlatDev& f, const uint64_t (&ca)[4], const uint64_t (&cb)[4],
                                         uint64_t& Ma, uint64_t& Mb, int lane, uint32_t prevbit) {
#ifdef CGX_JIT
  Ma = ca[CGX_JIT_REV_INIT];
  Mb = cb[CGX_JIT_REV_INIT];
#define CGX_QDEF(A, B)                                 \
  const uint64_t qa##A##_##B = shl1x(ca[A]) & ca[B];   \
  uint64_t qb##A##_##B = 0ull;                         \
  if (NT == 2) qb##A##_##B = shl1x(cb[A]) & cb[B];
  CGX_JIT_REV_FUSED(CGX_QDEF)
#define CGX_STEP(kind, cls) rev_step<cls, EXACT>(kind, ca, cb, Ma, Mb, lane, prevbit);
#define CGX_FUSE(A, B) rev_fused<B, EXACT>(ca, cb, qa##A##_##B, qb##A##_##B, Ma, Mb, lane, prevbit);
  CGX_JIT_REV_PASS(CGX_STEP, CGX_FUSE)
#undef CGX_STEP
#undef CGX_FUSE
#undef CGX_QDEF
#else
  const int init_cls = f.rev_init_class;
  Ma = init_cls == 0 ? ca[0] : init_cls == 1 ? ca[1] : init_cls == 2 ? ca[2] : ca[3];
  Mb = init_cls == 0 ? cb[0] : init_cls == 1 ? cb[1] : init_cls == 2 ? cb[2] : cb[3];
  const int rev_nops = f.rev_nops;
  for (int k = 0; k < rev_nops; k++) {
    const uint32_t op = f.rev_ops[k];
    const uint32_t kind = op & 3u;
    CGX_CLASS_SWITCH(op >> 2, (rev_step<C, EXACT>(kind, ca, cb, Ma, Mb, lane, prevbit)));
  }
#endif
}
template <bool EXACT>
__device__ __forceinline__ void fwd_pass(const FlatDev& f, const uint64_t (&ca)[4], const uint64_t (&cb)[4],
                                         uint64_t& Ta, uint64_t& Tb, int lane, uint32_t prevbit) {
#ifdef CGX_JIT
#define CGX_STEP(kind, cls) fwd_step<cls, EXACT>(kind, ca, cb, Ta, Tb, lane, prevbit);
  CGX_JIT_FWD_PASS(CGX_STEP)
#undef CGX_STEP
#else
  const int fwd_nops = f.fwd_nops;
  for (int k = 0; k < fwd_nops; k++) {
    const uint32_t op = f.fwd_ops[k];
    const uint32_t kind = op & 3u;
    CGX_CLASS_SWITCH(op >> 2, (fwd_step<C, EXACT>(kind, ca, cb, Ta, Tb, lane, prevbit)));
  }
#endif
}

// Do the starts S and ends E of one tile alternate (start, end, start, end ...; an end may share
// its position with the next start)?  With in = (starts below this lane's word) - (ends below it)
// from the rank scan — 1 when a match is open at the word's first bit — the word is consistent iff
// D = E - S - in, the span mask, has its edges exactly at S ^ E: D ^ (D << 1 | in) == S ^ E, and
// in is 0 or 1.  Together with equal totals this is exact (tests/test_sim_flat.py checks the
// criterion exhaustively on short words); it replaces a 6-step prefix parity.
__device__ __forceinline__ bool word_misordered(uint64_t S, uint64_t E, uint32_t in) {
  const uint64_t D = E - S - in;
  return in > 1u || (D ^ ((D << 1) | in)) != (S ^ E);
}

// ---- output ------------------------------------------------------------------------------------
struct Emit {
  const ScanArgs* ap;
  WarpSmem* w;
  int sb;                   // staging buffer in use
  bool direct;              // store straight to global memory (chunk redone after a staging overflow)
  bool far = false;         // a staged offset did not fit 16 bits

  __device__ __forceinline__ int64_t cb() const { return w->cb; }
  // rel = position relative to the chunk
  __device__ __forceinline__ void put(unsigned idx, int64_t rel, bool is_end) {
    const ScanArgs& a = *ap;
    if (a.mode != M_FINDALL) return;
    if (direct) {
      const unsigned long long gi = w->goff + idx;
      if ((int64_t)gi < a.cap) a.out[2 * gi + (is_end ? 1 : 0)] = w->cb + a.base + rel;
    } else {
      if (rel > 0xFFFF) far = true;
      else if (idx < (unsigned)CAP) (is_end ? w->stE[sb] : w->stS[sb])[idx] = (uint16_t)rel;
    }
  }
  // fast path of the fast path: staged FindAll output, offsets known to fit 16 bits
  __device__ __forceinline__ unsigned stage_bits32(uint16_t* st, uint32_t bits, unsigned idx, int rel0) {
    while (bits) {
      const int b = __ffs((int)bits) - 1;
      bits &= bits - 1;
      if (idx < (unsigned)CAP) st[idx] = (uint16_t)(rel0 + b);
      idx++;
    }
    return idx;
  }
  __device__ __forceinline__ void put_bits(uint64_t bits, unsigned idx, int rel0, bool is_end) {
    if (ap->mode != M_FINDALL) return;
    if (!direct) {
      uint16_t* st = is_end ? w->stE[sb] : w->stS[sb];
      idx = stage_bits32(st, (uint32_t)bits, idx, rel0);
      stage_bits32(st, (uint32_t)(bits >> 32), idx, rel0 + 32);
      return;
    }
    while (bits) {
      const int b = __ffsll((long long)bits) - 1;
      bits &= bits - 1;
      put(idx++, rel0 + b, is_end);
    }
  }
};

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

// One lane replays the reference loop (meta/findall.go:176-290 over meta/find_indices.go:1050-1088)
// from global position `from` until the candidate search meets a sync byte at or after `stop_min`
// (or the end of input).  Returns the number of matches emitted from index idx0 on.  Cold path.
// `em` travels by value and the result comes back packed (count | far << 31): a reference would
// pin the caller's emitter and counters in local memory for the hot path too.
__device__ __noinline__ unsigned serial_region_cold(const ScanArgs& a, Emit em, int64_t from, int64_t stop_min,
                                                    unsigned idx0, int lane) {
  unsigned added = 0;
  bool far = false;
  if (lane == 0) {
    atomicAdd(&a.total[2], 1ull);  // diagnostics: serial replays (cgx_debug_scratch)
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
        em.put(idx0 + added, d - em.cb(), false);
        em.put(idx0 + added, e - em.cb(), true);
        added++;
        pos = e > d ? e : d + 1;
      } else {
        pos = d + 1;
        // a failed run start fails for the whole run (digitRunSkipSafe, meta/strategy.go:525-560)
        if (a.flat.bs_runstart)
          while (pos < n && in_filter(a, __ldg(a.h + pos))) pos++;
      }
    }
    far = em.far;
  }
  return __shfl_sync(FULL, added | (far ? 0x80000000u : 0u), 0);
}
__device__ __forceinline__ unsigned serial_region(const ScanArgs& a, Emit& em, int64_t from, int64_t stop_min,
                                                  unsigned idx0, int lane) {
  const unsigned r = serial_region_cold(a, em, from, stop_min, idx0, lane);
  if (r >> 31) em.far = true;
  return r & 0x7FFFFFFFu;
}

// ---- one iteration: two overlapping tiles ------------------------------------------------------
struct TileOut {
  uint64_t S, E;      // forward orientation, owned starts and their ends
  uint64_t nz;        // this lane's sync bytes (bytes that belong to no class)
  uint32_t has;       // lanes whose piece holds a sync byte
  bool first;         // the tile starts the haystack
  bool owned;         // the tile owns any starts at all
  bool open;          // no sync byte in the last piece: the span after own_lim is finished serially
};

__device__ __forceinline__ int own_start(const TileOut& t, int lane);
__device__ __forceinline__ int own_open_lim(const TileOut& t, int lane);
// Ownership of one tile from U (union of the classes, forward orientation): owned starts run from
// the byte after the window's first sync byte (the start of the haystack for its first tile) up to
// the first sync byte of the last piece.  The mask costs one ballot and a few word operations;
// the positions themselves (own_start / own_open_lim) are only needed on the rare serial paths.
__device__ __forceinline__ uint64_t ownership(uint64_t U, bool first_tile, int lane, TileOut& t) {
  const uint64_t nz = ~U;
  const uint32_t has = __ballot_sync(FULL, nz != 0ull);
  t.nz = nz;
  t.has = has;
  t.first = first_tile;
  t.owned = first_tile || has != 0u;
  t.open = (has >> 31) == 0u;
  // nz == 0 gives upto == all ones: a lane below the first one with a sync byte owns nothing
  // without being told apart from it
  const uint64_t upto = nz ^ (nz - 1ull);  // bits up to and including this lane's first sync byte
  const uint32_t lt = (1u << lane) - 1u;
  uint64_t m = (first_tile || (has & lt)) ? ~0ull : ~upto;
  if (!t.open) {
    if (lane == 31) m &= upto;
  } else {
    // starts before (and at) the window's LAST sync byte; the tail span is replayed serially
    const int hl = has ? 31 - __clz((int)has) : -1;
    const int top = nz ? 63 - __clzll((long long)nz) : 0;
    const uint64_t below = top == 63 ? ~0ull : ((2ull << top) - 1ull);
    m &= lane > hl ? 0ull : (lane == hl ? below : ~0ull);
  }
  return m;
}
// first owned position (tile-relative); only valid when t.owned
__device__ __forceinline__ int own_start(const TileOut& t, int lane) {
  if (t.first) return 0;
  const int fl = __ffs((int)t.has) - 1;
  return 64 * fl + __shfl_sync(FULL, __ffsll((long long)t.nz) - 1, fl) + 1;
}
// open tiles: where the serially replayed tail span starts
__device__ __forceinline__ int own_open_lim(const TileOut& t, int lane) {
  const int a = own_start(t, lane);
  if (!t.has) return a;
  const int hl = 31 - __clz((int)t.has);
  const int z1 = 64 * hl + __shfl_sync(FULL, 63 - __clzll((long long)t.nz), hl) + 1;
  return a > z1 ? a : z1;
}

// trel = position of the tile's first byte relative to the chunk
__device__ __forceinline__ void finish_tile(const ScanArgs& a, Emit& em, const TileOut& t, int trel,
                                            unsigned totS, unsigned totE, unsigned exS, unsigned exE, bool bad,
                                            unsigned& cnt, int lane) {
  if (!t.owned) return;
  if (bad || totS != totE) {
    const int64_t tile_g = em.cb() + trel;
    cnt += serial_region(a, em, tile_g + own_start(t, lane), tile_g + STRIDE, cnt, lane);
    return;
  }
  if (totS) {
    const int rel0 = trel + 64 * lane;
    if (a.mode == M_FINDALL) {
      if (!em.direct) {
        // Rounds with a warp-uniform trip count (the largest number of starts or ends any lane
        // holds, usually 1 or 2): every round each lane stores its next start and its next end.
        // No divergent loop, hence no reconvergence bookkeeping.
        // A tile that does not fit the staging buffer any more stages nothing: the chunk has
        // overflowed and will be run again with direct stores (no per-store bound checks).
        if (cnt + totS <= (unsigned)CAP) {
          uint64_t sb = t.S, eb = t.E;
          uint16_t* ps_ = em.w->stS[em.sb] + cnt + exS;
          uint16_t* pe_ = em.w->stE[em.sb] + cnt + exE;
          const unsigned ps = __popcll(sb), pe = __popcll(eb);
          const unsigned rounds = __reduce_max_sync(FULL, ps > pe ? ps : pe);
          // straight-line rounds on 32-bit halves: the half that still has bits is picked by
          // select, the stores are predicated (no divergent branches, half the 64-bit arithmetic)
          uint32_t slo = (uint32_t)sb, shi = (uint32_t)(sb >> 32), elo = (uint32_t)eb, ehi = (uint32_t)(eb >> 32);
#pragma unroll 1
          for (unsigned j = 0; j < rounds; j++) {
            {
              const bool any = (slo | shi) != 0u, l = slo != 0u;
              const uint32_t w = l ? slo : shi;
              const int pos = rel0 + (l ? 0 : 32) + (__ffs((int)w) - 1);
              if (any) *ps_ = (uint16_t)pos;
              ps_ += any ? 1 : 0;
              const uint32_t w2 = w & (w - 1u);
              slo = l ? w2 : 0u;
              shi = l ? shi : w2;
            }
            {
              const bool any = (elo | ehi) != 0u, l = elo != 0u;
              const uint32_t w = l ? elo : ehi;
              const int pos = rel0 + (l ? 0 : 32) + (__ffs((int)w) - 1);
              if (any) *pe_ = (uint16_t)pos;
              pe_ += any ? 1 : 0;
              const uint32_t w2 = w & (w - 1u);
              elo = l ? w2 : 0u;
              ehi = l ? ehi : w2;
            }
          }
        }
      } else {
        em.put_bits(t.S, cnt + exS, rel0, false);
        em.put_bits(t.E, cnt + exE, rel0, true);
      }
    }
    cnt += totS;
  }
  if (t.open) cnt += serial_region(a, em, em.cb() + trel + own_open_lim(t, lane), em.cb() + trel + STRIDE, cnt, lane);
}

// bytes at or beyond the end of input belong to no class (last chunk only).  Inline on purpose: as
// a real call it would pin the class bitmaps in local memory (measured -12 %).
__device__ __forceinline__
void mask_tail(uint64_t (&ca)[4], uint64_t (&cb)[4], int64_t nv, int lane) {
  // reversed bit r of lane l <=> tile byte 2047 - (64 l + r); valid <=> byte < nv
  const int64_t ra = TILE - nv, rb = TILE - (nv - STRIDE);  // first valid reversed index
  const int64_t sa = ra - 64 * lane, sb = rb - 64 * lane;
  const uint64_t va = sa <= 0 ? ~0ull : (sa >= 64 ? 0ull : (~0ull << sa));
  const uint64_t vb = sb <= 0 ? ~0ull : (sb >= 64 ? 0ull : (~0ull << sb));
#pragma unroll
  for (int c = 0; c < 4; c++) {
    ca[c] &= va;
    cb[c] &= vb;
  }
}

// Processes the NT tiles whose windows start at `win` (position wrel in the chunk), win + STRIDE.
// nv = valid bytes from the start of tile A, clamped to SUPER; first = the window starts the haystack.
// after_classify() runs as soon as the window's bytes are in registers: the pair-load variant
// (CGX_PAIR) refills the window from there.
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
template <class Hook>
__device__ __forceinline__ void process_tiles(const ScanArgs& a, Emit& em, const uint8_t* win, int wrel, int nv,
                                              bool first, unsigned& cnt, int lane, Hook after_classify) {
  const FlatDev& f = a.flat;
  const int piece = 31 - lane;
  uint64_t ca[4], cb[4] = {0ull, 0ull, 0ull, 0ull};
  // a 1 the compiler cannot fold (a launch has at least one chunk): multiplier of mad_fma
  const uint32_t one = (uint32_t)(a.nchunks > 0);
  classify_piece(f, win + piece * 64, lane, one, ca);
  if (NT == 2) classify_piece(f, win + STRIDE + piece * 64, lane, one, cb);
  after_classify();
  // bytes at or beyond the end of input belong to no class
  if (nv < SUPER) mask_tail(ca, cb, nv, lane);

  // a class word of all ones (64 class bytes in one piece) needs the exact carry resolution
  uint32_t fullw = 0u;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    if (c < P_NCLASSES) {
      const uint32_t t = (uint32_t)ca[c] & (uint32_t)(ca[c] >> 32);
      fullw = t > fullw ? t : fullw;
      if (NT == 2) {
        const uint32_t u = (uint32_t)cb[c] & (uint32_t)(cb[c] >> 32);
        fullw = u > fullw ? u : fullw;
      }
    }
  }
  const bool exact = __any_sync(FULL, fullw == 0xFFFFFFFFu);
  const uint32_t prevbit = lane ? 1u << (lane - 1) : 0u;

  // ---- right to left: where can a match start ----
  uint64_t Ma, Mb;
  if (exact) rev_pass<true>(f, ca, cb, Ma, Mb, lane, prevbit);
  else rev_pass<false>(f, ca, cb, Ma, Mb, lane, prevbit);
  if (P_RUNSTART) {
    // only the first byte of a run of class 0 (pattern opens with C+); the byte before the window
    // counts as outside the class: position 0 is owned only when it is the start of the input
    uint32_t upa = __shfl_down_sync(FULL, (uint32_t)ca[0], 1) & 1u;
    if (lane == 31) upa = 0u;
    Ma &= ca[0] & ~((ca[0] >> 1) | ((uint64_t)upa << 63));
    if (NT == 2) {
      uint32_t upb = __shfl_down_sync(FULL, (uint32_t)cb[0], 1) & 1u;
      if (lane == 31) upb = 0u;
      Mb &= cb[0] & ~((cb[0] >> 1) | ((uint64_t)upb << 63));
    }
  }

  // ---- forward orientation from here on ----
#pragma unroll
  for (int c = 0; c < 4; c++) {
    if (c < P_NCLASSES) {
      ca[c] = flip(ca[c]);
      if (NT == 2) cb[c] = flip(cb[c]);
    }
  }
  TileOut ta, tb;
  const uint64_t oma = ownership(ca[0] | ca[1] | ca[2] | ca[3], first, lane, ta);
  uint64_t omb = 0ull;
  ta.S = flip(Ma) & oma;
  tb.S = tb.E = tb.nz = 0ull;
  tb.has = 0u;
  tb.first = tb.owned = tb.open = false;
  if (NT == 2) {
    omb = ownership(cb[0] | cb[1] | cb[2] | cb[3], false, lane, tb);
    tb.S = flip(Mb) & omb;
  }

  const uint32_t anyS = __ballot_sync(FULL, (ta.S | tb.S) != 0ull);
  if (a.mode == M_ISMATCH) {
    // starts inside the owned range are exact: a set bit is a match
    if (anyS) {
      if (lane == 0) a.total[1] = 1ull;
      cnt += 1;
      return;
    }
    // open tails may still hide a match
    const int64_t wg = em.cb() + wrel;
    if (ta.owned && ta.open) cnt += serial_region(a, em, wg + own_open_lim(ta, lane), wg + STRIDE, cnt, lane);
    if (NT == 2 && tb.owned && tb.open)
      cnt += serial_region(a, em, wg + STRIDE + own_open_lim(tb, lane), wg + 2 * STRIDE, cnt, lane);
    if (cnt && lane == 0) a.total[1] = 1ull;
    return;
  }

  unsigned exSa = 0, exEa = 0, exSb = 0, exEb = 0, totSa = 0, totEa = 0, totSb = 0, totEb = 0;
  bool bad = false, badb = false;
  ta.E = 0ull;
  if (anyS) {
    // ---- left to right: where do the matches end ----
    uint64_t Ta = ta.S, Tb = tb.S;
    if (exact) fwd_pass<true>(f, ca, cb, Ta, Tb, lane, prevbit);
    else fwd_pass<false>(f, ca, cb, Ta, Tb, lane, prevbit);
    // (a match from an owned start ends inside the owned range; the mask drops what a stray
    // marker of shl1x may have left before it)
    ta.E = Ta & oma;
    tb.E = Tb & omb;

    // ---- counts and ranks (tile A's matches precede tile B's) ----
    uint32_t xa = __popcll(ta.S) | (__popcll(ta.E) << 16), xb = __popcll(tb.S) | (__popcll(tb.E) << 16);
    const uint32_t va = xa, vb = xb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      xa = scan_step(xa, d, lane);
      if (NT == 2) xb = scan_step(xb, d, lane);
    }
    const uint32_t ta_tot = __shfl_sync(FULL, xa, 31);
    xa -= va;
    exSa = xa & 0xFFFFu; exEa = xa >> 16;
    totSa = ta_tot & 0xFFFFu; totEa = ta_tot >> 16;
    if (NT == 2) {
      const uint32_t tb_tot = __shfl_sync(FULL, xb, 31);
      xb -= vb;
      exSb = xb & 0xFFFFu; exEb = xb >> 16;
      totSb = tb_tot & 0xFFFFu; totEb = tb_tot >> 16;
    }

    // ---- starts and ends must alternate ----
    bool wa = word_misordered(ta.S, ta.E, exSa - exEa), wb = false;
    // (an end in the middle of a class-0 run: the reference resumes there, which is no run start)
    if (P_MIDRUN) wa |= (ta.E & ca[0] & shl1(ca[0], lane, 0u)) != 0ull;
    bad = __any_sync(FULL, wa);
    if (NT == 2) {
      wb = word_misordered(tb.S, tb.E, exSb - exEb);
      if (P_MIDRUN) wb |= (tb.E & cb[0] & shl1(cb[0], lane, 0u)) != 0ull;
      badb = __any_sync(FULL, wb);
    }
  }
  finish_tile(a, em, ta, wrel, totSa, totEa, exSa, exEa, bad, cnt, lane);
  if (NT == 2) finish_tile(a, em, tb, wrel + STRIDE, totSb, totEb, exSb, exEb, badb, cnt, lane);
}

// ---- two-level look-back ------------------------------------------------------------------------------
// Chunk c publishes its match count in status[c] (flag LB_AGG).  Chunks form groups of 32; the
// warp that finishes a group's last chunk publishes the group's sum in gstatus[g] (LB_AGG), and
// whoever learns the prefix at a group's start publishes it as gstatus[g-1] (LB_PREFIX, the
// inclusive prefix through group g-1).  The exclusive prefix of chunk c is therefore: counts of the
// earlier chunks of its own group (one 32-wide load) + a decoupled look-back over GROUP words
// (32 groups = 1024 chunks per step), instead of a walk over the ~3000 chunks that are in flight.
// Waiting is done by the CTA's resolver warp (and by a scanning warp only for a chunk that
// overflowed its staging buffer); scanning warps publish their count and move on.
__device__ unsigned long long lb_resolve(const ScanArgs& a, int64_t chunk, int lane) {
  const int64_t g = chunk >> 5;
  const int64_t idx1 = (g << 5) + lane;
  unsigned long long excl = 0, gpre = 0;
  bool have1 = false, have2 = g == 0;
  int64_t look = g - 1;
  for (;;) {
    // both levels are requested before either is looked at: one L2 round trip per attempt
    unsigned long long v1 = LB_AGG;     // lanes at or beyond the chunk contribute nothing
    unsigned long long v = LB_PREFIX;   // positions before group 0 act as a zero prefix
    if (!have1 && idx1 < chunk) v1 = ld_status(&a.status[idx1]);
    const int64_t idx2 = look - lane;
    if (!have2 && idx2 >= 0) v = ld_status(&a.gstatus[idx2]);
    if (!have1 && __all_sync(FULL, (v1 >> 62) != 0)) {
      excl = __reduce_add_sync(FULL, (unsigned)(v1 & 0xFFFFFFFFull));
      have1 = true;
    }
    if (!have2) {
      const uint32_t empty = __ballot_sync(FULL, (v >> 62) == 0);
      const uint32_t pm = __ballot_sync(FULL, (v >> 62) == 2);
      const int fe = empty ? __ffs((int)empty) - 1 : 32;
      const int fp = pm ? __ffs((int)pm) - 1 : 32;
      if (fp < fe) {
        // sums of the groups before the prefix, then the prefix itself
        const unsigned part = __reduce_add_sync(FULL, lane < fp ? (unsigned)(v & 0xFFFFFFFFull) : 0u);
        gpre += part + __shfl_sync(FULL, v & LB_VALUE, fp);
        have2 = true;
        // the inclusive prefix through the previous group, for everybody behind us
        if (lane == 0 && (look != g - 1 || fp != 0)) st_status(&a.gstatus[g - 1], LB_PREFIX | gpre);
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
// last arrival knows the group's total without re-reading anything: no fences.
__device__ __forceinline__ void publish_count(const ScanArgs& a, int64_t chunk, unsigned cnt, int lane) {
  if (lane == 0) {
    const int64_t g = chunk >> 5;
    const int64_t members = a.nchunks - (g << 5) < 32 ? a.nchunks - (g << 5) : 32;
    st_status(&a.status[chunk], LB_AGG | cnt);
    const unsigned long long old = atomicAdd(&a.gacc[g], (1ull << 40) | cnt);
    if ((int64_t)(old >> 40) == members - 1)
      st_status(&a.gstatus[g], LB_AGG | ((old & ((1ull << 40) - 1)) + cnt));
  }
}

// stores the staged matches of a resolved chunk in global match order
__device__ __forceinline__ void write_out(const ScanArgs& a, const WarpSmem& ws, int sb, int64_t chunk, unsigned cnt,
                                          unsigned long long excl, int lane) {
  if (lane == 0 && chunk == a.nchunks - 1) a.total[0] = excl + cnt;
  const int64_t b = chunk * (int64_t)CHUNKB + a.base;
  const uint16_t* ss = ws.stS[sb];
  const uint16_t* se = ws.stE[sb];
  if ((int64_t)(excl + cnt) <= a.cap) {  // the usual case: the whole chunk fits the output
    longlong2* o = reinterpret_cast<longlong2*>(a.out) + excl;
    for (unsigned i = lane; i < cnt; i += 32) o[i] = make_longlong2(b + ss[i], b + se[i]);
  } else {
    for (unsigned i = lane; i < cnt; i += 32) {
      const unsigned long long gi = excl + i;
      if ((int64_t)gi < a.cap)
        *reinterpret_cast<longlong2*>(a.out + 2 * gi) = make_longlong2(b + ss[i], b + se[i]);
    }
  }
}

// The resolver warp: takes the oldest handed-over chunk of its CTA (a younger one cannot resolve
// before it: both need every earlier count), waits for its offset and hands the offset back; the
// scanning warp stores its own staged matches when it next needs the buffer.  (Storing them here
// made this warp the bottleneck of a 15-warp CTA: 336 instructions per chunk, always busy.)
__device__ void resolver_warp(const ScanArgs& a, CtaSmem& cs, int lane) {
  for (;;) {
    // lane l looks at slots l and l + 32 (slot s = warp s >> 1, buffer s & 1)
    unsigned mine = 0xFFFFFFFFu;  // chunk numbers are 32-bit tickets
    int myslot = 0;
    bool all_done = true;
#pragma unroll
    for (int s = lane; s < 2 * FW_WARPS; s += 32) {
      const Mail& m = cs.mail[s >> 1][s & 1];
      all_done = all_done && cs.done[s >> 1] != 0;  // read BEFORE the slot: a warp hands over, then sets done
      cgx_fence_block();
      if (m.state == 1 && (unsigned)m.chunk < mine) {
        mine = (unsigned)m.chunk;
        myslot = s;
      }
    }
    const unsigned best = __reduce_min_sync(FULL, mine);
    if (best == 0xFFFFFFFFu) {
      if (__all_sync(FULL, all_done)) return;
      cgx_idle();  // nothing handed over: a chunk takes tens of microseconds to scan
      continue;
    }
    const int slot = __shfl_sync(FULL, myslot, __ffs((int)__ballot_sync(FULL, mine == best)) - 1);
    const unsigned long long excl = lb_resolve(a, (int64_t)best, lane);
    if (lane == 0) {
      Mail& m = cs.mail[slot >> 1][slot & 1];
      m.excl = excl;
      cgx_fence_block();
      m.state = 2;
    }
  }
}

}  // namespace

#ifdef CGX_JIT
#ifndef CGX_CPU_SIM
// the loader (jit.cu) reads the launch shape from the module it just built
extern "C" __device__ const int cgx_flat_jit_info[4] = {(int)sizeof(CtaSmem), FW_THREADS, FW_WARPS, FW_CTAS};
#endif
extern "C" __global__ void __launch_bounds__(FW_THREADS, FW_CTAS) cgx_flat_jit(const __grid_constant__ ScanArgs a) {
#else
namespace {
__global__ void __launch_bounds__(FW_THREADS, FW_CTAS) scan_flat_kernel(const __grid_constant__ ScanArgs a) {
#endif
  CGX_DYN_SMEM(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  CtaSmem& cs = *reinterpret_cast<CtaSmem*>(smem_raw);
  if (tid < 2 * FW_WARPS) {
    cs.mail[tid >> 1][tid & 1].state = 0;
    cs.done[tid >> 1] = 0;
  }
  static_assert(2 * FW_WARPS <= FW_THREADS, "one thread per mail slot at start-up");
  if (warp < FW_WARPS && lane == 0) {
    mbar_init(&cs.w[warp].mbar[0], 1);
    mbar_init(&cs.w[warp].mbar[1], 1);
    fence_mbar_init();
  }
  cgx_syncthreads();
  if (warp == FW_WARPS) {
    if (a.mode == M_FINDALL) resolver_warp(a, cs, lane);
    return;
  }
  WarpSmem& ws = cs.w[warp];

  uint32_t phase = 0;  // bit b = parity to wait for on mbar[b]
  auto take_ticket = [&]() -> unsigned {
    unsigned t = 0;
    if (lane == 0) {
      if (a.mode == M_ISMATCH && *((volatile unsigned long long*)&a.total[1])) t = 0xFFFFFFFFu;
      else t = atomicAdd(a.ticket, 1u);
    }
    return __shfl_sync(FULL, t, 0);
  };
  // starts the bulk copy of iteration `it` of `chunk` into window buffer `b`.  No proxy fence: the
  // buffer's previous contents were read with LDS whose results have all been consumed (they fed
  // the classification of an earlier iteration) before the warp gets here.
  auto issue = [&](int64_t chunk, int it, int b) {
    if (lane == 0) {
      const int64_t g = chunk * (int64_t)CHUNKB + it * (NT * STRIDE);
      if (g + SUPER <= a.n) {  // the common case: a whole window
        mbar_expect_tx(&ws.mbar[b], (uint32_t)SUPER);
        tma_load_1d(ws.win[b], a.h + g, (uint32_t)SUPER, &ws.mbar[b]);
      } else {
        const int64_t left = a.n - g;
        if (left > 0) {
          // whole 16-byte blocks: the last block may run up to 15 bytes past n, inside the caller's
          // 16-byte aligned allocation granule; those bytes are masked out (process_tiles, nv)
          const uint32_t bulk = (uint32_t)((left + 15) & ~(int64_t)15);
          mbar_expect_tx(&ws.mbar[b], bulk);
          tma_load_1d(ws.win[b], a.h + g, bulk, &ws.mbar[b]);
        } else {
          mbar_arrive(&ws.mbar[b]);
        }
      }
    }
  };
#if CGX_PAIR
  static_assert(NT == 1 && TPC % 2 == 0, "CGX_PAIR is a variant of the one-tile build");
  constexpr int PSUPER = STRIDE + TILE;  // two overlapping tiles
  // starts the bulk copy of tile pair `p` of `chunk` into the warp's (single) window
  auto issue_pair = [&](int64_t chunk, int p) {
    if (lane == 0) {
      const int64_t g = chunk * (int64_t)CHUNKB + p * (2 * STRIDE);
      if (g + PSUPER <= a.n) {
        mbar_expect_tx(&ws.mbar[0], (uint32_t)PSUPER);
        tma_load_1d(ws.win[0], a.h + g, (uint32_t)PSUPER, &ws.mbar[0]);
      } else {
        const int64_t left = a.n - g;
        if (left > 0) {
          const uint32_t bulk = (uint32_t)((left + 15) & ~(int64_t)15);  // < PSUPER + 16 <= 2 * SUPER
          mbar_expect_tx(&ws.mbar[0], bulk);
          tma_load_1d(ws.win[0], a.h + g, bulk, &ws.mbar[0]);
        } else {
          mbar_arrive(&ws.mbar[0]);
        }
      }
    }
  };
#endif
  // first window of a chunk (into the buffer the chunk will start with)
  auto issue_first = [&](int64_t chunk, int b) {
#if CGX_PAIR
    issue_pair(chunk, 0);
#else
    issue(chunk, 0, b);
#endif
  };
  auto wait = [&](int b) {
    mbar_wait(&ws.mbar[b], (phase >> b) & 1u);
    phase ^= 1u << b;
  };
  // Stores the chunk staged in buffer b (if any) once the resolver has supplied its offset;
  // returns whether the buffer is free afterwards.  Blocking, or one look.
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
    write_out(a, ws, b, m.chunk, m.cnt, m.excl, lane);
    __syncwarp();
    if (lane == 0) m.state = 0;
    __syncwarp();
    return true;
  };

  // Rule that keeps the look-back free of convoys: a warp only ever WAITS (for a staging buffer,
  // or for the offset of an overflowed chunk) while it holds no ticket it has not finished — every
  // ticket anybody holds is being scanned, so every count a waiter needs arrives within one chunk
  // time.  Hence the next ticket is drawn early (for the prefetch of its first window) only when
  // the staging buffer it will need is already free.
  // chunk numbers are 32-bit tickets (the host refuses haystacks of more than 0xFFFF0000 chunks)
  const unsigned none = 0xFFFFFFFEu;
  const unsigned nch = (unsigned)a.nchunks;
  unsigned cur = take_ticket();
  unsigned nxt = none;
  bool prefetched = false;  // window 0 of `nxt` is in flight into the buffer the next chunk starts with
  int kb = 0, sb = 0;
  if (cur < nch) issue_first(cur, 0);
  bool direct = false;
  while (cur < nch) {
    // one pass over the chunk: staged (normal) or, after a staging overflow, with direct stores
    {
      const int64_t cbeg = cur * (int64_t)CHUNKB;
      if (lane == 0) {
        ws.cb = cbeg;
        ws.csrc = a.h + cbeg;
        ws.whole = cbeg + (CHUNKB + TILE - STRIDE) <= a.n;
        ws.chunk0 = cur == 0;
      }
      __syncwarp();
    }
    Emit em{&a, &ws, sb, direct};
    unsigned cnt = 0;
#if CGX_PAIR
    for (int p = 0; p < TPC / 2; p++) {
      // what fills the window once this pair's second tile is classified — decided here, while
      // nothing of a tile is live in registers
      unsigned fill = none;
      int fillp = 0;
      if (p + 1 < TPC / 2) {
        fill = cur;
        fillp = p + 1;
      } else {
        if (nxt == none &&
            (a.mode != M_FINDALL || direct || (cnt <= (unsigned)CAP && flush(sb ^ 1, false))))
          nxt = take_ticket();
        if (nxt != none && nxt < nch) {
          fill = nxt;
          prefetched = true;
        }
      }
      auto refill = [&]() {
        __syncwarp();  // every lane has its piece of the window in registers
        if (fill != none) issue_pair(fill, fillp);
      };
      wait(0);
      const int wrel = p * (2 * STRIDE);
      int nvp = PSUPER;
      if (!ws.whole) {
        const int64_t left = a.n - (ws.cb + wrel);
        nvp = left >= PSUPER ? PSUPER : (left > 0 ? (int)left : 0);
      }
      const int nva = nvp < TILE ? nvp : TILE;
      const int nvb = nvp - STRIDE < TILE ? nvp - STRIDE : TILE;
      const bool first = p == 0 && ws.chunk0 != 0;
      // one copy of the tile code, run once or twice (a second inlined copy would double the hot
      // loop's instruction-cache footprint)
      const int ntl = nvb > 0 ? 2 : (nva > 0 ? 1 : 0);
      if (ntl == 0) refill();
#pragma unroll 1
      for (int t = 0; t < ntl; t++) {
        auto hook = [&]() {
          if (t == ntl - 1) refill();
        };
        process_tiles(a, em, ws.win[0] + t * STRIDE, wrel + t * STRIDE, t ? nvb : nva, first && t == 0, cnt, lane,
                      hook);
      }
    }
#else
#pragma unroll IT_UNROLL
    for (int it = 0; it < ITERS; it++) {
      // the other window buffer was last read an iteration ago: refill it now
      __syncwarp();
      const int whole = ws.whole;
      if (it + 1 < ITERS) {
        if (whole) {
          // a whole window of a chunk that lies inside the input: no bounds to look at
          if (lane == 0) {
            mbar_expect_tx(&ws.mbar[kb ^ 1], (uint32_t)SUPER);
            tma_load_1d(ws.win[kb ^ 1], ws.csrc + (it + 1) * (NT * STRIDE), (uint32_t)SUPER, &ws.mbar[kb ^ 1]);
          }
        } else {
          issue(cur, it + 1, kb ^ 1);
        }
      } else {
        if (nxt == none &&
            (a.mode != M_FINDALL || direct || (cnt <= (unsigned)CAP && flush(sb ^ 1, false))))
          nxt = take_ticket();
        if (nxt != none && nxt < nch) {
          issue(nxt, 0, kb ^ 1);
          prefetched = true;
        }
      }
      wait(kb);
      const int wrel = it * (NT * STRIDE);
      int nv = SUPER;
      if (!whole) {
        const int64_t left = a.n - (ws.cb + wrel);
        nv = left >= SUPER ? SUPER : (left > 0 ? (int)left : 0);
      }
      if (nv > 0) process_tiles(a, em, ws.win[kb], wrel, nv, it == 0 && ws.chunk0 != 0, cnt, lane, NoHook());
      kb ^= 1;
    }
#endif
    if (a.mode != M_FINDALL) {
      if (cnt && lane == 0) {
        atomicAdd(a.total, (unsigned long long)cnt);
        a.total[1] = 1ull;
      }
    } else if (!direct) {
      publish_count(a, cur, cnt, lane);  // everybody behind us can go on
      if (cnt > (unsigned)CAP || __any_sync(FULL, em.far)) {
        // more matches than the staging buffer holds: get the offset now and run the chunk again
        // with direct stores.  A prefetched window of the next chunk is dropped and re-issued later.
        const unsigned long long goff = lb_resolve(a, cur, lane);
        if (lane == 0) {
          ws.goff = goff;
          atomicAdd(&a.total[3], 1ull);  // diagnostics: chunks redone with direct stores
          if (cur == nch - 1u) a.total[0] = goff + cnt;
        }
        direct = true;
        if (prefetched) wait(kb);
        prefetched = false;
        __syncwarp();
        issue_first(cur, kb);
        continue;
      }
      // hand the chunk to the resolver warp
      __syncwarp();  // staged matches written
      if (lane == 0) {
        Mail& m = cs.mail[warp][sb];
        m.chunk = cur;
        m.cnt = cnt;
        cgx_fence_block();
        m.state = 1;
      }
      sb ^= 1;
      flush(sb, true);  // (no unfinished ticket is held here unless the buffer was free)
    }
    direct = false;
    if (nxt == none) nxt = take_ticket();
    if (!prefetched && nxt < nch) {
      __syncwarp();
      issue_first(nxt, kb);
    }
    prefetched = false;
    cur = nxt;
    nxt = none;
  }
  if (a.mode == M_FINDALL) {
    flush(sb, true);
    flush(sb ^ 1, true);
  }
  __syncwarp();
  if (lane == 0) {
    cgx_fence_block();
    cs.done[warp] = 1;
  }
}
#ifndef CGX_JIT
}  // namespace
#endif

#if !defined(CGX_JIT) || defined(CGX_CPU_SIM)
int64_t scan_flat_chunks(int64_t n) {
  if (n <= 0) return 0;
  const int64_t tiles = (n + STRIDE - 1) / STRIDE;
  return (tiles + TPC - 1) / TPC;
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
// Launches the scan on `stream`.  ticket/status/total must be zeroed by the caller.
cudaError_t launch_scan_flat(const ScanArgs& a, int sm_count, cudaStream_t stream) {
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
  scan_flat_kernel<<<(unsigned)grid, FW_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}
#endif

#endif  // host side

}  // namespace cgx
