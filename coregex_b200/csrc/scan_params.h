// scan_params.h — plain structs shared by the host engine and the sm_100a scan kernels.
#pragma once
#ifdef __CUDACC_RTC__
// NVRTC (csrc/jit.cpp) compiles without host headers
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef signed char int8_t;
typedef short int16_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned long size_t;
#else
#include <cstdint>
#endif

namespace cgx {

// Which candidate filter phase A runs (what the reference calls the prefilter):
//   F_RUNSTART : first byte of every run of bytes in `ranges`  (DigitPrefilter + digitRunSkipSafe,
//                reference meta/find_indices.go:1058-1085, meta/strategy.go:525-560)
//   F_BYTESET  : every byte in `ranges` (<= 4 ranges, SWAR compares)
//   F_LUT      : every byte b with lut[b] != 0 (arbitrary first-byte sets)
enum FilterKind : int { F_RUNSTART = 0, F_BYTESET = 1, F_LUT = 2 };

enum ScanMode : int { M_FINDALL = 0, M_COUNT = 1, M_ISMATCH = 2 };

struct DfaDev {
  const uint16_t* trans;  // nstates*256 (global; staged to shared memory by the kernel)
  const uint8_t* eoi;     // nstates
  int nstates;
  uint16_t start[5];
  uint8_t kind_lut_needed;  // 1 when start state depends on the previous byte
};

struct FilterDev {
  int kind;
  int nranges;
  uint8_t lo[4], hi[4];
  const uint8_t* lut;  // 256 bytes (F_LUT)
};

// Bit-parallel "can a match start here" filter for flat patterns (a concatenation of byte-class
// items with quantifiers).  Evaluated right-to-left over per-class position bitmaps, one 32-bit
// word per lane, with warp-wide carry chains (the position-parallel formulation of an NFA walk).
// It only has to be a superset of the true match starts: every survivor is still verified by the
// anchored DFA.  kind: 0 = one byte of the class, 1 = class+, 2 = class*, 3 = class?
struct FlatDev {
  int nops;                // 0 = no second-level filter
  int nclasses;            // <= 4
  int first_is_filter;     // class 0 has exactly the first-level filter's ranges
  uint8_t op_kind[24];
  uint8_t op_class[24];    // ops in PATTERN order (host/debug view)
  // what the kernel executes: evaluation runs right to left; trailing nullable items are dropped
  // and the first concrete item becomes the initial marker set (M = class rev_init_class);
  // rev_ops[i] = kind | class<<2 for the remaining items, already in right-to-left order
  int rev_nops;
  int rev_init_class;
  uint8_t rev_ops[24];
  uint8_t cls_nranges[4];
  uint8_t cls_lo[4][4], cls_hi[4][4];
  // SWAR constants per (class, range), precomputed on the host so the kernel reads them straight
  // from the constant bank: mode 0 = XOR-alignable range (3 ops/word), 1 = generic (5 ops/word)
  uint8_t cls_mode[4][4];
  uint32_t cls_k1[4][4], cls_k2[4][4];
  // ---- bitstream engine (scan_bits.cu) --------------------------------------------------------
  // bs_ok: the pattern is flat AND deterministic (every repeated/optional item's class is disjoint
  // from whatever can follow it), so the leftmost-first match from a start is the forced greedy
  // one and its END can be computed by a forward marker pass over the same class bitmaps.
  int bs_ok;
  int bs_runstart;       // starts are restricted to the first byte of a run of class 0 (pattern opens with C+)
  int bs_midrun_check;   // a match may end in the middle of a class-0 run: such tiles take the serial path
  int fwd_nops;
  uint8_t fwd_ops[24];   // kind | class<<2, pattern order
  uint32_t sync_lut[8];  // bit b set <=> byte b belongs to no class: no match can contain it
};

// Multi-literal engine tables (reference prefilter/teddy.go, teddy_fat.go), device resident.
struct TeddyDev {
  const uint32_t* fp;        // 256 entries: fp0[b] | fp1[b] << 16  (bucket masks per byte value)
  // per literal id: its first 8 bytes as a little-endian word (zero padded), then npat byte masks
  // (0xFF for the bytes that exist): a candidate is compared 8 bytes at a time
  const uint64_t* lit8;
  // 256 entries: bucket masks per byte value of the THIRD byte (every literal has three bytes or
  // more): the candidate filter tests three bytes, an order of magnitude fewer candidates to verify
  const uint16_t* fp2;
  // 1: phase A tests the third byte as well.  It pays when two-byte candidates are frequent in text
  // (literals that open with two letters: 5 % of positions for 16 English words, C3 892 -> 1308
  // GB/s); for sets like `k00z..k63z` they are rare anyway and the third lookup per byte only costs
  // (C5 965 -> 890 GB/s), so the host switches it per literal set.
  int use_fp2;
  const uint8_t* bytes;      // concatenated literals
  const int32_t* offs;       // npat + 1
  const uint16_t* order;     // literal ids, bucket-major (SIMD-regime verify order)
  const uint16_t* bucket_off;  // nbuckets + 1 offsets into `order`
  int npat, nbuckets, min_len, max_len, bytes_len;
  int blob_bytes;            // size of the single allocation that starts at `fp`
};

// Record engine tables: unanchored forward DFA (u*) and reverse DFA (r*), one u16 blob
// utrans[un*256] | rtrans[rn*256] | ueoi[un] | reoi[rn]; entries = next | flag<<15 as in DfaDev.
struct LineDev {
  const uint16_t* blob;
  int un, rn;
  uint16_t ustart[5], rstart[5];
  uint8_t ukinds, rkinds;  // 1 when the start state depends on the neighbouring byte
  int blob_bytes;
};

enum EngineSel : int { SEL_DFA = 0, SEL_TEDDY = 1, SEL_LINE = 2 };

struct ScanArgs {
  const uint8_t* h;   // device haystack, 16-byte aligned
  int64_t n;
  int64_t base;       // position of h[0] in the logical haystack: added to every reported offset;
                      // base > 0 means h[-1] is the record delimiter (shards are cut at records)
  int64_t after;      // bytes of the logical haystack after h[n-1] (0: h ends the haystack)
  DfaDev dfa;
  FilterDev filter;
  FlatDev flat;
  TeddyDev teddy;
  LineDev line;
  int engine;         // EngineSel
  int skip_safe;      // 1: after a match, a candidate in the middle of a run must still be tried
  uint8_t delim;      // record delimiter no match can contain
  int mode;
  int64_t* out;       // int64 pairs (start,end), capacity cap pairs
  int64_t cap;
  unsigned long long* total;   // [0]=match count, [1]=is-match flag
  unsigned int* ticket;        // chunk ticket counter (zeroed before launch)
  unsigned long long* status;  // nchunks look-back words (zeroed before launch)
  int64_t nchunks;
  // bitstream kernel: one word and one arrival accumulator per group of 32 chunks (zeroed before launch)
  unsigned long long* gstatus;
  unsigned long long* gacc;
  // bitstream kernel: look-back words of another epoch read as empty (nothing is cleared between
  // launches; 1 .. 0xFFFFF, the host clears the words when the counter wraps)
  unsigned int epoch;
  // bitstream kernel, FindAll: {total, is-match flag} also written here by the kernel itself
  // (device pointer, may be null), so that a FindAll call is exactly one launch
  unsigned long long* result;
};

}  // namespace cgx
