// scan_params.h — plain structs shared by the host engine and the sm_100a scan kernels.
#pragma once
#include <cstdint>

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

struct ScanArgs {
  const uint8_t* h;   // device haystack, 16-byte aligned
  int64_t n;
  int64_t base;       // added to every reported offset (shard base)
  DfaDev dfa;
  FilterDev filter;
  int skip_safe;      // 1: after a match, a candidate in the middle of a run must still be tried
  uint8_t delim;      // record delimiter no match can contain
  int mode;
  int64_t* out;       // int64 pairs (start,end), capacity cap pairs
  int64_t cap;
  unsigned long long* total;   // [0]=match count, [1]=is-match flag
  unsigned int* ticket;        // chunk ticket counter (zeroed before launch)
  unsigned long long* status;  // nchunks look-back words (zeroed before launch)
  int64_t nchunks;
};

}  // namespace cgx
