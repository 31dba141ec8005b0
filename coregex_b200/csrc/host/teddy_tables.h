// teddy_tables.h — tables for the GPU multi-literal engine.
//
// Same mathematics as the reference's Teddy (prefilter/teddy.go:271-311 buildMasks,
// prefilter/teddy_fat.go:200-236 buildFatMasks): pattern id -> bucket id % nbuckets, per
// fingerprint position a low-nibble and a high-nibble table of bucket bitmasks; a position is a
// candidate when AND over fingerprint bytes of (lo[p][b&15] & hi[p][b>>4]) is non-zero.
// The GPU indexes one 256-entry byte table per position (shared memory has no PSHUFB to feed), so
// the table can be EXACT — fp[p][b] = buckets holding a literal whose byte p is b — instead of the
// nibble product lo[p][b&15] & hi[p][b>>4], which also lets through every byte that shares a low
// nibble with one literal and a high nibble with another.  Candidates are only ever a superset of the
// match starts and every candidate is verified byte for byte, so the result is the reference's; the
// exact table just verifies fewer positions.
// Verification order is part of the observable behaviour (reference prefilter/teddy.go:415-428,
// :540-546 vs :447-458): `order_simd` lists pattern ids bucket-major (low bucket first, insertion
// order inside a bucket); the "scalar" regime (haystack[start:] shorter than 16 bytes) uses plain
// pattern-id order.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cgx {

struct TeddyTables {
  int npat = 0;
  int nbuckets = 0;
  int fp_len = 0;
  int min_len = 0, max_len = 0;
  std::vector<uint8_t> bytes;      // concatenated patterns
  std::vector<int32_t> offs;       // npat+1
  std::vector<uint16_t> fp0, fp1;  // 256-entry byte tables for fingerprint positions 0 and 1
  std::vector<uint16_t> order_simd;  // pattern ids, bucket-major
  std::vector<uint8_t> bucket_of;    // per pattern id
  std::vector<uint16_t> bucket_off;  // nbuckets+1 offsets into order_simd
  std::vector<uint32_t> fp_packed;   // fp0 | fp1<<16
};

// patterns in reference literal order; returns false when the reference's NewTeddy/NewFatTeddy
// would refuse (count out of range, a pattern shorter than 3 bytes).  max_patterns = 64 is the
// reference's Fat Teddy limit; larger sets (the reference's Aho-Corasick strategy) pass their own
// bound and get 16 buckets like Fat Teddy.
bool BuildTeddyTables(const std::vector<std::string>& patterns, TeddyTables& out, size_t max_patterns = 64);

}  // namespace cgx
