// oracle/teddy.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// CPU restatement of the reference's Teddy multi-literal searchers.
// Slim Teddy (2-32 literals, 8 buckets):
//   reference prefilter/teddy.go:189-256 (NewTeddy), :271-311 (buildMasks, bucket = id % min(8,n)),
//     :327-389 (Find), :391-445 (FindMatch), :447-458 (findMatchScalar, haystack[start:] < 16 bytes),
//     :491-521 (findScalarCandidate — the executable spec of the SSSE3/AVX2 kernels),
//     :532-550 (verifyBucket: insertion order inside a bucket, buckets tried low -> high)
// Fat Teddy (33-64 literals, 16 buckets):
//   reference prefilter/teddy_fat.go:127-197, :200-236 (buildFatMasks, bucket = id % 16),
//     :348-393 (FindMatch), :407-424 (scalar regimes), :426-459 (findScalarCandidate), :461-477
// The asm kernels (prefilter/teddy_ssse3_amd64.s:48,273; prefilter/teddy_avx2_amd64.s:43) return
// the same (pos, bucketMask) as the scalar twins for every input (first i with i+fpLen<=len and a
// non-zero mask); the fat AVX2 kernel may add spurious buckets at position 0 (prev0=0xFF), which
// full verification makes unobservable.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace oracle {

class TeddyBase {
 public:
  // nbuckets_max = 8 (slim) or 16 (fat)
  TeddyBase(const std::vector<std::string>& patterns, int nbuckets_max, bool slim);
  bool ok() const { return ok_; }
  // (start,end) of the first match at or after `start`; false if none
  bool FindMatch(const uint8_t* h, int64_t n, int64_t start, int64_t& ms, int64_t& me) const;
  int64_t Find(const uint8_t* h, int64_t n, int64_t start) const;

  int fingerprint_len() const { return fp_len_; }
  const std::vector<std::vector<int>>& buckets() const { return buckets_; }
  const std::vector<std::string>& patterns() const { return patterns_; }
  // nibble tables: lo_[pos][nibble] / hi_[pos][nibble] as 16-bit bucket masks
  uint16_t lo(int pos, int nib) const { return lo_[pos][nib]; }
  uint16_t hi(int pos, int nib) const { return hi_[pos][nib]; }

 private:
  bool ok_ = false;
  std::vector<std::string> patterns_;
  std::vector<std::vector<int>> buckets_;
  int fp_len_ = 0;
  size_t min_len_ = 0;
  uint16_t lo_[4][16], hi_[4][16];

  // first candidate in h[0..n): returns pos or -1, mask out
  int64_t candidate(const uint8_t* h, int64_t n, uint32_t& mask) const;
  bool matchScalar(const uint8_t* h, int64_t n, int64_t& ms, int64_t& me) const;
};

class Teddy : public TeddyBase {
 public:
  explicit Teddy(const std::vector<std::string>& p) : TeddyBase(p, 8, true) {}
};
class FatTeddy : public TeddyBase {
 public:
  explicit FatTeddy(const std::vector<std::string>& p) : TeddyBase(p, 16, false) {}
};

}  // namespace oracle
