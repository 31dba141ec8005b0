// oracle/teddy.cpp — TEST INFRASTRUCTURE ONLY.  See teddy.h for the reference citations.
#include "teddy.h"

#include <cstring>

namespace oracle {

TeddyBase::TeddyBase(const std::vector<std::string>& patterns, int nbuckets_max, bool slim) {
  const size_t kMinPatterns = 2, kMinLen = 3;
  size_t max_patterns = slim ? 32 : 64;
  if (patterns.size() < kMinPatterns || patterns.size() > max_patterns) return;
  size_t mn = patterns[0].size();
  for (auto& p : patterns) {
    if (p.size() < kMinLen) return;
    if (p.size() < mn) mn = p.size();
  }
  min_len_ = mn;
  fp_len_ = 2;
  if ((size_t)fp_len_ > mn) fp_len_ = (int)mn;
  if (fp_len_ > 4) fp_len_ = 4;
  patterns_ = patterns;
  memset(lo_, 0, sizeof lo_);
  memset(hi_, 0, sizeof hi_);
  // slim: numBuckets = min(8, n) (teddy.go:277-281); fat: always 16 (teddy_fat.go:205)
  int nb = nbuckets_max;
  if (slim && (int)patterns.size() < nb) nb = (int)patterns.size();
  buckets_.assign(nb, {});
  for (size_t id = 0; id < patterns.size(); id++) {
    int b = (int)(id % nb);
    buckets_[b].push_back((int)id);
    uint16_t bit = (uint16_t)(1u << b);
    for (int pos = 0; pos < fp_len_; pos++) {
      uint8_t c = (uint8_t)patterns[id][pos];
      lo_[pos][c & 15] |= bit;
      hi_[pos][c >> 4] |= bit;
    }
  }
  ok_ = true;
}

int64_t TeddyBase::candidate(const uint8_t* h, int64_t n, uint32_t& mask) const {
  for (int64_t i = 0; i + fp_len_ <= n; i++) {
    uint32_t m = 0xFFFF;
    for (int p = 0; p < fp_len_; p++) {
      uint8_t b = h[i + p];
      m &= lo_[p][b & 15] & hi_[p][b >> 4];
    }
    if (m) {
      mask = m;
      return i;
    }
  }
  return -1;
}

bool TeddyBase::matchScalar(const uint8_t* h, int64_t n, int64_t& ms, int64_t& me) const {
  for (int64_t i = 0; i < n - (int64_t)min_len_ + 1; i++)
    for (auto& p : patterns_)
      if (i + (int64_t)p.size() <= n && memcmp(h + i, p.data(), p.size()) == 0) {
        ms = i;
        me = i + (int64_t)p.size();
        return true;
      }
  return false;
}

bool TeddyBase::FindMatch(const uint8_t* h0, int64_t n0, int64_t start, int64_t& ms,
                          int64_t& me) const {
  if (start < 0 || start >= n0) return false;
  const uint8_t* h = h0 + start;
  int64_t n = n0 - start;
  if (n < 16) {
    if (!matchScalar(h, n, ms, me)) return false;
    ms += start;
    me += start;
    return true;
  }
  int64_t acc = 0;
  uint32_t mask;
  int64_t pos = candidate(h, n, mask);
  while (pos != -1) {
    while (mask) {
      int bucket = __builtin_ctz(mask);
      mask &= mask - 1;
      if (bucket < (int)buckets_.size()) {
        int64_t rem = n - acc;
        for (int id : buckets_[bucket]) {
          const std::string& p = patterns_[id];
          int64_t end = pos + (int64_t)p.size();
          if (end <= rem && memcmp(h + acc + pos, p.data(), p.size()) == 0) {
            ms = start + acc + pos;
            me = ms + (int64_t)p.size();
            return true;
          }
        }
      }
    }
    int64_t next = acc + pos + 1;
    if (next >= n) break;
    acc = next;
    pos = candidate(h + acc, n - acc, mask);
  }
  return false;
}

int64_t TeddyBase::Find(const uint8_t* h, int64_t n, int64_t start) const {
  int64_t ms, me;
  if (!FindMatch(h, n, start, ms, me)) return -1;
  return ms;
}

}  // namespace oracle
