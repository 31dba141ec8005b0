// oracle/simd.cpp — TEST INFRASTRUCTURE ONLY.  See simd.h for the reference citations.
#include "simd.h"

#if defined(__AVX2__)
#include <immintrin.h>
#endif

namespace oracle {

static inline bool is_word(uint8_t b) {
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}

int64_t memchr1(const uint8_t* h, int64_t n, uint8_t needle) {
  int64_t i = 0;
#if defined(__AVX2__)
  if (n >= 32) {
    const __m256i v = _mm256_set1_epi8((char)needle);
    for (; i + 32 <= n; i += 32) {
      __m256i x = _mm256_loadu_si256((const __m256i*)(h + i));
      uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, v));
      if (m) return i + __builtin_ctz(m);
    }
  }
#endif
  for (; i < n; i++)
    if (h[i] == needle) return i;
  return -1;
}

int64_t memchr2(const uint8_t* h, int64_t n, uint8_t a, uint8_t b) {
  int64_t i = 0;
#if defined(__AVX2__)
  if (n >= 32) {
    const __m256i va = _mm256_set1_epi8((char)a), vb = _mm256_set1_epi8((char)b);
    for (; i + 32 <= n; i += 32) {
      __m256i x = _mm256_loadu_si256((const __m256i*)(h + i));
      __m256i e = _mm256_or_si256(_mm256_cmpeq_epi8(x, va), _mm256_cmpeq_epi8(x, vb));
      uint32_t m = (uint32_t)_mm256_movemask_epi8(e);
      if (m) return i + __builtin_ctz(m);
    }
  }
#endif
  for (; i < n; i++)
    if (h[i] == a || h[i] == b) return i;
  return -1;
}

int64_t memchr3(const uint8_t* h, int64_t n, uint8_t a, uint8_t b, uint8_t c) {
  for (int64_t i = 0; i < n; i++)
    if (h[i] == a || h[i] == b || h[i] == c) return i;
  return -1;
}

int64_t memchr_digit(const uint8_t* h, int64_t n) {
  int64_t i = 0;
#if defined(__AVX2__)
  // reference simd/memchr_digit_amd64.go:25: AVX2 only when len >= 32, scalar tail
  if (n >= 32) {
    const __m256i lo = _mm256_set1_epi8('0' - 1), hi = _mm256_set1_epi8('9' + 1);
    for (; i + 32 <= n; i += 32) {
      __m256i x = _mm256_loadu_si256((const __m256i*)(h + i));
      // signed compares are fine: bytes >= 0x80 are negative, hence < '0'
      __m256i ge = _mm256_cmpgt_epi8(x, lo);
      __m256i le = _mm256_cmpgt_epi8(hi, x);
      uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_and_si256(ge, le));
      if (m) return i + __builtin_ctz(m);
    }
  }
#endif
  for (; i < n; i++)
    if (h[i] >= '0' && h[i] <= '9') return i;
  return -1;
}

int64_t memchr_digit_at(const uint8_t* h, int64_t n, int64_t at) {
  if (at < 0 || at >= n) return -1;
  int64_t p = memchr_digit(h + at, n - at);
  return p < 0 ? -1 : p + at;
}

int64_t memchr_word(const uint8_t* h, int64_t n) {
  for (int64_t i = 0; i < n; i++)
    if (is_word(h[i])) return i;
  return -1;
}

int64_t memchr_not_word(const uint8_t* h, int64_t n) {
  for (int64_t i = 0; i < n; i++)
    if (!is_word(h[i])) return i;
  return -1;
}

}  // namespace oracle
