// oracle/simd.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// CPU restatement of the reference's SIMD primitives.  The scalar twins are the executable
// spec (reference simd/memchr_generic_impl.go:23-253); when the host has AVX2 the 32-byte
// loops mirror the shape of the reference's assembly so the CPU baseline is a fair stand-in:
//   memchr        — reference simd/memchr_amd64.go:67, asm simd/memchr_amd64.s:25
//   memchr2/3     — reference simd/memchr_amd64.go:114,159, asm :140,:241
//   memchr_digit  — reference simd/memchr_digit_amd64.go:17,34, asm simd/memchr_digit_amd64.s:26
//   memchr_word / memchr_not_word — reference simd/memchr_class_amd64.go:35,58
#pragma once
#include <cstdint>

namespace oracle {

int64_t memchr1(const uint8_t* h, int64_t n, uint8_t needle);
int64_t memchr2(const uint8_t* h, int64_t n, uint8_t a, uint8_t b);
int64_t memchr3(const uint8_t* h, int64_t n, uint8_t a, uint8_t b, uint8_t c);
int64_t memchr_digit(const uint8_t* h, int64_t n);
// reference simd/memchr_digit_amd64.go:34 MemchrDigitAt
int64_t memchr_digit_at(const uint8_t* h, int64_t n, int64_t at);
int64_t memchr_word(const uint8_t* h, int64_t n);
int64_t memchr_not_word(const uint8_t* h, int64_t n);

}  // namespace oracle
