// synth.h — deterministic synthetic corpora, identical on host and device.
//
// The corpus is a sequence of independent fixed-size blocks; block i depends only on
// (kind, seed, i), so a 16 GB corpus can be generated shard by shard directly in HBM and any
// slice can be regenerated on the host for the CPU oracle.  Every block ends with '\n'.
//   kind 0  access-log lines, 4096-byte blocks:
//           {a}.{b}.{c}.{d} - {user} [{dd}/{Mon}/{yyyy}:{hh}:{mm}:{ss} +0000] "{METHOD} /{path} HTTP/1.1" {status} {bytes}
//           5 % of lines carry a second IP inside the path, 5 % a near-miss (1.2.3 / 1.2.3. / 1..2.3.4)
//   kind 1  lowercase word text with literals planted about every 200 bytes, 4096-byte blocks
//   kind 2  80-byte lines ([a-z ] filler, exactly one user@host.tld per line), 80-byte blocks
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CGX_HD __host__ __device__
#else
#define CGX_HD
#endif

namespace cgx {
namespace synth {

constexpr int kBlock01 = 4096;
constexpr int kBlock2 = 80;

struct Rng {
  uint64_t s;
  CGX_HD explicit Rng(uint64_t seed, uint64_t block) {
    s = seed * 0x9E3779B97F4A7C15ull + block * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
    next();
    next();
  }
  CGX_HD uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  CGX_HD uint32_t below(uint32_t n) { return (uint32_t)((next() >> 33) % n); }
};

CGX_HD inline int put_uint(uint8_t* o, uint32_t v, int min_digits) {
  char tmp[12];
  int k = 0;
  do {
    tmp[k++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (k < min_digits) tmp[k++] = '0';
  for (int i = 0; i < k; i++) o[i] = (uint8_t)tmp[k - 1 - i];
  return k;
}
CGX_HD inline int put_str(uint8_t* o, const char* s) {
  int k = 0;
  while (s[k]) {
    o[k] = (uint8_t)s[k];
    k++;
  }
  return k;
}
CGX_HD inline int put_ip(uint8_t* o, Rng& r) {
  int k = 0;
  for (int i = 0; i < 4; i++) {
    k += put_uint(o + k, r.below(256), 1);
    if (i < 3) o[k++] = '.';
  }
  return k;
}

CGX_HD inline int log_line(uint8_t* o, Rng& r) {
  const char* months[12] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
  const char* methods[4] = {"GET", "POST", "PUT", "DELETE"};
  const char pathc[41] = "abcdefghijklmnopqrstuvwxyz0123456789/_.-";
  int k = 0;
  k += put_ip(o + k, r);
  k += put_str(o + k, " - ");
  int ul = 3 + (int)r.below(6);
  for (int i = 0; i < ul; i++) o[k++] = (uint8_t)('a' + r.below(26));
  k += put_str(o + k, " [");
  k += put_uint(o + k, 1 + r.below(28), 2);
  o[k++] = '/';
  k += put_str(o + k, months[r.below(12)]);
  o[k++] = '/';
  k += put_uint(o + k, 2000 + r.below(27), 4);
  o[k++] = ':';
  k += put_uint(o + k, r.below(24), 2);
  o[k++] = ':';
  k += put_uint(o + k, r.below(60), 2);
  o[k++] = ':';
  k += put_uint(o + k, r.below(60), 2);
  k += put_str(o + k, " +0000] \"");
  k += put_str(o + k, methods[r.below(4)]);
  k += put_str(o + k, " /");
  int pl = 8 + (int)r.below(33);
  uint32_t special = r.below(100);
  for (int i = 0; i < pl; i++) o[k++] = (uint8_t)pathc[r.below(40)];
  if (special < 5) {  // second IP inside the path
    o[k++] = '/';
    k += put_ip(o + k, r);
  } else if (special < 10) {  // near misses
    o[k++] = '/';
    uint32_t v = r.below(3);
    if (v == 0) k += put_str(o + k, "1.2.3");
    else if (v == 1) k += put_str(o + k, "1.2.3.");
    else k += put_str(o + k, "1..2.3.4x");
  }
  k += put_str(o + k, " HTTP/1.1\" ");
  const uint32_t st[6] = {200, 200, 200, 301, 404, 500};
  k += put_uint(o + k, st[r.below(6)], 3);
  o[k++] = ' ';
  k += put_uint(o + k, r.below(100000), 1);
  o[k++] = '\n';
  return k;
}

constexpr int kMaxLogLine = 176;

CGX_HD inline void fill_tail(uint8_t* o, int rem, Rng& r) {
  // one filler line of letters so the block ends exactly with '\n'
  if (rem <= 0) return;
  for (int i = 0; i < rem - 1; i++) o[i] = (uint8_t)('a' + r.below(26));
  o[rem - 1] = '\n';
}

CGX_HD inline void gen_block_log(uint8_t* o, uint64_t seed, uint64_t block) {
  Rng r(seed, block);
  int k = 0;
  while (kBlock01 - k >= kMaxLogLine) k += log_line(o + k, r);
  fill_tail(o + k, kBlock01 - k, r);
}

CGX_HD inline void gen_block_text(uint8_t* o, uint64_t seed, uint64_t block, const uint8_t* lits,
                                  const int32_t* offs, int nlit) {
  Rng r(seed ^ 0x7E47, block);
  int k = 0;
  int line_left = 60 + (int)r.below(61);
  while (k < kBlock01 - 1) {
    int wl;
    bool plant = nlit > 0 && r.below(31) == 0;
    int li = 0;
    if (plant) {
      li = (int)r.below((uint32_t)nlit);
      wl = offs[li + 1] - offs[li];
    } else {
      wl = 2 + (int)r.below(8);
    }
    if (k + wl + 1 > kBlock01 - 1) break;
    if (plant)
      for (int i = 0; i < wl; i++) o[k++] = lits[offs[li] + i];
    else
      for (int i = 0; i < wl; i++) o[k++] = (uint8_t)('a' + r.below(26));
    line_left -= wl + 1;
    if (line_left <= 0) {
      o[k++] = '\n';
      line_left = 60 + (int)r.below(61);
    } else {
      o[k++] = ' ';
    }
  }
  while (k < kBlock01 - 1) o[k++] = ' ';
  o[kBlock01 - 1] = '\n';
}

CGX_HD inline void gen_block_email(uint8_t* o, uint64_t seed, uint64_t line) {
  Rng r(seed ^ 0xE3A11, line);
  for (int i = 0; i < 79; i++) {
    uint32_t v = r.below(30);
    o[i] = v < 26 ? (uint8_t)('a' + v) : (uint8_t)' ';
  }
  o[79] = '\n';
  int ul = 3 + (int)r.below(6), hl = 3 + (int)r.below(6), tl = 2 + (int)r.below(2);
  int total = ul + 1 + hl + 1 + tl;
  int col = 1 + (int)r.below((uint32_t)(79 - total - 1));
  int k = col;
  o[k - 1] = ' ';
  for (int i = 0; i < ul; i++) o[k++] = (uint8_t)('a' + r.below(26));
  o[k++] = '@';
  for (int i = 0; i < hl; i++) o[k++] = (uint8_t)('a' + r.below(26));
  o[k++] = '.';
  for (int i = 0; i < tl; i++) o[k++] = (uint8_t)('a' + r.below(26));
  if (k < 79) o[k] = ' ';
}

}  // namespace synth
}  // namespace cgx
