// coregex.hpp — header-only C++ mirror of the reference's public Go API for the bulk-scan path
// (reference regex.go: Compile :110, MustCompile :129, Match :282, FindAllIndex :695, Count :1349,
// FindAllSubmatchIndex :1423, NumSubexp :552, String :444) over the C ABI in coregex_b200.h.
// The reference is compiled code (Go); with no Go toolchain in this image the exercised host-side
// mirror above the C ABI is this header (and the ctypes binding used by the tests).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "coregex_b200.h"

namespace coregex {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class Regex {
 public:
  Regex(const Regex&) = delete;
  Regex& operator=(const Regex&) = delete;
  Regex(Regex&& o) noexcept : h_(o.h_), pattern_(std::move(o.pattern_)) { o.h_ = nullptr; }
  ~Regex() { if (h_) cgx_free(h_); }

  // reference regex.go:110 — returns (regex, error) in Go; throws Error with the same message here
  static Regex Compile(const std::string& pattern) {
    cgx_regex* h = nullptr;
    char err[1024] = {0};
    if (cgx_compile(pattern.data(), pattern.size(), &h, err, sizeof err) != CGX_OK) throw Error(err);
    return Regex(h, pattern);
  }
  // reference regex.go:129 — panic text "regexp: Compile(`pat`): <error>"
  static Regex MustCompile(const std::string& pattern) {
    try {
      return Compile(pattern);
    } catch (const Error& e) {
      throw Error("regexp: Compile(`" + pattern + "`): " + e.what());
    }
  }

  const std::string& String() const { return pattern_; }
  int NumSubexp() const { return cgx_num_captures(h_) - 1; }
  std::string Strategy() const { return cgx_strategy(h_); }

  bool Match(const uint8_t* b, size_t n) const {
    int m = 0;
    check(cgx_is_match(h_, b, n, &m));
    return m != 0;
  }
  size_t Count(const uint8_t* b, size_t n, int64_t limit = -1) const {
    size_t c = 0;
    check(cgx_count(h_, b, n, limit, &c));
    return c;
  }
  // [][2]int of reference meta/findall.go:155; empty vector == Go's nil
  std::vector<std::pair<int64_t, int64_t>> FindAllIndex(const uint8_t* b, size_t n, int64_t limit = -1) const {
    std::vector<std::pair<int64_t, int64_t>> out;
    if (limit == 0) return out;
    size_t cap = n / 100 + 256, c = 0;
    for (;;) {
      out.resize(cap);
      check(cgx_find_all_index(h_, b, n, limit, reinterpret_cast<int64_t*>(out.data()), cap, &c));
      if (c <= cap) break;
      cap = c;
    }
    out.resize(c);
    return out;
  }
  // flat rows of 2*(NumSubexp()+1) int64, -1 for unmatched groups (reference regex.go:1423)
  std::vector<int64_t> FindAllSubmatchIndex(const uint8_t* b, size_t n, int64_t limit = -1) const {
    std::vector<int64_t> out;
    if (limit == 0) return out;
    const size_t stride = 2 * (size_t)(NumSubexp() + 1);
    size_t cap = n / 64 + 256, c = 0;
    for (;;) {
      out.resize(cap * stride);
      check(cgx_find_all_submatch_index(h_, b, n, limit, out.data(), cap, &c));
      if (c <= cap) break;
      cap = c;
    }
    out.resize(c * stride);
    return out;
  }
  cgx_regex* handle() const { return h_; }

 private:
  Regex(cgx_regex* h, std::string p) : h_(h), pattern_(std::move(p)) {}
  static void check(int rc) {
    if (rc != CGX_OK) throw Error(std::string("coregex_b200: ") + cgx_last_error());
  }
  cgx_regex* h_;
  std::string pattern_;
};

}  // namespace coregex
