// coregex.hpp — header-only C++ mirror of the reference's public Go API for the bulk-scan path
// (reference regex.go: Compile :110, MustCompile :129, CompileWithConfig :198, Match :282,
// FindAllIndex :695, Count :1349, FindAllSubmatchIndex :1423, NumSubexp :552, SubexpNames :575,
// String :444, Longest :464) over the C ABI in coregex_b200.h, plus the wrappers that are host
// plumbing over those batch results (FindIndex :342, Find :307, FindAll :376, ReplaceAllLiteral :790,
// Expand :951, ReplaceAll :1006, ReplaceAllFunc :1136, Split :1288, QuoteMeta :233).
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

  // reference regex.go:198 — config checked field by field (meta/config.go:132-170)
  static Regex CompileWithConfig(const std::string& pattern, const cgx_config& cfg) {
    cgx_regex* h = nullptr;
    char err[1024] = {0};
    if (cgx_compile_cfg(pattern.data(), pattern.size(), &cfg, &h, err, sizeof err) != CGX_OK) throw Error(err);
    return Regex(h, pattern);
  }
  // reference regex.go:464
  void Longest() { check(cgx_set_longest(h_, 1)); }

  const std::string& String() const { return pattern_; }
  int NumSubexp() const { return cgx_num_captures(h_) - 1; }
  std::string Strategy() const { return cgx_strategy(h_); }
  // reference regex.go:575 / :593
  std::vector<std::string> SubexpNames() const {
    std::vector<std::string> names;
    for (int i = 0; i <= NumSubexp(); i++) names.emplace_back(cgx_subexp_name(h_, i));
    return names;
  }
  int SubexpIndex(const std::string& name) const {
    if (name.empty()) return -1;
    for (int i = 0; i <= NumSubexp(); i++)
      if (name == cgx_subexp_name(h_, i)) return i;
    return -1;
  }

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
  // ---- string forms and the wrappers over the batch results ------------------------------------
  using Pairs = std::vector<std::pair<int64_t, int64_t>>;
  static const uint8_t* u8(const std::string& s) { return reinterpret_cast<const uint8_t*>(s.data()); }
  bool MatchString(const std::string& s) const { return Match(u8(s), s.size()); }
  Pairs FindAllStringIndex(const std::string& s, int64_t limit = -1) const { return FindAllIndex(u8(s), s.size(), limit); }
  // reference regex.go:342: {-1,-1} == Go's nil
  std::pair<int64_t, int64_t> FindStringIndex(const std::string& s) const {
    Pairs m = FindAllStringIndex(s, 1);
    return m.empty() ? std::pair<int64_t, int64_t>(-1, -1) : m[0];
  }
  std::string FindString(const std::string& s) const {
    auto m = FindStringIndex(s);
    return m.first < 0 ? std::string() : s.substr((size_t)m.first, (size_t)(m.second - m.first));
  }
  std::vector<std::string> FindAllString(const std::string& s, int64_t limit = -1) const {
    std::vector<std::string> out;
    for (auto& m : FindAllStringIndex(s, limit)) out.push_back(s.substr((size_t)m.first, (size_t)(m.second - m.first)));
    return out;
  }
  // reference regex.go:790
  std::string ReplaceAllLiteralString(const std::string& src, const std::string& repl) const {
    std::string out;
    size_t last = 0;
    for (auto& m : FindAllStringIndex(src)) {
      out.append(src, last, (size_t)m.first - last);
      out += repl;
      last = (size_t)m.second;
    }
    out.append(src, last, std::string::npos);
    return out;
  }
  // reference regex.go:951: $0-$9 (one digit), $$; `${` and unknown escapes keep the `$`
  static void Expand(std::string& dst, const std::string& tmpl, const std::string& src, const int64_t* match, size_t nmatch) {
    for (size_t i = 0; i < tmpl.size();) {
      if (tmpl[i] != '$' || i + 1 >= tmpl.size()) {
        dst += tmpl[i++];
        continue;
      }
      const char nx = tmpl[i + 1];
      if (nx >= '0' && nx <= '9') {
        const size_t g = 2 * (size_t)(nx - '0');
        if (g + 1 < nmatch && match[g] >= 0) dst.append(src, (size_t)match[g], (size_t)(match[g + 1] - match[g]));
        i += 2;
      } else if (nx == '$') {
        dst += '$';
        i += 2;
      } else {
        dst += '$';
        i++;
      }
    }
  }
  // reference regex.go:1006
  std::string ReplaceAllString(const std::string& src, const std::string& repl) const {
    if (repl.find('$') == std::string::npos) return ReplaceAllLiteralString(src, repl);
    const size_t stride = 2 * (size_t)(NumSubexp() + 1);
    const std::vector<int64_t> rows = FindAllSubmatchIndex(u8(src), src.size());
    std::string out;
    size_t last = 0;
    for (size_t k = 0; k + stride <= rows.size(); k += stride) {
      out.append(src, last, (size_t)rows[k] - last);
      Expand(out, repl, src, &rows[k], stride);
      last = (size_t)rows[k + 1];
    }
    out.append(src, last, std::string::npos);
    return out;
  }
  // reference regex.go:1209
  template <class F>
  std::string ReplaceAllStringFunc(const std::string& src, F fn) const {
    std::string out;
    size_t last = 0;
    for (auto& m : FindAllStringIndex(src)) {
      out.append(src, last, (size_t)m.first - last);
      out += fn(src.substr((size_t)m.first, (size_t)(m.second - m.first)));
      last = (size_t)m.second;
    }
    out.append(src, last, std::string::npos);
    return out;
  }
  // reference regex.go:1288 (n == 0: empty vector, Go's nil)
  std::vector<std::string> Split(const std::string& s, int n = -1) const {
    std::vector<std::string> out;
    if (n == 0) return out;
    const Pairs idx = FindAllStringIndex(s);
    if (idx.empty()) return {s};
    size_t last = 0;
    for (auto& m : idx) {
      if (last == 0 && m.first == 0 && m.second == 0) continue;
      if ((size_t)m.first == s.size() && (size_t)m.second == s.size()) break;
      out.push_back(s.substr(last, (size_t)m.first - last));
      last = (size_t)m.second;
      if (n > 0 && (int)out.size() >= n - 1) {
        out.push_back(s.substr(last));
        return out;
      }
    }
    out.push_back(s.substr(last));
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

// reference regex.go:233
inline std::string QuoteMeta(const std::string& s) {
  static const std::string special = "\\.+*?()|[]{}^$";
  std::string out;
  for (char c : s) {
    if (special.find(c) != std::string::npos) out += '\\';
    out += c;
  }
  return out;
}

}  // namespace coregex
