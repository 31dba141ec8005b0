// parse.cpp — restatement of Go regexp/syntax/parse.go (see syntax.h for scope and limits).
#include "syntax.h"

#include <algorithm>
#include <cstring>

namespace gosyntax {

namespace {

struct Error {
  std::string code, expr;
};

const char* ErrInternalError = "regexp/syntax: internal error";
const char* ErrInvalidCharClass = "invalid character class";
const char* ErrInvalidCharRange = "invalid character class range";
const char* ErrInvalidEscape = "invalid escape sequence";
const char* ErrInvalidNamedCapture = "invalid named capture";
const char* ErrInvalidPerlOp = "invalid or unsupported Perl syntax";
const char* ErrInvalidRepeatOp = "invalid nested repetition operator";
const char* ErrInvalidRepeatSize = "invalid repeat count";
const char* ErrInvalidUTF8 = "invalid UTF-8";
const char* ErrMissingBracket = "missing closing ]";
const char* ErrMissingParen = "missing closing )";
const char* ErrMissingRepeatArgument = "missing argument to repetition operator";
const char* ErrTrailingBackslash = "trailing backslash at end of expression";
const char* ErrUnexpectedParen = "unexpected )";
const char* ErrUnsupportedUnicode = "unsupported: Unicode class";

using Runes = std::vector<int32_t>;
using sv = std::string;  // we pass (string, offset) pairs as std::string substrings for clarity

// ---- Unicode data: \p{..} range tables and simple-case-folding orbits ------------------------
// The reference parses with Go's regexp/syntax (standard library, not in the reference tree), which
// reads unicode.Categories / unicode.Scripts / unicode.SimpleFold (Unicode 15.0.0).  The tables here
// are regenerated from an independent 15.0.0 database by tools/gen_unicode_tables.py.
struct UnicodeTable {
  const char* name;
  const int32_t* r;  // inclusive ranges lo,hi,...
  int n;             // number of int32 in r
  int is_script;
  const int32_t* fold;  // runes outside the table whose case orbit reaches into it (may be null)
  int nfold;
};
struct UnicodeAlias {
  const char* alias;
  const char* name;
};
#include "unicode_tables.inc"

// unicode.SimpleFold: the next larger rune of r's case orbit, wrapping around to the smallest
int32_t simpleFold(int32_t r) {
  size_t lo = 0, hi = sizeof(kCaseOrbit) / sizeof(kCaseOrbit[0]);
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (kCaseOrbit[mid][0] < r) lo = mid + 1; else hi = mid;
  }
  if (lo < sizeof(kCaseOrbit) / sizeof(kCaseOrbit[0]) && kCaseOrbit[lo][0] == r) return kCaseOrbit[lo][1];
  return r;
}
const int32_t minFold = 0x0041, maxFold = 0x1e943;

int32_t minFoldRune(int32_t r) {
  if (r < minFold || r > maxFold) return r;
  int32_t m = r, r0 = r;
  for (r = simpleFold(r); r != r0; r = simpleFold(r)) m = std::min(m, r);
  return m;
}

// ---- class helpers (appendRange / cleanClass / negateClass) ---------------------------------
void appendRange(Runes& r, int32_t lo, int32_t hi) {
  size_t n = r.size();
  for (size_t i = 2; i <= 4; i += 2) {
    if (n >= i) {
      int32_t rlo = r[n - i], rhi = r[n - i + 1];
      if (lo <= rhi + 1 && rlo <= hi + 1) {
        if (lo < rlo) r[n - i] = lo;
        if (hi > rhi) r[n - i + 1] = hi;
        return;
      }
    }
  }
  r.push_back(lo);
  r.push_back(hi);
}

void appendFoldedRange(Runes& r, int32_t lo, int32_t hi) {
  if (lo <= minFold && hi >= maxFold) return appendRange(r, lo, hi);
  if (hi < minFold || lo > maxFold) return appendRange(r, lo, hi);
  if (lo < minFold) {
    appendRange(r, lo, minFold - 1);
    lo = minFold;
  }
  if (hi > maxFold) {
    appendRange(r, maxFold + 1, hi);
    hi = maxFold;
  }
  // brute force; appendRange coalesces on the fly
  for (int32_t c = lo; c <= hi; c++) {
    appendRange(r, c, c);
    for (int32_t f = simpleFold(c); f != c; f = simpleFold(f)) appendRange(r, f, f);
  }
}

void appendLiteral(Runes& r, int32_t x, uint16_t flags) {
  if (flags & FoldCase)
    appendFoldedRange(r, x, x);
  else
    appendRange(r, x, x);
}

void appendClass(Runes& r, const Runes& x) {
  for (size_t i = 0; i + 1 < x.size(); i += 2) appendRange(r, x[i], x[i + 1]);
}

void appendFoldedClass(Runes& r, const Runes& x) {
  for (size_t i = 0; i + 1 < x.size(); i += 2) appendFoldedRange(r, x[i], x[i + 1]);
}

void appendNegatedClass(Runes& r, const Runes& x) {
  int32_t nextLo = 0;
  for (size_t i = 0; i + 1 < x.size(); i += 2) {
    int32_t lo = x[i], hi = x[i + 1];
    if (nextLo <= lo - 1) appendRange(r, nextLo, lo - 1);
    nextLo = hi + 1;
  }
  if (nextLo <= kMaxRune) appendRange(r, nextLo, kMaxRune);
}

void cleanClass(Runes& r) {
  // sort by lo increasing, hi decreasing to break ties
  size_t n = r.size() / 2;
  std::vector<std::pair<int32_t, int32_t>> v(n);
  for (size_t i = 0; i < n; i++) v[i] = {r[2 * i], r[2 * i + 1]};
  std::sort(v.begin(), v.end(), [](auto& a, auto& b) {
    return a.first < b.first || (a.first == b.first && a.second > b.second);
  });
  Runes out;
  for (auto& p : v) {
    if (!out.empty() && p.first <= out.back() + 1) {
      if (p.second > out.back()) out.back() = p.second;
      continue;
    }
    out.push_back(p.first);
    out.push_back(p.second);
  }
  r.swap(out);
}

void negateClass(Runes& r) {
  Runes out;
  int32_t nextLo = 0;
  for (size_t i = 0; i + 1 < r.size(); i += 2) {
    int32_t lo = r[i], hi = r[i + 1];
    if (nextLo <= lo - 1) {
      out.push_back(nextLo);
      out.push_back(lo - 1);
    }
    nextLo = hi + 1;
  }
  if (nextLo <= kMaxRune) {
    out.push_back(nextLo);
    out.push_back(kMaxRune);
  }
  r.swap(out);
}

struct CharGroup {
  int sign;
  Runes cls;
};

// Go 1.25 regexp/syntax unicodeTable: names compare case-insensitively with spaces, underscores and
// hyphens ignored; Any / ASCII / Assigned are built in, categories answer to their aliases too
std::string canonicalUnicodeName(const std::string& name) {
  std::string out;
  for (char c : name) {
    if (c == ' ' || c == '_' || c == '-') continue;
    out.push_back((c >= 'A' && c <= 'Z') ? (char)(c + 32) : c);
  }
  return out;
}

// false: no such table.  sign is flipped for names defined as a complement (Assigned = not Cn)
bool unicodeTable(const std::string& name, Runes& tab, Runes& fold, int& sign) {
  const std::string want = canonicalUnicodeName(name);
  if (want.empty()) return false;
  if (want == "any") {
    tab = {0, kMaxRune};
    return true;
  }
  if (want == "ascii") {
    tab = {0, 0x7F};
    return true;
  }
  std::string key = want == "assigned" ? "cn" : want;
  if (want == "assigned") sign = -sign;
  for (const UnicodeAlias& a : kUnicodeAliases)
    if (canonicalUnicodeName(a.alias) == key) key = canonicalUnicodeName(a.name);
  for (int pass = 0; pass < 2; pass++)  // categories first, then scripts
    for (const UnicodeTable& t : kUnicodeTables)
      if (t.is_script == pass && canonicalUnicodeName(t.name) == key) {
        tab.assign(t.r, t.r + t.n);
        fold.assign(t.fold, t.fold + t.nfold);
        return true;
      }
  return false;
}

const Runes code_d = {'0', '9'};
const Runes code_s = {0x9, 0xa, 0xc, 0xd, 0x20, 0x20};
const Runes code_w = {'0', '9', 'A', 'Z', '_', '_', 'a', 'z'};

bool perlGroup(const std::string& name, CharGroup& g) {
  if (name == "\\d") g = {+1, code_d};
  else if (name == "\\D") g = {-1, code_d};
  else if (name == "\\s") g = {+1, code_s};
  else if (name == "\\S") g = {-1, code_s};
  else if (name == "\\w") g = {+1, code_w};
  else if (name == "\\W") g = {-1, code_w};
  else return false;
  return true;
}

bool posixGroup(const std::string& name, CharGroup& g) {
  struct E {
    const char* n;
    Runes c;
  };
  static const E tab[] = {
      {"alnum", {'0', '9', 'A', 'Z', 'a', 'z'}},
      {"alpha", {'A', 'Z', 'a', 'z'}},
      {"ascii", {0x0, 0x7f}},
      {"blank", {'\t', '\t', ' ', ' '}},
      {"cntrl", {0x0, 0x1f, 0x7f, 0x7f}},
      {"digit", {'0', '9'}},
      {"graph", {'!', '~'}},
      {"lower", {'a', 'z'}},
      {"print", {' ', '~'}},
      {"punct", {'!', '/', ':', '@', '[', '`', '{', '~'}},
      {"space", {'\t', '\r', ' ', ' '}},
      {"upper", {'A', 'Z'}},
      {"word", {'0', '9', 'A', 'Z', '_', '_', 'a', 'z'}},
      {"xdigit", {'0', '9', 'A', 'F', 'a', 'f'}},
  };
  // name is "[:alpha:]" or "[:^alpha:]"
  if (name.size() < 4) return false;
  std::string inner = name.substr(2, name.size() - 4);
  int sign = +1;
  if (!inner.empty() && inner[0] == '^') {
    sign = -1;
    inner = inner.substr(1);
  }
  for (auto& e : tab)
    if (inner == e.n) {
      g = {sign, e.c};
      return true;
    }
  return false;
}

bool isalnum_(int32_t c) {
  return ('0' <= c && c <= '9') || ('A' <= c && c <= 'Z') || ('a' <= c && c <= 'z');
}
int unhex(int32_t c) {
  if ('0' <= c && c <= '9') return c - '0';
  if ('a' <= c && c <= 'f') return c - 'a' + 10;
  if ('A' <= c && c <= 'F') return c - 'A' + 10;
  return -1;
}

bool isCharClass(const Regexp* re) {
  return (re->op == OpLiteral && re->rune.size() == 1) || re->op == OpCharClass ||
         re->op == OpAnyCharNotNL || re->op == OpAnyChar;
}

bool matchRune(const Regexp* re, int32_t r) {
  switch (re->op) {
    case OpLiteral:
      return re->rune.size() == 1 && re->rune[0] == r;
    case OpCharClass:
      for (size_t i = 0; i + 1 < re->rune.size(); i += 2)
        if (re->rune[i] <= r && r <= re->rune[i + 1]) return true;
      return false;
    case OpAnyCharNotNL:
      return r != '\n';
    case OpAnyChar:
      return true;
    default:
      return false;
  }
}

void mergeCharClass(Regexp* dst, Regexp* src) {
  switch (dst->op) {
    case OpAnyChar:
      break;
    case OpAnyCharNotNL:
      if (matchRune(src, '\n')) dst->op = OpAnyChar;
      break;
    case OpCharClass:
      if (src->op == OpLiteral)
        appendLiteral(dst->rune, src->rune[0], src->flags);
      else
        appendClass(dst->rune, src->rune);
      break;
    case OpLiteral: {
      if (src->rune[0] == dst->rune[0] && src->flags == dst->flags) break;
      dst->op = OpCharClass;
      int32_t d0 = dst->rune[0];
      dst->rune.clear();
      appendLiteral(dst->rune, d0, dst->flags);
      appendLiteral(dst->rune, src->rune[0], src->flags);
      break;
    }
    default:
      break;
  }
}

void cleanAlt(Regexp* re) {
  if (re->op != OpCharClass) return;
  cleanClass(re->rune);
  if (re->rune.size() == 2 && re->rune[0] == 0 && re->rune[1] == kMaxRune) {
    re->rune.clear();
    re->op = OpAnyChar;
    return;
  }
  if (re->rune.size() == 4 && re->rune[0] == 0 && re->rune[1] == '\n' - 1 &&
      re->rune[2] == '\n' + 1 && re->rune[3] == kMaxRune) {
    re->rune.clear();
    re->op = OpAnyCharNotNL;
    return;
  }
}

// repeatIsValid: nested repeat product must stay <= n
bool repeatIsValid(const Regexp* re, int n) {
  if (re->op == OpRepeat) {
    int m = re->max;
    if (m == 0) return true;
    if (m < 0) m = re->min;
    if (m > n) return false;
    if (m > 0) n /= m;
  }
  for (auto* s : re->sub)
    if (!repeatIsValid(s, n)) return false;
  return true;
}

struct Parser {
  Arena& arena;
  uint16_t flags;
  std::vector<Regexp*> stack;
  int numCap = 0;
  std::string whole;

  explicit Parser(Arena& a, uint16_t f) : arena(a), flags(f) {}

  Regexp* newRegexp(Op op) { return arena.make(op); }

  // -- UTF-8 decoding of the pattern -----------------------------------------------------------
  // returns false on invalid UTF-8
  static bool nextRune(const std::string& s, size_t& i, int32_t& c) {
    unsigned char b0 = s[i];
    if (b0 < 0x80) {
      c = b0;
      i += 1;
      return true;
    }
    int need = 0;
    int32_t v = 0;
    if (b0 >= 0xC2 && b0 <= 0xDF) {
      need = 1;
      v = b0 & 0x1F;
    } else if (b0 >= 0xE0 && b0 <= 0xEF) {
      need = 2;
      v = b0 & 0x0F;
    } else if (b0 >= 0xF0 && b0 <= 0xF4) {
      need = 3;
      v = b0 & 0x07;
    } else
      return false;
    if (i + (size_t)need >= s.size()) return false;
    for (int k = 1; k <= need; k++) {
      unsigned char b = s[i + k];
      if ((b & 0xC0) != 0x80) return false;
      v = (v << 6) | (b & 0x3F);
    }
    if ((need == 2 && v < 0x800) || (need == 3 && v < 0x10000) || v > kMaxRune ||
        (v >= 0xD800 && v <= 0xDFFF))
      return false;
    c = v;
    i += need + 1;
    return true;
  }

  // -- stack machinery ------------------------------------------------------------------------
  bool maybeConcat(int32_t r, uint16_t fl) {
    size_t n = stack.size();
    if (n < 2) return false;
    Regexp* re1 = stack[n - 1];
    Regexp* re2 = stack[n - 2];
    if (re1->op != OpLiteral || re2->op != OpLiteral ||
        (re1->flags & FoldCase) != (re2->flags & FoldCase))
      return false;
    re2->rune.insert(re2->rune.end(), re1->rune.begin(), re1->rune.end());
    if (r >= 0) {
      re1->rune.assign(1, r);
      re1->flags = fl;
      return true;
    }
    stack.pop_back();
    return false;
  }

  Regexp* push(Regexp* re) {
    if (re->op == OpCharClass && re->rune.size() == 2 && re->rune[0] == re->rune[1]) {
      if (maybeConcat(re->rune[0], flags & ~FoldCase)) return nullptr;
      re->op = OpLiteral;
      re->rune.resize(1);
      re->flags = flags & ~FoldCase;
    } else if ((re->op == OpCharClass && re->rune.size() == 4 && re->rune[0] == re->rune[1] &&
                re->rune[2] == re->rune[3] && simpleFold(re->rune[0]) == re->rune[2] &&
                simpleFold(re->rune[2]) == re->rune[0]) ||
               (re->op == OpCharClass && re->rune.size() == 2 && re->rune[0] + 1 == re->rune[1] &&
                simpleFold(re->rune[0]) == re->rune[1] && simpleFold(re->rune[1]) == re->rune[0])) {
      if (maybeConcat(re->rune[0], flags | FoldCase)) return nullptr;
      re->op = OpLiteral;
      re->rune.resize(1);
      re->flags = flags | FoldCase;
    } else {
      maybeConcat(-1, 0);
    }
    stack.push_back(re);
    return re;
  }

  void literal(int32_t r) {
    Regexp* re = newRegexp(OpLiteral);
    re->flags = flags;
    if (flags & FoldCase) r = minFoldRune(r);
    re->rune.assign(1, r);
    push(re);
  }

  Regexp* op(Op o) {
    Regexp* re = newRegexp(o);
    re->flags = flags;
    return push(re);
  }

  bool repeat(Op o, int min, int max, const std::string& t, size_t before, size_t& after,
              size_t lastRepeat /* npos if none */, Error& err) {
    uint16_t fl = flags;
    if (flags & PerlX) {
      if (after < t.size() && t[after] == '?') {
        after++;
        fl ^= NonGreedy;
      }
      if (lastRepeat != std::string::npos) {
        err = {ErrInvalidRepeatOp, t.substr(lastRepeat, after - lastRepeat)};
        return false;
      }
    }
    size_t n = stack.size();
    if (n == 0) {
      err = {ErrMissingRepeatArgument, t.substr(before, after - before)};
      return false;
    }
    Regexp* sub = stack[n - 1];
    if (sub->op >= opPseudo) {
      err = {ErrMissingRepeatArgument, t.substr(before, after - before)};
      return false;
    }
    Regexp* re = newRegexp(o);
    re->min = min;
    re->max = max;
    re->flags = fl;
    re->sub.assign(1, sub);
    stack[n - 1] = re;
    if (o == OpRepeat && (min >= 2 || max >= 2) && !repeatIsValid(re, 1000)) {
      err = {ErrInvalidRepeatSize, t.substr(before, after - before)};
      return false;
    }
    return true;
  }

  Regexp* collapse(std::vector<Regexp*> subs, Op o) {
    if (subs.size() == 1) return subs[0];
    Regexp* re = newRegexp(o);
    for (auto* s : subs) {
      if (s->op == o)
        re->sub.insert(re->sub.end(), s->sub.begin(), s->sub.end());
      else
        re->sub.push_back(s);
    }
    if (o == OpAlternate) {
      re->sub = factor(re->sub);
      if (re->sub.size() == 1) re = re->sub[0];
    }
    return re;
  }

  Regexp* concat() {
    maybeConcat(-1, 0);
    size_t i = stack.size();
    while (i > 0 && stack[i - 1]->op < opPseudo) i--;
    std::vector<Regexp*> subs(stack.begin() + i, stack.end());
    stack.resize(i);
    if (subs.empty()) return push(newRegexp(OpEmptyMatch));
    return push(collapse(subs, OpConcat));
  }

  Regexp* alternate() {
    size_t i = stack.size();
    while (i > 0 && stack[i - 1]->op < opPseudo) i--;
    std::vector<Regexp*> subs(stack.begin() + i, stack.end());
    stack.resize(i);
    if (!subs.empty()) cleanAlt(subs.back());
    if (subs.empty()) return push(newRegexp(OpNoMatch));
    return push(collapse(subs, OpAlternate));
  }

  bool swapVerticalBar() {
    size_t n = stack.size();
    if (n >= 3 && stack[n - 2]->op == opVerticalBar && isCharClass(stack[n - 1]) &&
        isCharClass(stack[n - 3])) {
      Regexp* re1 = stack[n - 1];
      Regexp* re3 = stack[n - 3];
      if (re1->op > re3->op) {
        std::swap(re1, re3);
        stack[n - 3] = re3;
      }
      mergeCharClass(re3, re1);
      stack.pop_back();
      return true;
    }
    if (n >= 2) {
      Regexp* re1 = stack[n - 1];
      Regexp* re2 = stack[n - 2];
      if (re2->op == opVerticalBar) {
        if (n >= 3) cleanAlt(stack[n - 3]);
        stack[n - 2] = re1;
        stack[n - 1] = re2;
        return true;
      }
    }
    return false;
  }

  void parseVerticalBar() {
    concat();
    if (!swapVerticalBar()) op(opVerticalBar);
  }

  bool parseRightParen(Error& err) {
    concat();
    if (swapVerticalBar()) stack.pop_back();
    alternate();
    size_t n = stack.size();
    if (n < 2) {
      err = {ErrUnexpectedParen, whole};
      return false;
    }
    Regexp* re1 = stack[n - 1];
    Regexp* re2 = stack[n - 2];
    stack.resize(n - 2);
    if (re2->op != opLeftParen) {
      err = {ErrUnexpectedParen, whole};
      return false;
    }
    flags = re2->flags;
    if (re2->cap == 0) {
      push(re1);
    } else {
      re2->op = OpCapture;
      re2->sub.assign(1, re1);
      push(re2);
    }
    return true;
  }

  // -- factor(): alternation factoring, rounds 1..4 -------------------------------------------
  static void leadingString(Regexp* re, const Runes*& str, uint16_t& fl) {
    if (re->op == OpConcat && !re->sub.empty()) re = re->sub[0];
    if (re->op != OpLiteral) {
      str = nullptr;
      fl = 0;
      return;
    }
    str = &re->rune;
    fl = re->flags & FoldCase;
  }

  Regexp* removeLeadingString(Regexp* re, size_t n) {
    if (re->op == OpConcat && !re->sub.empty()) {
      Regexp* sub = removeLeadingString(re->sub[0], n);
      re->sub[0] = sub;
      if (sub->op == OpEmptyMatch) {
        switch (re->sub.size()) {
          case 0:
          case 1:
            re->op = OpEmptyMatch;
            re->sub.clear();
            break;
          case 2:
            re = re->sub[1];
            break;
          default:
            re->sub.erase(re->sub.begin());
        }
      }
      return re;
    }
    if (re->op == OpLiteral) {
      re->rune.erase(re->rune.begin(), re->rune.begin() + n);
      if (re->rune.empty()) re->op = OpEmptyMatch;
    }
    return re;
  }

  static Regexp* leadingRegexp(Regexp* re) {
    if (re->op == OpEmptyMatch) return nullptr;
    if (re->op == OpConcat && !re->sub.empty()) {
      Regexp* sub = re->sub[0];
      if (sub->op == OpEmptyMatch) return nullptr;
      return sub;
    }
    return re;
  }

  Regexp* removeLeadingRegexp(Regexp* re) {
    if (re->op == OpConcat && !re->sub.empty()) {
      re->sub.erase(re->sub.begin());
      switch (re->sub.size()) {
        case 0:
          re->op = OpEmptyMatch;
          re->sub.clear();
          break;
        case 1:
          re = re->sub[0];
          break;
      }
      return re;
    }
    return newRegexp(OpEmptyMatch);
  }

  std::vector<Regexp*> factor(std::vector<Regexp*> sub) {
    if (sub.size() < 2) return sub;

    // Round 1: factor out common literal prefixes.
    {
      Runes str;
      uint16_t strflags = 0;
      size_t start = 0;
      std::vector<Regexp*> out;
      for (size_t i = 0; i <= sub.size(); i++) {
        const Runes* istr = nullptr;
        uint16_t iflags = 0;
        if (i < sub.size()) {
          leadingString(sub[i], istr, iflags);
          if (iflags == strflags) {
            size_t same = 0;
            size_t ilen = istr ? istr->size() : 0;
            while (same < str.size() && same < ilen && str[same] == (*istr)[same]) same++;
            if (same > 0) {
              str.resize(same);
              continue;
            }
          }
        }
        if (i == start) {
        } else if (i == start + 1) {
          out.push_back(sub[start]);
        } else {
          Regexp* prefix = newRegexp(OpLiteral);
          prefix->flags = strflags;
          prefix->rune = str;
          for (size_t j = start; j < i; j++) sub[j] = removeLeadingString(sub[j], str.size());
          Regexp* suffix =
              collapse(std::vector<Regexp*>(sub.begin() + start, sub.begin() + i), OpAlternate);
          Regexp* re = newRegexp(OpConcat);
          re->sub = {prefix, suffix};
          out.push_back(re);
        }
        start = i;
        str = istr ? *istr : Runes();
        strflags = iflags;
      }
      sub.swap(out);
    }

    // Round 2: factor out common simple prefixes (first piece of each concatenation).
    {
      size_t start = 0;
      std::vector<Regexp*> out;
      Regexp* first = nullptr;
      for (size_t i = 0; i <= sub.size(); i++) {
        Regexp* ifirst = nullptr;
        if (i < sub.size()) {
          ifirst = leadingRegexp(sub[i]);
          if (first != nullptr && ifirst != nullptr && first->equal(ifirst) &&
              (isCharClass(first) || (first->op == OpRepeat && first->min == first->max &&
                                      isCharClass(first->sub[0]))))
            continue;
        }
        if (i == start) {
        } else if (i == start + 1) {
          out.push_back(sub[start]);
        } else {
          Regexp* prefix = first;
          for (size_t j = start; j < i; j++) sub[j] = removeLeadingRegexp(sub[j]);
          Regexp* suffix =
              collapse(std::vector<Regexp*>(sub.begin() + start, sub.begin() + i), OpAlternate);
          Regexp* re = newRegexp(OpConcat);
          re->sub = {prefix, suffix};
          out.push_back(re);
        }
        start = i;
        first = ifirst;
      }
      sub.swap(out);
    }

    // Round 3: collapse runs of single literals or character classes.
    {
      size_t start = 0;
      std::vector<Regexp*> out;
      for (size_t i = 0; i <= sub.size(); i++) {
        if (i < sub.size() && isCharClass(sub[i])) continue;
        if (i == start) {
        } else if (i == start + 1) {
          out.push_back(sub[start]);
        } else {
          size_t mx = start;
          for (size_t j = start + 1; j < i; j++)
            if (sub[mx]->op < sub[j]->op ||
                (sub[mx]->op == sub[j]->op && sub[mx]->rune.size() < sub[j]->rune.size()))
              mx = j;
          std::swap(sub[start], sub[mx]);
          for (size_t j = start + 1; j < i; j++) mergeCharClass(sub[start], sub[j]);
          cleanAlt(sub[start]);
          out.push_back(sub[start]);
        }
        if (i < sub.size()) out.push_back(sub[i]);
        start = i + 1;
      }
      sub.swap(out);
    }

    // Round 4: collapse runs of empty matches into a single empty match.
    {
      std::vector<Regexp*> out;
      for (size_t i = 0; i < sub.size(); i++) {
        if (i + 1 < sub.size() && sub[i]->op == OpEmptyMatch && sub[i + 1]->op == OpEmptyMatch)
          continue;
        out.push_back(sub[i]);
      }
      sub.swap(out);
    }
    return sub;
  }

  // -- lexing helpers -------------------------------------------------------------------------
  // parseInt at t[i...]; returns -1 if none; caps at "huge" like Go (>= 1e8 -> -1... invalid size)
  static int parseInt(const std::string& t, size_t& i) {
    if (i >= t.size() || t[i] < '0' || t[i] > '9') return -1;
    // disallow leading zeros
    if (i + 1 < t.size() && t[i] == '0' && t[i + 1] >= '0' && t[i + 1] <= '9') return -1;
    size_t j = i;
    while (j < t.size() && t[j] >= '0' && t[j] <= '9') j++;
    int n = 0;
    for (size_t k = i; k < j; k++) {
      if (n >= 100000000) {
        n = -2;  // overflow marker: treated as "too big" by the caller
        break;
      }
      n = n * 10 + (t[k] - '0');
    }
    i = j;
    return n;
  }

  // {n}, {n,}, {n,m}; returns false if not a repeat (brace is then literal)
  static bool parseRepeat(const std::string& t, size_t i, int& min, int& max, size_t& rest) {
    if (i >= t.size() || t[i] != '{') return false;
    i++;
    min = parseInt(t, i);
    if (min == -1) return false;
    if (i >= t.size()) return false;
    if (t[i] != ',') {
      max = min;
    } else {
      i++;
      if (i >= t.size()) return false;
      if (t[i] == '}') {
        max = -1;
      } else {
        max = parseInt(t, i);
        if (max == -1) return false;
        if (max == -2) min = -2;  // propagate "too big"
      }
    }
    if (i >= t.size() || t[i] != '}') return false;
    rest = i + 1;
    if (min == -2) min = 1001;  // force ErrInvalidRepeatSize upstream
    return true;
  }

  bool parseEscape(const std::string& t, size_t& i, int32_t& r, Error& err) {
    size_t s0 = i;
    i++;  // backslash
    if (i >= t.size()) {
      err = {ErrTrailingBackslash, ""};
      return false;
    }
    int32_t c;
    if (!nextRune(t, i, c)) {
      err = {ErrInvalidUTF8, t.substr(s0)};
      return false;
    }
    switch (c) {
      default:
        if (c < 0x80 && !isalnum_(c)) {
          r = c;
          return true;
        }
        break;
      case '1': case '2': case '3': case '4': case '5': case '6': case '7':
        if (i >= t.size() || t[i] < '0' || t[i] > '7') break;
        [[fallthrough]];
      case '0': {
        r = c - '0';
        for (int k = 1; k < 3; k++) {
          if (i >= t.size() || t[i] < '0' || t[i] > '7') break;
          r = r * 8 + (t[i] - '0');
          i++;
        }
        return true;
      }
      case 'x': {
        if (i >= t.size()) break;
        if (!nextRune(t, i, c)) {
          err = {ErrInvalidUTF8, t.substr(s0)};
          return false;
        }
        if (c == '{') {
          int nhex = 0;
          r = 0;
          bool bad = false;
          for (;;) {
            if (i >= t.size()) { bad = true; break; }
            if (!nextRune(t, i, c)) {
              err = {ErrInvalidUTF8, t.substr(s0)};
              return false;
            }
            if (c == '}') break;
            int v = unhex(c);
            if (v < 0) { bad = true; break; }
            r = r * 16 + v;
            if (r > kMaxRune) { bad = true; break; }
            nhex++;
          }
          if (bad || nhex == 0) break;
          return true;
        }
        int x = unhex(c);
        if (i >= t.size()) break;
        if (!nextRune(t, i, c)) {
          err = {ErrInvalidUTF8, t.substr(s0)};
          return false;
        }
        int y = unhex(c);
        if (x < 0 || y < 0) break;
        r = x * 16 + y;
        return true;
      }
      case 'a': r = 7; return true;
      case 'f': r = '\f'; return true;
      case 'n': r = '\n'; return true;
      case 'r': r = '\r'; return true;
      case 't': r = '\t'; return true;
      case 'v': r = '\v'; return true;
    }
    err = {ErrInvalidEscape, t.substr(s0, i - s0)};
    return false;
  }

  void appendGroup(Runes& r, const CharGroup& g) {
    if (!(flags & FoldCase)) {
      if (g.sign < 0)
        appendNegatedClass(r, g.cls);
      else
        appendClass(r, g.cls);
    } else {
      Runes tmp;
      appendFoldedClass(tmp, g.cls);
      cleanClass(tmp);
      if (g.sign < 0)
        appendNegatedClass(r, tmp);
      else
        appendClass(r, tmp);
    }
  }

  // \p{Name} \pN \P{Name} \p{^Name}: 0 = not one, 1 = consumed, -1 = error
  int parseUnicodeClass(const std::string& t, size_t& i, Runes& r, Error& err) {
    if (!(flags & UnicodeGroups) || i + 2 > t.size() || t[i] != '\\' || (t[i + 1] != 'p' && t[i + 1] != 'P')) return 0;
    int sign = t[i + 1] == 'P' ? -1 : +1;
    size_t k = i + 2;
    std::string seq, name;
    if (k < t.size() && t[k] == '{') {
      size_t end = t.find('}', i);
      if (end == std::string::npos) {
        err = {ErrInvalidCharRange, t.substr(i)};
        return -1;
      }
      seq = t.substr(i, end + 1 - i);
      name = t.substr(i + 3, end - (i + 3));
      k = end + 1;
    } else {
      int32_t c = 0;
      if (k < t.size() && !nextRune(t, k, c)) {
        err = {ErrInvalidUTF8, t.substr(k)};
        return -1;
      }
      seq = t.substr(i, k - i);
      name = seq.substr(2);
    }
    if (!name.empty() && name[0] == '^') {
      sign = -sign;
      name = name.substr(1);
    }
    Runes tab, fold;
    if (!unicodeTable(name, tab, fold, sign)) {
      err = {ErrInvalidCharRange, seq};
      return -1;
    }
    if ((flags & FoldCase) && !fold.empty()) {
      tab.insert(tab.end(), fold.begin(), fold.end());
      cleanClass(tab);
    }
    if (sign > 0)
      appendClass(r, tab);
    else
      appendNegatedClass(r, tab);
    i = k;
    return 1;
  }

  // returns true if a Perl class escape was consumed
  bool parsePerlClassEscape(const std::string& t, size_t& i, Runes& r) {
    if (!(flags & PerlX) || i + 2 > t.size() || t[i] != '\\') return false;
    CharGroup g;
    if (!perlGroup(t.substr(i, 2), g)) return false;
    appendGroup(r, g);
    i += 2;
    return true;
  }

  // 0 = not a named class, 1 = consumed, -1 = error
  int parseNamedClass(const std::string& t, size_t& i, Runes& r, Error& err) {
    if (i + 2 > t.size() || t[i] != '[' || t[i + 1] != ':') return 0;
    size_t e = t.find(":]", i + 2);
    if (e == std::string::npos) return 0;
    std::string name = t.substr(i, e + 2 - i);
    CharGroup g;
    if (!posixGroup(name, g)) {
      err = {ErrInvalidCharRange, name};
      return -1;
    }
    appendGroup(r, g);
    i = e + 2;
    return 1;
  }

  bool parseClassChar(const std::string& t, size_t& i, size_t wholeStart, int32_t& r, Error& err) {
    if (i >= t.size()) {
      err = {ErrMissingBracket, t.substr(wholeStart)};
      return false;
    }
    if (t[i] == '\\') return parseEscape(t, i, r, err);
    if (!nextRune(t, i, r)) {
      err = {ErrInvalidUTF8, t.substr(i)};
      return false;
    }
    return true;
  }

  bool parseClass(const std::string& t, size_t& i, Error& err) {
    size_t s0 = i;
    i++;  // chop [
    Regexp* re = newRegexp(OpCharClass);
    re->flags = flags;
    int sign = +1;
    if (i < t.size() && t[i] == '^') {
      sign = -1;
      i++;
      if (!(flags & ClassNL)) {
        re->rune.push_back('\n');
        re->rune.push_back('\n');
      }
    }
    Runes& cls = re->rune;
    bool first = true;
    while (i >= t.size() || t[i] != ']' || first) {
      if (i < t.size() && t[i] == '-' && !(flags & PerlX) && !first &&
          (i + 1 == t.size() || t[i + 1] != ']')) {
        err = {ErrInvalidCharRange, t.substr(i, 1)};
        return false;
      }
      first = false;
      if (i + 2 < t.size() && t[i] == '[' && t[i + 1] == ':') {
        int k = parseNamedClass(t, i, cls, err);
        if (k < 0) return false;
        if (k > 0) continue;
      }
      {
        int k = parseUnicodeClass(t, i, cls, err);
        if (k < 0) return false;
        if (k > 0) continue;
      }
      if (parsePerlClassEscape(t, i, cls)) continue;
      size_t rng = i;
      int32_t lo, hi;
      if (!parseClassChar(t, i, s0, lo, err)) return false;
      hi = lo;
      if (i + 1 < t.size() && t[i] == '-' && t[i + 1] != ']') {
        i++;
        if (!parseClassChar(t, i, s0, hi, err)) return false;
        if (hi < lo) {
          err = {ErrInvalidCharRange, t.substr(rng, i - rng)};
          return false;
        }
      }
      if (!(flags & FoldCase))
        appendRange(cls, lo, hi);
      else
        appendFoldedRange(cls, lo, hi);
    }
    i++;  // chop ]
    cleanClass(cls);
    if (sign < 0) negateClass(cls);
    push(re);
    return true;
  }

  static bool isValidCaptureName(const std::string& name) {
    if (name.empty()) return false;
    for (unsigned char c : name)
      if (c != '_' && !isalnum_(c)) return false;
    return true;
  }

  bool parsePerlFlags(const std::string& t, size_t& i, Error& err) {
    size_t s0 = i;
    bool startsWithP = t.size() - i > 4 && t[i + 2] == 'P' && t[i + 3] == '<';
    bool startsWithName = t.size() - i > 3 && t[i + 2] == '<';
    if (startsWithP || startsWithName) {
      size_t exprStart = startsWithName ? 3 : 4;
      size_t end = t.find('>', i);
      if (end == std::string::npos) {
        err = {ErrInvalidNamedCapture, t.substr(s0)};
        return false;
      }
      std::string capture = t.substr(i, end + 1 - i);
      std::string name = t.substr(i + exprStart, end - (i + exprStart));
      if (!isValidCaptureName(name)) {
        err = {ErrInvalidNamedCapture, capture};
        return false;
      }
      numCap++;
      Regexp* re = op(opLeftParen);
      re->cap = numCap;
      re->name = name;
      i = end + 1;
      return true;
    }
    i += 2;  // "(?"
    uint16_t fl = flags;
    int sign = +1;
    bool sawFlag = false;
    while (i < t.size()) {
      int32_t c;
      if (!nextRune(t, i, c)) {
        err = {ErrInvalidUTF8, t.substr(s0)};
        return false;
      }
      switch (c) {
        default:
          goto bad;
        case 'i': fl |= FoldCase; sawFlag = true; break;
        case 'm': fl &= ~OneLine; sawFlag = true; break;
        case 's': fl |= DotNL; sawFlag = true; break;
        case 'U': fl |= NonGreedy; sawFlag = true; break;
        case '-':
          if (sign < 0) goto bad;
          sign = -1;
          fl = ~fl;
          sawFlag = false;
          break;
        case ':':
        case ')':
          if (sign < 0) {
            if (!sawFlag) goto bad;
            fl = ~fl;
          }
          if (c == ':') op(opLeftParen);
          flags = fl;
          return true;
      }
    }
  bad:
    err = {ErrInvalidPerlOp, t.substr(s0, i - s0)};
    return false;
  }

  Regexp* parse(const std::string& s, Error& err) {
    whole = s;
    const std::string& t = s;
    size_t i = 0;
    size_t lastRepeat = std::string::npos;
    while (i < t.size()) {
      size_t repeatPos = std::string::npos;
      switch (t[i]) {
        default: {
          int32_t c;
          if (!nextRune(t, i, c)) {
            err = {ErrInvalidUTF8, t.substr(i)};
            return nullptr;
          }
          literal(c);
          break;
        }
        case '(':
          if ((flags & PerlX) && i + 1 < t.size() && t[i + 1] == '?') {
            if (!parsePerlFlags(t, i, err)) return nullptr;
            break;
          }
          numCap++;
          op(opLeftParen)->cap = numCap;
          i++;
          break;
        case '|':
          parseVerticalBar();
          i++;
          break;
        case ')':
          if (!parseRightParen(err)) return nullptr;
          i++;
          break;
        case '^':
          if (flags & OneLine)
            op(OpBeginText);
          else
            op(OpBeginLine);
          i++;
          break;
        case '$':
          if (flags & OneLine)
            op(OpEndText)->flags |= WasDollar;
          else
            op(OpEndLine);
          i++;
          break;
        case '.':
          if (flags & DotNL)
            op(OpAnyChar);
          else
            op(OpAnyCharNotNL);
          i++;
          break;
        case '[':
          if (!parseClass(t, i, err)) return nullptr;
          break;
        case '*':
        case '+':
        case '?': {
          size_t before = i;
          Op o = t[i] == '*' ? OpStar : t[i] == '+' ? OpPlus : OpQuest;
          size_t after = i + 1;
          if (!repeat(o, 0, 0, t, before, after, lastRepeat, err)) return nullptr;
          repeatPos = before;
          i = after;
          break;
        }
        case '{': {
          size_t before = i;
          int mn, mx;
          size_t after;
          if (!parseRepeat(t, i, mn, mx, after)) {
            literal('{');
            i++;
            break;
          }
          if (mn < 0 || mn > 1000 || mx > 1000 || (mx >= 0 && mn > mx)) {
            err = {ErrInvalidRepeatSize, t.substr(before, after - before)};
            return nullptr;
          }
          if (!repeat(OpRepeat, mn, mx, t, before, after, lastRepeat, err)) return nullptr;
          repeatPos = before;
          i = after;
          break;
        }
        case '\\': {
          if ((flags & PerlX) && i + 1 < t.size()) {
            bool handled = true;
            switch (t[i + 1]) {
              case 'A': op(OpBeginText); i += 2; break;
              case 'b': op(OpWordBoundary); i += 2; break;
              case 'B': op(OpNoWordBoundary); i += 2; break;
              case 'C':
                err = {ErrInvalidEscape, t.substr(i, 2)};
                return nullptr;
              case 'Q': {
                size_t e = t.find("\\E", i + 2);
                std::string lit = e == std::string::npos ? t.substr(i + 2) : t.substr(i + 2, e - (i + 2));
                size_t k = 0;
                while (k < lit.size()) {
                  int32_t c;
                  if (!nextRune(lit, k, c)) {
                    err = {ErrInvalidUTF8, lit.substr(k)};
                    return nullptr;
                  }
                  literal(c);
                }
                i = e == std::string::npos ? t.size() : e + 2;
                break;
              }
              case 'z': op(OpEndText); i += 2; break;
              default: handled = false;
            }
            if (handled) break;
          }
          Regexp* re = newRegexp(OpCharClass);
          re->flags = flags;
          if (i + 1 < t.size() && (t[i + 1] == 'p' || t[i + 1] == 'P')) {
            int k = parseUnicodeClass(t, i, re->rune, err);
            if (k < 0) return nullptr;
            if (k > 0) {
              push(re);
              break;
            }
          }
          if (parsePerlClassEscape(t, i, re->rune)) {
            push(re);
            break;
          }
          int32_t c;
          if (!parseEscape(t, i, c, err)) return nullptr;
          literal(c);
          break;
        }
      }
      lastRepeat = repeatPos;
    }
    concat();
    if (swapVerticalBar()) stack.pop_back();
    alternate();
    if (stack.size() != 1) {
      err = {ErrMissingParen, s};
      return nullptr;
    }
    return stack[0];
  }
};

}  // namespace

bool Regexp::equal(const Regexp* y) const {
  const Regexp* x = this;
  if (x == nullptr || y == nullptr) return x == y;
  if (x->op != y->op) return false;
  switch (x->op) {
    case OpEndText:
      if ((x->flags & WasDollar) != (y->flags & WasDollar)) return false;
      break;
    case OpLiteral:
    case OpCharClass:
      if (x->rune != y->rune) return false;
      if (x->op == OpLiteral && (x->flags & FoldCase) != (y->flags & FoldCase)) return false;
      break;
    case OpAlternate:
    case OpConcat:
      if (x->sub.size() != y->sub.size()) return false;
      for (size_t i = 0; i < x->sub.size(); i++)
        if (!x->sub[i]->equal(y->sub[i])) return false;
      break;
    case OpStar:
    case OpPlus:
    case OpQuest:
      if ((x->flags & NonGreedy) != (y->flags & NonGreedy) || !x->sub[0]->equal(y->sub[0]))
        return false;
      break;
    case OpRepeat:
      if ((x->flags & NonGreedy) != (y->flags & NonGreedy) || x->min != y->min ||
          x->max != y->max || !x->sub[0]->equal(y->sub[0]))
        return false;
      break;
    case OpCapture:
      if (x->cap != y->cap || x->name != y->name || !x->sub[0]->equal(y->sub[0])) return false;
      break;
    default:
      break;
  }
  return true;
}

ParseResult Parse(const std::string& pattern, uint16_t flags, Arena& arena) {
  ParseResult out;
  Parser p(arena, flags);
  Error err;
  Regexp* re = p.parse(pattern, err);
  if (!re) {
    out.err = "error parsing regexp: " + err.code;
    if (!err.expr.empty() || err.code != ErrTrailingBackslash)
      out.err += ": `" + err.expr + "`";
    return out;
  }
  out.re = re;
  out.num_cap = p.numCap;
  return out;
}

static void dumpRunes(const Regexp* re, std::string& s, bool pairs) {
  char buf[32];
  for (size_t i = 0; i < re->rune.size(); i++) {
    if (pairs) {
      if (i % 2 == 0) {
        if (i) s += ' ';
        snprintf(buf, sizeof buf, "0x%x", re->rune[i]);
        s += buf;
      } else {
        snprintf(buf, sizeof buf, "-0x%x", re->rune[i]);
        s += buf;
      }
    } else {
      int32_t r = re->rune[i];
      if (r >= 0x20 && r < 0x7f && r != '{' && r != '}' && r != '\\') {
        s += (char)r;
      } else {
        snprintf(buf, sizeof buf, "\\x{%x}", r);
        s += buf;
      }
    }
  }
}

static void dump(const Regexp* re, std::string& s) {
  static const char* names[] = {"",     "no",   "emp",  "lit",  "cc",   "dnl",  "dot",
                                "bol",  "eol",  "bot",  "eot",  "wb",   "nwb",  "cap",
                                "star", "plus", "que",  "rep",  "cat",  "alt"};
  switch (re->op) {
    case OpStar: case OpPlus: case OpQuest: case OpRepeat:
      if (re->flags & NonGreedy) s += 'n';
      break;
    case OpLiteral:
      if (re->flags & FoldCase) s += "fold";
      break;
    default: break;
  }
  s += names[re->op];
  s += '{';
  switch (re->op) {
    case OpLiteral: dumpRunes(re, s, false); break;
    case OpCharClass: dumpRunes(re, s, true); break;
    case OpRepeat: {
      char buf[48];
      snprintf(buf, sizeof buf, "%d,%d ", re->min, re->max);
      s += buf;
      break;
    }
    case OpCapture:
      if (!re->name.empty()) s += re->name + ":";
      break;
    default: break;
  }
  for (auto* x : re->sub) dump(x, s);
  s += '}';
}

std::string Dump(const Regexp* re) {
  std::string s;
  dump(re, s);
  return s;
}

int MaxCap(const Regexp* re) {
  int m = 0;
  if (re->op == OpCapture) m = re->cap;
  for (auto* s : re->sub) m = std::max(m, MaxCap(s));
  return m;
}

}  // namespace gosyntax
