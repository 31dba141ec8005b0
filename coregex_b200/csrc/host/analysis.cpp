// analysis.cpp — see analysis.h.
#include "analysis.h"

#include <algorithm>
#include <set>

namespace cgx {

using namespace gosyntax;

const char* RefStrategyName(int s) {
  static const char* n[] = {"UseNFA", "UseDFA", "UseBoth", "UseReverseAnchored", "UseReverseSuffix",
                            "UseOnePass", "UseReverseInner", "UseBoundedBacktracker", "UseTeddy",
                            "UseReverseSuffixSet", "UseCharClassSearcher", "UseCompositeSearcher",
                            "UseBranchDispatch", "UseDigitPrefilter", "UseAhoCorasick",
                            "UseAnchoredLiteral", "UseMultilineReverseSuffix"};
  return (s >= 0 && s < 17) ? n[s] : "?";
}

namespace {

constexpr int kMaxLiterals = 256, kMaxLiteralLen = 64, kMaxClassSize = 10, kCrossLimit = 250;

// ---- how many states the reference's Thompson compiler would allocate (nfa/compile.go) --------
int utf8Len(int32_t r) { return r < 0x80 ? 1 : r < 0x800 ? 2 : r < 0x10000 ? 3 : 4; }

bool nullable(const Regexp* re) {
  switch (re->op) {
    case OpEmptyMatch: return true;
    case OpLiteral: return re->rune.empty();
    case OpCharClass: case OpAnyCharNotNL: case OpAnyChar: case OpNoMatch: return false;
    case OpCapture: return re->sub.empty() || nullable(re->sub[0]);
    case OpStar: case OpQuest: return true;
    case OpPlus: return !re->sub.empty() && nullable(re->sub[0]);
    case OpRepeat: return re->min == 0 || (!re->sub.empty() && nullable(re->sub[0]));
    case OpConcat:
      for (auto* s : re->sub) if (!nullable(s)) return false;
      return true;
    case OpAlternate:
      for (auto* s : re->sub) if (nullable(s)) return true;
      return false;
    default: return true;
  }
}

int refStates(const Regexp* re) {
  switch (re->op) {
    case OpLiteral: {
      if (re->rune.empty()) return 1;
      int n = 0;
      for (int32_t r : re->rune) {
        bool letter = (r >= 'a' && r <= 'z') || (r >= 'A' && r <= 'Z');
        if ((re->flags & FoldCase) && letter) n += 4; else n += utf8Len(r);
      }
      return n;
    }
    case OpCharClass:
      if (re->rune.empty()) return 2;
      return re->rune.size() == 2 ? 1 : 2;
    case OpConcat: {
      int n = 0;
      for (auto* s : re->sub) n += refStates(s);
      return re->sub.empty() ? 1 : n;
    }
    case OpAlternate: {
      if (re->sub.size() == 1) return refStates(re->sub[0]);
      int n = 0;
      for (auto* s : re->sub) n += refStates(s);
      return n + (int)re->sub.size() - 1 + 1;
    }
    case OpStar: return refStates(re->sub[0]) + (nullable(re->sub[0]) ? 3 : 2);
    case OpPlus: case OpQuest: return refStates(re->sub[0]) + 2;
    case OpRepeat: {
      int s = refStates(re->sub[0]);
      if (re->max == -1) {
        if (re->min == 0) return s + (nullable(re->sub[0]) ? 3 : 2);
        return re->min * s + s + (nullable(re->sub[0]) ? 3 : 2);
      }
      if (re->min == re->max) return re->min == 0 ? 1 : re->min * s;
      return re->min * s + (re->max - re->min) * (s + 2);
    }
    case OpCapture: return re->sub.empty() ? 1 : refStates(re->sub[0]) + 2;
    default: return 1;
  }
}

// ---- predicates over the AST (reference meta/strategy.go) --------------------------------------
bool anyOp(const Regexp* re, bool (*pred)(Op)) {
  if (pred(re->op)) return true;
  for (auto* s : re->sub)
    if (anyOp(s, pred)) return true;
  return false;
}

bool digitOnlyClass(const std::vector<int32_t>& r) {
  if (r.empty() || r.size() % 2) return false;
  for (size_t i = 0; i < r.size(); i += 2)
    if (r[i] < '0' || r[i + 1] > '9') return false;
  return true;
}

bool digitLead(const Regexp* re);

bool optionalDigitOnly(const Regexp* re) {
  if (re->sub.empty()) return false;
  const Regexp* s = re->sub[0];
  if (s->op == OpCharClass) return digitOnlyClass(s->rune);
  if (s->op == OpLiteral) {
    for (int32_t r : s->rune)
      if (r < '0' || r > '9') return false;
    return !s->rune.empty();
  }
  return digitLead(s);
}

bool digitLead(const Regexp* re) {
  switch (re->op) {
    case OpCharClass: return digitOnlyClass(re->rune);
    case OpLiteral: return !re->rune.empty() && re->rune[0] >= '0' && re->rune[0] <= '9';
    case OpAlternate:
      if (re->sub.empty()) return false;
      for (auto* s : re->sub) if (!digitLead(s)) return false;
      return true;
    case OpConcat:
      for (auto* s : re->sub) {
        bool optional = s->op == OpQuest || s->op == OpStar || (s->op == OpRepeat && s->min == 0);
        if (optional) {
          if (!optionalDigitOnly(s)) return false;
          continue;
        }
        return digitLead(s);
      }
      return false;
    case OpCapture: case OpPlus: return !re->sub.empty() && digitLead(re->sub[0]);
    case OpRepeat: return !re->sub.empty() && re->min >= 1 && digitLead(re->sub[0]);
    default: return false;
  }
}

bool digitRunSkipSafe(const Regexp* re) {
  switch (re->op) {
    case OpConcat: case OpCapture: return !re->sub.empty() && digitRunSkipSafe(re->sub[0]);
    case OpPlus: case OpStar:
      return re->sub.size() == 1 && re->sub[0]->op == OpCharClass && digitOnlyClass(re->sub[0]->rune);
    case OpRepeat:
      return re->max == -1 && re->sub.size() == 1 && re->sub[0]->op == OpCharClass &&
             digitOnlyClass(re->sub[0]->rune);
    default: return false;
  }
}

bool simpleCharClass(const Regexp* re) {
  switch (re->op) {
    case OpCharClass: return true;
    case OpPlus: case OpStar: case OpQuest: case OpRepeat:
      return re->sub.size() == 1 && simpleCharClass(re->sub[0]);
    case OpConcat:
      for (auto* s : re->sub) if (!simpleCharClass(s)) return false;
      return true;
    case OpCapture: return re->sub.size() == 1 && simpleCharClass(re->sub[0]);
    default: return false;
  }
}

bool endAnchoredTail(const Regexp* re) {
  switch (re->op) {
    case OpEndText: return true;
    case OpConcat: return !re->sub.empty() && endAnchoredTail(re->sub.back());
    case OpCapture: return !re->sub.empty() && endAnchoredTail(re->sub[0]);
    default: return false;
  }
}

// ---- literal sequences ---------------------------------------------------------------------------
using Lits = std::vector<Lit>;

std::string utf8(const std::vector<int32_t>& rs) {
  std::string s;
  for (int32_t r : rs) {
    if (r < 0x80) s += (char)r;
    else if (r < 0x800) { s += (char)(0xC0 | (r >> 6)); s += (char)(0x80 | (r & 0x3F)); }
    else if (r < 0x10000) {
      s += (char)(0xE0 | (r >> 12)); s += (char)(0x80 | ((r >> 6) & 0x3F)); s += (char)(0x80 | (r & 0x3F));
    } else {
      s += (char)(0xF0 | (r >> 18)); s += (char)(0x80 | ((r >> 12) & 0x3F));
      s += (char)(0x80 | ((r >> 6) & 0x3F)); s += (char)(0x80 | (r & 0x3F));
    }
  }
  return s;
}

void keepFirst(Lits& v, size_t n) {
  for (auto& l : v)
    if (l.bytes.size() > n) { l.bytes.resize(n); l.complete = false; }
}
void dedup(Lits& v) {
  std::set<std::string> seen;
  Lits out;
  for (auto& l : v) if (seen.insert(l.bytes).second) out.push_back(l);
  v.swap(out);
}
void inexact(Lits& v) { for (auto& l : v) l.complete = false; }

int32_t fold1(int32_t r) {
  if (r == 'K') return 'k'; if (r == 'k') return 0x212A; if (r == 0x212A) return 'K';
  if (r == 'S') return 's'; if (r == 's') return 0x17F; if (r == 0x17F) return 'S';
  if (r >= 'A' && r <= 'Z') return r + 32;
  if (r >= 'a' && r <= 'z') return r - 32;
  return r;
}

struct Extract {
  Lits foldLiteral(const std::vector<int32_t>& runes) {
    if (runes.empty()) return {};
    std::vector<std::vector<int32_t>> sets(runes.size());
    long total = 1;
    size_t filled = 0;
    for (size_t i = 0; i < runes.size(); i++) {
      sets[i] = {runes[i]};
      for (int32_t f = fold1(runes[i]); f != runes[i]; f = fold1(f)) sets[i].push_back(f);
      filled = i + 1;
      total *= (long)sets[i].size();
      if (total > kCrossLimit) break;
    }
    auto gen = [&](size_t n) {
      std::vector<std::vector<int32_t>> var{{}};
      for (size_t i = 0; i < n; i++) {
        std::vector<std::vector<int32_t>> nx;
        for (auto& p : var) for (int32_t r : sets[i]) { auto e = p; e.push_back(r); nx.push_back(e); }
        var.swap(nx);
      }
      Lits out;
      for (auto& v : var) {
        std::string b = utf8(v);
        if ((int)b.size() > kMaxLiteralLen) b.resize(kMaxLiteralLen);
        out.push_back({b, true});
      }
      return out;
    };
    if (total <= kMaxLiterals && filled == runes.size()) return gen(filled);
    size_t trim = sets.size();
    long prod = 1;
    for (size_t i = 0; i < sets.size(); i++) {
      prod *= (long)sets[i].size();
      if (prod > kMaxLiterals) { trim = i; break; }
    }
    if (trim == 0) return {};
    Lits r = gen(trim);
    inexact(r);
    dedup(r);
    if ((int)r.size() > kMaxLiterals) r.resize(kMaxLiterals);
    return r;
  }

  Lits expandClass(const Regexp* re) {
    long count = 0;
    for (size_t i = 0; i + 1 < re->rune.size(); i += 2) {
      count += re->rune[i + 1] - re->rune[i] + 1;
      if (count > kMaxClassSize) return {};
    }
    Lits out;
    for (size_t i = 0; i + 1 < re->rune.size(); i += 2)
      for (int32_t r = re->rune[i]; r <= re->rune[i + 1]; r++) {
        out.push_back({utf8({r}), true});
        if ((int)out.size() >= kMaxLiterals) return out;
      }
    return out;
  }

  // nil contribution == (false, {})
  bool contribution(const Regexp* sub, int depth, Lits& out) {
    switch (sub->op) {
      case OpLiteral:
        out = (sub->flags & FoldCase) ? foldLiteral(sub->rune) : Lits{{utf8(sub->rune), true}};
        return true;
      case OpCharClass:
        out = expandClass(sub);
        return !out.empty();
      case OpAlternate: {
        Lits all;
        bool over = false;
        for (auto* s : sub->sub) {
          Lits q = prefixes(s, depth + 1);
          if (q.empty()) return false;
          if (over) {
            for (auto& l : q) { std::string b = l.bytes.substr(0, 3); all.push_back({b, false}); }
            if ((int)all.size() > kCrossLimit) dedup(all);
            continue;
          }
          for (auto& l : q) all.push_back(l);
          if ((int)all.size() > kCrossLimit) { over = true; keepFirst(all, 3); inexact(all); dedup(all); }
        }
        if (over || (int)all.size() > kMaxLiterals) {
          keepFirst(all, 3); inexact(all); dedup(all);
          if ((int)all.size() > kMaxLiterals) all.resize(kMaxLiterals);
        }
        out = all;
        return true;
      }
      case OpCapture: return !sub->sub.empty() && contribution(sub->sub[0], depth, out);
      case OpRepeat:
        if (sub->min >= 1 && !sub->sub.empty()) {
          if (!contribution(sub->sub[0], depth, out)) return false;
          inexact(out);
          return true;
        }
        return false;
      case OpWordBoundary: case OpNoWordBoundary:
        out = {{"", true}};
        return true;
      default: return false;
    }
  }

  Lits prefixes(const Regexp* re, int depth) {
    if (depth > 100) return {};
    switch (re->op) {
      case OpLiteral: {
        if (re->flags & FoldCase) return foldLiteral(re->rune);
        std::string b = utf8(re->rune);
        if ((int)b.size() > kMaxLiteralLen) b.resize(kMaxLiteralLen);
        return {{b, true}};
      }
      case OpConcat: {
        size_t start = 0;
        while (start < re->sub.size() && (re->sub[start]->op == OpBeginLine || re->sub[start]->op == OpBeginText)) start++;
        if (start >= re->sub.size()) return {};
        Lits acc{{"", true}};
        for (size_t i = start; i < re->sub.size(); i++) {
          bool anyExact = false;
          for (auto& l : acc) anyExact |= l.complete;
          if (!anyExact) break;
          Lits c;
          if (!contribution(re->sub[i], depth, c)) { inexact(acc); break; }
          Lits nx;
          for (auto& l : acc) {
            if (!l.complete) { nx.push_back(l); continue; }
            for (auto& r : c) nx.push_back({l.bytes + r.bytes, r.complete});
          }
          if (!acc.empty() && !c.empty()) acc.swap(nx);
          if ((int)acc.size() > kCrossLimit || (int)acc.size() > kMaxLiterals) {
            keepFirst(acc, 4); inexact(acc); dedup(acc);
            if ((int)acc.size() > kMaxLiterals) acc.resize(kMaxLiterals);
            break;
          }
          for (auto& l : acc)
            if ((int)l.bytes.size() > kMaxLiteralLen) { l.bytes.resize(kMaxLiteralLen); l.complete = false; }
        }
        if (acc.size() == 1 && acc[0].bytes.empty()) return {};
        return acc;
      }
      case OpAlternate: {
        Lits all;
        bool over = false;
        for (auto* s : re->sub) {
          Lits q = prefixes(s, depth + 1);
          if (q.empty()) return {};
          for (auto& l : q) {
            all.push_back(l);
            if ((int)all.size() > kCrossLimit) { over = true; break; }
          }
          if (over) break;
        }
        if (over || (int)all.size() > kMaxLiterals) {
          keepFirst(all, 3); inexact(all); dedup(all);
          if ((int)all.size() > kMaxLiterals) all.resize(kMaxLiterals);
        }
        return all;
      }
      case OpCharClass: return expandClass(re);
      case OpCapture: return re->sub.empty() ? Lits{} : prefixes(re->sub[0], depth + 1);
      default: return {};
    }
  }

  Lits suffixes(const Regexp* re, int depth) {
    if (depth > 100) return {};
    switch (re->op) {
      case OpLiteral: {
        if (re->flags & FoldCase) return foldLiteral(re->rune);
        std::string b = utf8(re->rune);
        if ((int)b.size() > kMaxLiteralLen) b = b.substr(b.size() - kMaxLiteralLen);
        return {{b, true}};
      }
      case OpConcat: {
        int last = (int)re->sub.size() - 1;
        while (last >= 0) {
          Op o = re->sub[last]->op;
          if (o != OpEndLine && o != OpEndText && o != OpWordBoundary && o != OpNoWordBoundary) break;
          last--;
        }
        if (last < 0) return {};
        Lits suf = suffixes(re->sub[last], depth + 1);
        if (suf.empty()) return {};
        for (int i = last - 1; i >= 0; i--) {
          const Regexp* s = re->sub[i];
          if (s->op == OpWordBoundary || s->op == OpNoWordBoundary) continue;
          if (s->op != OpLiteral) { inexact(suf); return suf; }
          std::string pre = utf8(s->rune);
          for (auto& l : suf) {
            l.bytes = pre + l.bytes;
            if ((int)l.bytes.size() > kMaxLiteralLen) l.bytes = l.bytes.substr(l.bytes.size() - kMaxLiteralLen);
          }
          if ((int)suf.size() > kMaxLiterals) return suf;
        }
        return suf;
      }
      case OpAlternate: {
        Lits all;
        for (auto* s : re->sub) {
          Lits q = suffixes(s, depth + 1);
          if (q.empty()) return {};
          for (auto& l : q) { all.push_back(l); if ((int)all.size() >= kMaxLiterals) return all; }
        }
        return all;
      }
      case OpCharClass: return expandClass(re);
      case OpCapture: return re->sub.empty() ? Lits{} : suffixes(re->sub[0], depth + 1);
      default: return {};
    }
  }

  Lits inner(const Regexp* re, int depth) {
    if (depth > 100) return {};
    switch (re->op) {
      case OpLiteral: {
        Lits r = (re->flags & FoldCase) ? foldLiteral(re->rune) : Lits{{utf8(re->rune).substr(0, kMaxLiteralLen), false}};
        inexact(r);
        return r;
      }
      case OpConcat:
        for (auto* s : re->sub) { Lits q = inner(s, depth + 1); if (!q.empty()) return q; }
        return {};
      case OpAlternate: {
        Lits all;
        for (auto* s : re->sub) {
          Lits q = inner(s, depth + 1);
          if (q.empty()) return {};
          for (auto& l : q) { all.push_back(l); if ((int)all.size() >= kMaxLiterals) return all; }
        }
        return all;
      }
      case OpCharClass: return expandClass(re);
      case OpCapture: return re->sub.empty() ? Lits{} : inner(re->sub[0], depth + 1);
      default: return {};
    }
  }
};

std::string lcp(const Lits& v) {
  if (v.empty()) return "";
  std::string p = v[0].bytes;
  for (size_t i = 1; i < v.size(); i++) {
    size_t k = 0;
    while (k < p.size() && k < v[i].bytes.size() && p[k] == v[i].bytes[k]) k++;
    p.resize(k);
  }
  return p;
}
std::string lcs(const Lits& v) {
  if (v.empty()) return "";
  std::string s = v[0].bytes;
  for (size_t i = 1; i < v.size(); i++) {
    const std::string& b = v[i].bytes;
    size_t k = 0;
    while (k < s.size() && k < b.size() && s[s.size() - 1 - k] == b[b.size() - 1 - k]) k++;
    s = s.substr(s.size() - k);
  }
  return s;
}
bool allComplete(const Lits& v) {
  if (v.empty()) return false;
  for (auto& l : v) if (!l.complete) return false;
  return true;
}
size_t minLen(const Lits& v) {
  size_t m = (size_t)-1;
  for (auto& l : v) m = std::min(m, l.bytes.size());
  return m;
}

bool wildcardOrRep(const Regexp* re) {
  switch (re->op) {
    case OpStar: case OpPlus: case OpQuest: case OpRepeat: case OpAnyChar: case OpAnyCharNotNL: return true;
    case OpConcat: case OpAlternate:
      for (auto* s : re->sub) if (wildcardOrRep(s)) return true;
      return false;
    case OpCapture: return !re->sub.empty() && wildcardOrRep(re->sub[0]);
    default: return false;
  }
}
bool wildcardSub(const Regexp* re) {
  while (re->op == OpCapture && !re->sub.empty()) re = re->sub[0];
  if ((re->op == OpStar || re->op == OpPlus) && !re->sub.empty() &&
      (re->sub[0]->op == OpAnyChar || re->sub[0]->op == OpAnyCharNotNL)) return true;
  if (re->op == OpPlus && !re->sub.empty() && re->sub[0]->op == OpCharClass) return true;
  return re->op == OpRepeat && re->min >= 1;
}
bool containsAnchor(const Regexp* re) {
  switch (re->op) {
    case OpBeginLine: case OpEndLine: case OpBeginText: case OpEndText: return true;
    case OpConcat: case OpAlternate:
      for (auto* s : re->sub) if (containsAnchor(s)) return true;
      return false;
    case OpCapture: case OpStar: case OpPlus: case OpQuest: case OpRepeat:
      return !re->sub.empty() && containsAnchor(re->sub[0]);
    default: return false;
  }
}
bool safeReverseSuffix(const Regexp* re) {
  if (re->op == OpCapture) return !re->sub.empty() && safeReverseSuffix(re->sub[0]);
  if (re->op != OpConcat || re->sub.size() < 2) return false;
  int wc = 0;
  for (size_t i = 0; i + 1 < re->sub.size(); i++) wc += wildcardSub(re->sub[i]);
  if (!wc) return false;
  for (size_t i = 1; i + 1 < re->sub.size(); i++) if (containsAnchor(re->sub[i])) return false;
  return true;
}
bool safeReverseInner(const Regexp* re) {
  if (re->op == OpCapture) return !re->sub.empty() && safeReverseInner(re->sub[0]);
  if (re->op != OpConcat || re->sub.size() < 2) return false;
  const Regexp* f = re->sub[0];
  if ((f->op == OpStar || f->op == OpPlus) && !f->sub.empty() &&
      (f->sub[0]->op == OpAnyChar || f->sub[0]->op == OpAnyCharNotNL)) return true;
  return f->op == OpPlus && !f->sub.empty() && f->sub[0]->op == OpCharClass;
}

}  // namespace

Analysis Analyze(const Regexp* re, int, bool anchored_start, const AnalysisConfig& cfg) {
  Analysis a;
  Extract ex;
  a.digit_lead = digitLead(re);
  a.digit_run_skip_safe = digitRunSkipSafe(re);
  a.can_match_empty = nullable(re);
  a.has_anchors = anyOp(re, [](Op o) {
    return o == OpBeginLine || o == OpEndLine || o == OpBeginText || o == OpEndText ||
           o == OpWordBoundary || o == OpNoWordBoundary;
  });
  bool nonLine = anyOp(re, [](Op o) {
    return o == OpEndLine || o == OpEndText || o == OpBeginText || o == OpWordBoundary || o == OpNoWordBoundary;
  });
  bool hasWB = anyOp(re, [](Op o) { return o == OpWordBoundary || o == OpNoWordBoundary; });
  bool hasML = anyOp(re, [](Op o) { return o == OpBeginLine || o == OpEndLine; });
  bool hasStartAnchor = anyOp(re, [](Op o) { return o == OpBeginText; });
  bool hasEndText = anyOp(re, [](Op o) { return o == OpEndText; });

  const size_t minlit = (size_t)(cfg.min_literal_len > 0 ? cfg.min_literal_len : 1);
  if (!anchored_start && cfg.enable_prefilter) {  // reference meta/compile.go:466
    a.prefixes = ex.prefixes(re, 0);
    if (a.prefixes.size() > 64) {  // reference literal/extractor.go:135-149
      Lits orig = a.prefixes;
      for (int keep : {4, 3, 2}) {
        if (a.prefixes.size() <= 64) break;
        keepFirst(a.prefixes, keep);
        dedup(a.prefixes);
      }
      if (a.prefixes.size() > 64) a.prefixes = orig;
    }
  }
  a.prefixes_all_complete = allComplete(a.prefixes);
  const Lits& P = a.prefixes;
  const int nfaSize = refStates(re) + 1 + (anchored_start ? 0 : 2);

  auto decide = [&]() -> int {
    if (cfg.enable_dfa && endAnchoredTail(re) && !anchored_start && !hasStartAnchor) return RS_UseReverseAnchored;
    if (anchored_start) return RS_UseBoundedBacktracker;
    // --- selectReverseStrategy (needs the DFA and the prefilter, meta/strategy.go:976)
    if (!hasWB && !hasEndText && cfg.enable_dfa && cfg.enable_prefilter) {
      bool fastPrefix = !P.empty() && (lcp(P).size() >= minlit || P.size() == 1 || minLen(P) >= 3);
      if (!fastPrefix) {
        Lits suf = ex.suffixes(re, 0);
        if (!suf.empty() && lcs(suf).size() >= minlit) {
          if (safeReverseSuffix(re)) return RS_UseReverseSuffix;
          goto no_reverse;
        }
        if (safeReverseSuffix(re) && !suf.empty()) {
          bool exactAlt = allComplete(P) && P.size() == suf.size();
          bool ok = !exactAlt && suf.size() >= 2 && suf.size() <= 32;
          for (auto& l : suf) if (l.bytes.size() < 2) ok = false;
          if (ok) return RS_UseReverseSuffixSet;
        }
        if (re->op == OpConcat && re->sub.size() >= 3) {
          for (size_t i = 1; i + 1 < re->sub.size(); i++) {
            Lits in = ex.inner(re->sub[i], 0);
            if (in.empty()) continue;
            bool before = false, after = false;
            for (size_t j = 0; j < i; j++) before |= wildcardOrRep(re->sub[j]);
            for (size_t j = i + 1; j < re->sub.size(); j++) after |= wildcardOrRep(re->sub[j]);
            if (before && after) {
              std::string p = lcp(in);
              if (p.size() == 1 && a.digit_lead) goto no_reverse;
              if (p.size() >= minlit) {
                if (!safeReverseInner(re)) goto no_reverse;
                a.inner_idx = (int)i;
                a.inner_literal = p;
                return RS_UseReverseInner;
              }
              break;
            }
          }
        }
      }
    }
  no_reverse:
    if (!cfg.enable_dfa) return RS_UseNFA;  // meta/strategy.go:1447
    bool hasGood = !P.empty() && lcp(P).size() >= minlit;
    bool hasTeddy = false, hasAC = false;
    if (P.size() >= 2 && P.size() <= 64) {
      hasTeddy = true;
      for (auto& l : P) if (l.bytes.size() < 3) hasTeddy = false;
    }
    if (P.size() > 64) {
      hasAC = true;
      for (auto& l : P) if (l.bytes.empty()) hasAC = false;
    }
    if (!hasGood && !hasTeddy && simpleCharClass(re)) {
      bool single = re->op == OpPlus || re->op == OpCapture;
      return single ? RS_UseCharClassSearcher : RS_UseBoundedBacktracker;
    }
    if (!P.empty()) {
      if (hasTeddy && allComplete(P) && !(a.has_anchors && nonLine)) return RS_UseTeddy;
      if (hasAC && allComplete(P)) return RS_UseAhoCorasick;
    }
    if (nfaSize <= 100 && a.digit_lead && cfg.enable_prefilter) return RS_UseDigitPrefilter;
    if (nfaSize < 20) {
      if ((hasWB && a.has_anchors) || a.can_match_empty || hasML) return RS_UseNFA;
      return RS_UseDFA;
    }
    if (!hasGood && !hasTeddy && a.can_match_empty) return RS_UseNFA;
    if (hasGood || hasTeddy) {
      if (nfaSize > 200 && !allComplete(P)) return RS_UseNFA;
      return RS_UseDFA;
    }
    if (nfaSize > 100) return RS_UseNFA;
    return RS_UseBoth;
  };
  a.strategy = decide();
  return a;
}

}  // namespace cgx
