// oracle/meta.cpp — TEST INFRASTRUCTURE ONLY.  See meta.h for the reference citations.
#include "meta.h"

#include "literal.h"
#include "pikevm.h"
#include "revsearch.h"
#include "simd.h"
#include "teddy.h"

namespace oracle {

using namespace gosyntax;

const char* StrategyName(int s) {
  static const char* n[] = {"UseNFA", "UseDFA", "UseBoth", "UseReverseAnchored", "UseReverseSuffix",
                            "UseOnePass", "UseReverseInner", "UseBoundedBacktracker", "UseTeddy",
                            "UseReverseSuffixSet", "UseCharClassSearcher", "UseCompositeSearcher",
                            "UseBranchDispatch", "UseDigitPrefilter", "UseAhoCorasick",
                            "UseAnchoredLiteral", "UseMultilineReverseSuffix"};
  if (s < 0 || s >= (int)(sizeof n / sizeof *n)) return "?";
  return n[s];
}

// ---- AST predicates (reference meta/strategy.go) -------------------------------------------

static bool isDigitOnlyClass(const std::vector<int32_t>& r) {
  if (r.empty() || r.size() % 2) return false;
  for (size_t i = 0; i < r.size(); i += 2)
    if (r[i] < '0' || r[i + 1] > '9') return false;
  return true;
}

static bool isOptionalElement(const Regexp* re) {
  return re->op == OpQuest || re->op == OpStar || (re->op == OpRepeat && re->min == 0);
}

// reference meta/strategy.go:365-390
static bool isOptionalDigitOnly(const Regexp* re) {
  if (re->sub.empty()) return false;
  const Regexp* sub = re->sub[0];
  switch (sub->op) {
    case OpCharClass: return isDigitOnlyClass(sub->rune);
    case OpLiteral:
      for (int32_t r : sub->rune)
        if (r < '0' || r > '9') return false;
      return !sub->rune.empty();
    default: return isDigitLeadPattern(sub);
  }
}

bool isDigitLeadPattern(const Regexp* re) {
  if (!re) return false;
  switch (re->op) {
    case OpCharClass: return isDigitOnlyClass(re->rune);
    case OpLiteral: return !re->rune.empty() && re->rune[0] >= '0' && re->rune[0] <= '9';
    case OpAlternate:
      if (re->sub.empty()) return false;
      for (auto* s : re->sub)
        if (!isDigitLeadPattern(s)) return false;
      return true;
    case OpConcat:
      if (re->sub.empty()) return false;
      for (auto* s : re->sub) {
        if (isOptionalElement(s)) {
          if (!isOptionalDigitOnly(s)) return false;
          continue;
        }
        return isDigitLeadPattern(s);
      }
      return false;
    case OpCapture:
    case OpPlus:
      return !re->sub.empty() && isDigitLeadPattern(re->sub[0]);
    case OpRepeat:
      return !re->sub.empty() && re->min >= 1 && isDigitLeadPattern(re->sub[0]);
    default: return false;
  }
}

bool isDigitRunSkipSafe(const Regexp* re) {
  if (!re) return false;
  switch (re->op) {
    case OpConcat:
    case OpCapture:
      return !re->sub.empty() && isDigitRunSkipSafe(re->sub[0]);
    case OpPlus:
    case OpStar:
      return re->sub.size() == 1 && re->sub[0]->op == OpCharClass && isDigitOnlyClass(re->sub[0]->rune);
    case OpRepeat:
      return re->max == -1 && re->sub.size() == 1 && re->sub[0]->op == OpCharClass &&
             isDigitOnlyClass(re->sub[0]->rune);
    default: return false;
  }
}

static bool anyOp(const Regexp* re, bool (*pred)(Op)) {
  if (pred(re->op)) return true;
  for (auto* s : re->sub)
    if (anyOp(s, pred)) return true;
  return false;
}
static bool hasAnchorAssertions(const Regexp* re) {
  return anyOp(re, [](Op o) {
    return o == OpBeginLine || o == OpEndLine || o == OpBeginText || o == OpEndText ||
           o == OpWordBoundary || o == OpNoWordBoundary;
  });
}
static bool hasNonLineAnchors(const Regexp* re) {
  return anyOp(re, [](Op o) {
    return o == OpEndLine || o == OpEndText || o == OpBeginText || o == OpWordBoundary ||
           o == OpNoWordBoundary;
  });
}
static bool hasWordBoundary(const Regexp* re) {
  return anyOp(re, [](Op o) { return o == OpWordBoundary || o == OpNoWordBoundary; });
}
static bool hasMultilineLineAnchor(const Regexp* re) {
  return anyOp(re, [](Op o) { return o == OpBeginLine || o == OpEndLine; });
}
static bool containsEndAnchor(const Regexp* re) {
  return anyOp(re, [](Op o) { return o == OpEndText; });
}
static bool containsStartAnchor(const Regexp* re) {
  return anyOp(re, [](Op o) { return o == OpBeginText; });
}

// reference meta/strategy.go:1210-1254
static bool metaCanMatchEmpty(const Regexp* re) {
  switch (re->op) {
    case OpEmptyMatch: case OpNoMatch: return true;
    case OpLiteral: case OpCharClass: case OpAnyCharNotNL: case OpAnyChar: return false;
    case OpBeginLine: case OpEndLine: case OpBeginText: case OpEndText:
    case OpWordBoundary: case OpNoWordBoundary: return true;
    case OpCapture: return metaCanMatchEmpty(re->sub[0]);
    case OpStar: case OpQuest: return true;
    case OpPlus: return metaCanMatchEmpty(re->sub[0]);
    case OpRepeat: return re->min == 0 || metaCanMatchEmpty(re->sub[0]);
    case OpConcat:
      for (auto* s : re->sub) if (!metaCanMatchEmpty(s)) return false;
      return true;
    case OpAlternate:
      for (auto* s : re->sub) if (metaCanMatchEmpty(s)) return true;
      return false;
    default: return false;
  }
}

// reference meta/strategy.go:1100-1131
static bool isSimpleCharClass(const Regexp* re) {
  switch (re->op) {
    case OpCharClass: return true;
    case OpPlus: case OpStar: case OpQuest: case OpRepeat:
      return re->sub.size() == 1 && isSimpleCharClass(re->sub[0]);
    case OpConcat:
      for (auto* s : re->sub) if (!isSimpleCharClass(s)) return false;
      return true;
    case OpCapture: return re->sub.size() == 1 && isSimpleCharClass(re->sub[0]);
    default: return false;
  }
}

static bool isEndAnchoredTail(const Regexp* re) {
  switch (re->op) {
    case OpEndText: return true;
    case OpConcat: return !re->sub.empty() && isEndAnchoredTail(re->sub.back());
    case OpCapture: return !re->sub.empty() && isEndAnchoredTail(re->sub[0]);
    default: return false;
  }
}

Engine::~Engine() = default;

struct EngineBuilder {
  // Returns the strategy; sets exact=false when an un-restated predicate was needed.
  static int Select(Engine& e, const Seq& literals, bool& exact) {
    const Regexp* re = e.re_;
    const NFA& n = e.nfa_;
    exact = true;
    bool isStartAnchored = n.anchored;
    bool isEndAnchored = isEndAnchoredTail(re);  // approximation of nfa.IsPatternEndAnchored
    bool hasStartAnchor = containsStartAnchor(re);
    if (isEndAnchored && !isStartAnchored && !hasStartAnchor) {
      exact = false;  // internal-end-anchor check not restated
      return UseReverseAnchored;
    }
    if (isStartAnchored) {
      exact = false;
      return UseBoundedBacktracker;
    }
    // ---- selectReverseStrategy (reference meta/strategy.go:974-1083) ----
    {
      int rs = revsearch::SelectReverseStrategy(re, n, literals, exact);
      if (rs != 0) return rs;
    }
    int nfaSize = (int)n.states.size();
    bool hasGood = false, hasTeddy = false, hasAC = false;
    if (!literals.empty()) {
      if (literals.longest_common_prefix().size() >= 1) hasGood = true;
      size_t cnt = literals.len();
      if (cnt >= 2 && cnt <= 64) {
        bool all = true;
        for (auto& l : literals.lits)
          if (l.bytes.size() < 3) all = false;
        if (all) hasTeddy = true;
      }
      if (cnt > 64) {
        bool all = true;
        for (auto& l : literals.lits)
          if (l.bytes.empty()) all = false;
        if (all) hasAC = true;
      }
    }
    bool hasAnchors = hasAnchorAssertions(re);
    bool hasNonLine = hasAnchors && hasNonLineAnchors(re);
    if (!hasGood && !hasTeddy) {
      // CharClassSearcher / CompositeSearcher / BoundedBacktracker: engines not restated;
      // all are asserted equal to leftmost-first by the reference's own tests.
      if (isSimpleCharClass(re)) {
        exact = false;
        bool single = re->op == OpPlus || (re->op == OpCapture);
        return single ? UseCharClassSearcher : UseBoundedBacktracker;
      }
    }
    if (!literals.empty()) {
      if (hasTeddy && literals.all_complete() && !hasNonLine) return UseTeddy;
      if (hasAC && literals.all_complete()) {
        exact = false;  // external ahocorasick v0.3.0: arithmetic not in tree
        return UseAhoCorasick;
      }
    }
    if (nfaSize <= 100 && isDigitLeadPattern(re)) return UseDigitPrefilter;
    if (nfaSize < 20) {
      bool wbcombo = hasWordBoundary(re) && hasAnchors;
      if (wbcombo || metaCanMatchEmpty(re) || hasMultilineLineAnchor(re)) return UseNFA;
      return UseDFA;
    }
    if (!hasGood && !hasTeddy && metaCanMatchEmpty(re)) return UseNFA;
    if (hasGood || hasTeddy) {
      if (nfaSize > 200 && !literals.all_complete()) return UseNFA;
      return UseDFA;
    }
    if (nfaSize > 100) return UseNFA;
    return UseBoth;
  }
};

std::unique_ptr<Engine> Engine::Compile(const std::string& pattern, std::string& err) {
  std::unique_ptr<Engine> e(new Engine());
  ParseResult pr = Parse(pattern, Perl, e->arena_);
  if (!pr.re) {
    err = pr.err;
    return nullptr;
  }
  e->re_ = pr.re;
  std::string nerr = CompileNFA(pr.re, false, e->nfa_);
  if (!nerr.empty()) {
    err = nerr;
    return nullptr;
  }
  e->can_match_empty_ = metaCanMatchEmpty(pr.re);
  Seq literals;
  if (!e->nfa_.anchored) literals = ExtractPrefixes(pr.re);
  bool exact = true;
  e->strategy_ = EngineBuilder::Select(*e, literals, exact);
  e->strategy_exact_ = exact;
  e->pikevm_.reset(new PikeVM(&e->nfa_));

  switch (e->strategy_) {
    case UseDigitPrefilter:
      e->dfa_.reset(new LazyDFA(&e->nfa_, LazyConfig{true}));
      e->digit_run_skip_safe_ = isDigitRunSkipSafe(pr.re);
      break;
    case UseTeddy: {
      std::vector<std::string> pats;
      for (auto& l : literals.lits) pats.push_back(l.bytes);
      if (pats.size() <= 32) {
        e->teddy_.reset(new Teddy(pats));
        if (!e->teddy_->ok()) e->teddy_.reset();
      } else {
        e->fat_teddy_.reset(new FatTeddy(pats));
        if (!e->fat_teddy_->ok()) e->fat_teddy_.reset();
      }
      if (!e->teddy_ && !e->fat_teddy_) {
        e->strategy_ = UseNFA;
        e->strategy_exact_ = false;
      }
      // reference meta/compile.go:663-686 adjustForAnchors: (?m)^ + complete literals and no
      // other anchors -> WrapLineAnchor (the wrapper has no FindMatch, so findIndicesTeddyAt
      // takes its Find + LiteralLen / NFA branch)
      if (hasAnchorAssertions(pr.re) && hasMultilineLineAnchor(pr.re) && !hasNonLineAnchors(pr.re)) {
        e->teddy_line_anchor_ = true;
        size_t ul = pats[0].size();
        for (auto& p : pats)
          if (p.size() != ul) ul = 0;
        e->teddy_uniform_len_ = (int64_t)ul;
      }
      break;
    }
    case UseDFA:
      // reference meta/compile.go:160-219: forward DFA + reverse DFA of the ReverseAnchored NFA
      // (BreakAtMatch=false); non-greedy patterns get no reverse DFA and take PikeVM bounds.  The
      // prefilter skip-ahead of find_indices.go:342-354 only moves the start of the same search.
      if (revsearch::BuildBidirectional(e->re_, e->nfa_, e->rev_nfa_)) {
        e->dfa_.reset(new LazyDFA(&e->nfa_, LazyConfig{true}));
        e->rev_dfa_.reset(new LazyDFA(&e->rev_nfa_, LazyConfig{false}));
      }
      // the bidirectional search is built and tested (set_bidirectional) but not the default path:
      // by default UseDFA is pinned to leftmost-first through the PikeVM restatement
      e->strategy_exact_ = false;
      break;
    case UseBoth:
      // reference meta/compile.go:196 builds a reverse DFA for UseDFA only; the adaptive searcher
      // (find_indices.go:406-460, PikeVM bounds from an estimated start) is not restated
      e->strategy_exact_ = false;
      break;
    case UseReverseInner:
      e->rinner_ = revsearch::BuildReverseInner(e->re_, e->nfa_);
      if (!e->rinner_) {
        e->strategy_ = UseNFA;
        e->strategy_exact_ = false;
      }
      break;
    default:
      break;
  }
  return e;
}

// reference meta/find_indices.go:1050-1088
bool Engine::findDigitPrefilterAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e) {
  if (at >= n) return findNFAAt(h, n, at, s, e);
  int64_t pos = at;
  while (pos < n) {
    int64_t d = memchr_digit_at(h, n, pos);
    if (d < 0) return false;
    int64_t end = dfa_->SearchAtAnchored(h, n, d);
    if (end != -1) {
      s = d;
      e = end;
      return true;
    }
    pos = d + 1;
    if (digit_run_skip_safe_)
      while (pos < n && h[pos] >= '0' && h[pos] <= '9') pos++;
  }
  return false;
}

// reference meta/find_indices.go:1172-1212 (no prefilter, no backtracker in the pooled state)
bool Engine::findNFAAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e) {
  return pikevm_->SearchAt(h, n, at, s, e);
}

// reference meta/find_indices.go:925-951
bool Engine::findTeddyAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e) {
  if (at >= n) return findNFAAt(h, n, at, s, e);
  if (teddy_line_anchor_) {
    // reference prefilter/wrap.go:48-62 (lineAnchorWrapper.Find) + meta/find_indices.go:940-950
    int64_t pos = at;
    for (;;) {
      int64_t c = teddy_ ? teddy_->Find(h, n, pos) : fat_teddy_->Find(h, n, pos);
      if (c == -1) return false;
      if (c == 0 || h[c - 1] == '\n') {
        if (teddy_uniform_len_ > 0) {
          s = c;
          e = c + teddy_uniform_len_;
          return true;
        }
        return findNFAAt(h, n, c, s, e);
      }
      pos = c + 1;
    }
  }
  if (teddy_) return teddy_->FindMatch(h, n, at, s, e);
  return fat_teddy_->FindMatch(h, n, at, s, e);
}

// reference meta/findall.go:216-239 (useDFADirect) restated as a per-call helper
bool Engine::findDFAAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e) {
  int64_t end = dfa_->SearchAt(h, n, at);
  if (end < 0) return false;
  if (end == at) {
    s = e = at;
    return true;
  }
  if (nfa_.anchored) {  // reference meta/find_indices.go:697-700: no reverse search when always anchored
    s = at;
    e = end;
    return true;
  }
  int64_t st = rev_dfa_->SearchReverse(h, n, at, end);
  if (st < 0) return false;
  s = st;
  e = end;
  return true;
}

bool Engine::FindIndicesAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e) {
  if (at > 0 && nfa_.anchored) return false;
  switch (strategy_) {
    case UseDigitPrefilter: return findDigitPrefilterAt(h, n, at, s, e);
    case UseTeddy: return findTeddyAt(h, n, at, s, e);
    case UseDFA:
    case UseBoth:
      if (use_bidir_ && strategy_ == UseDFA && dfa_ && rev_dfa_) return findDFAAt(h, n, at, s, e);
      return findNFAAt(h, n, at, s, e);
    case UseReverseInner:
      return revsearch::ReverseInnerFindAt(*rinner_, *pikevm_, h, n, at, s, e);
    default: return findNFAAt(h, n, at, s, e);
  }
}

// reference meta/findall.go:176-290
int64_t Engine::FindAll(const uint8_t* h, int64_t n, int64_t limit, std::vector<int64_t>& out) {
  int64_t count = 0;
  int64_t pos = 0, lastMatchEnd = -1;
  if (nfa_.anchored) {
    int64_t s, e;
    if (FindIndicesAt(h, n, 0, s, e)) {
      out.push_back(s);
      out.push_back(e);
      return 1;
    }
    return 0;
  }
  while (limit <= 0 || count < limit) {
    int64_t s, e;
    if (!FindIndicesAt(h, n, pos, s, e)) break;
    if (s == e && s == lastMatchEnd) {
      pos++;
      if (pos > n) break;
      continue;
    }
    out.push_back(s);
    out.push_back(e);
    count++;
    if (s != e) lastMatchEnd = e;
    if (s == e)
      pos = e + 1;
    else if (e > pos)
      pos = e;
    else
      pos++;
    if (pos > n) break;
  }
  return count;
}

// reference meta/findall.go:297-380
int64_t Engine::Count(const uint8_t* h, int64_t n, int64_t limit) {
  if (limit == 0) return 0;
  int64_t count = 0, pos = 0, lastNonEmptyEnd = -1;
  while (pos <= n) {
    int64_t s, e;
    if (!FindIndicesAt(h, n, pos, s, e)) break;
    if (s == e && s == lastNonEmptyEnd) {
      pos++;
      if (pos > n) break;
      continue;
    }
    count++;
    if (s != e) lastNonEmptyEnd = e;
    if (s == e)
      pos = e + 1;
    else if (e > pos)
      pos = e;
    else
      pos++;
    if (limit > 0 && count >= limit) break;
  }
  return count;
}

// reference meta/ismatch.go:27 — every restated strategy reduces to "a match exists"
bool Engine::IsMatch(const uint8_t* h, int64_t n) {
  int64_t s, e;
  return FindIndicesAt(h, n, 0, s, e);
}

// reference meta/findall.go:63-133, :390-447
int64_t Engine::FindAllSubmatch(const uint8_t* h, int64_t n, int64_t limit, std::vector<int64_t>& out) {
  if (limit == 0) return 0;
  int64_t count = 0, pos = 0, lastMatchEnd = -1;
  int stride = nfa_.capture_count * 2;
  std::vector<int64_t> slots;
  while (pos <= n) {
    bool found = false;
    switch (strategy_) {
      case UseBoundedBacktracker: case UseNFA: case UseDFA: case UseBoth: case UseDigitPrefilter:
        found = pikevm_->SearchCapturesAt(h, n, pos, slots);
        break;
      default: {
        int64_t s, e;
        if (!FindIndicesAt(h, n, pos, s, e)) break;
        if (nfa_.capture_count <= 1) {
          slots.assign(stride, -1);
          slots[0] = s;
          slots[1] = e;
          found = true;
        } else {
          // reference: SearchWithCapturesInSpan(h, start, end) then fallback.  The span search is
          // an anchored capture search on [start,end): restated via captures-at from `s`.
          found = pikevm_->SearchCapturesAt(h, n, s, slots);
        }
      }
    }
    if (!found) break;
    int64_t ms = slots[0], me = slots[1];
    if (ms == me && ms == lastMatchEnd) {
      pos++;
      if (pos > n) break;
      continue;
    }
    out.insert(out.end(), slots.begin(), slots.begin() + stride);
    count++;
    if (ms != me) lastMatchEnd = me;
    if (ms == me)
      pos = me + 1;
    else if (me > pos)
      pos = me;
    else
      pos++;
    if (limit > 0 && count >= limit) break;
  }
  return count;
}

}  // namespace oracle
