// oracle/revsearch.cpp — TEST INFRASTRUCTURE ONLY.  See revsearch.h for the reference citations.
#include "revsearch.h"

#include "meta.h"
#include "simd.h"

namespace oracle {
namespace revsearch {

using namespace gosyntax;

namespace {

bool anyOp(const Regexp* re, bool (*pred)(Op)) {
  if (pred(re->op)) return true;
  for (auto* s : re->sub)
    if (anyOp(s, pred)) return true;
  return false;
}

bool isWildcardOrRepetition(const Regexp* re) {
  switch (re->op) {
    case OpStar: case OpPlus: case OpQuest: case OpRepeat: return true;
    case OpAnyChar: case OpAnyCharNotNL: return true;
    case OpConcat: case OpAlternate:
      for (auto* s : re->sub) if (isWildcardOrRepetition(s)) return true;
      return false;
    case OpCapture: return !re->sub.empty() && isWildcardOrRepetition(re->sub[0]);
    default: return false;
  }
}

bool isWildcardSubexpression(const Regexp* re) {
  while (re->op == OpCapture && !re->sub.empty()) re = re->sub[0];
  if ((re->op == OpStar || re->op == OpPlus) && !re->sub.empty() &&
      (re->sub[0]->op == OpAnyChar || re->sub[0]->op == OpAnyCharNotNL))
    return true;
  if (re->op == OpPlus && !re->sub.empty() && re->sub[0]->op == OpCharClass) return true;
  if (re->op == OpRepeat && re->min >= 1) return true;
  return false;
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

bool isSafeForReverseSuffix(const Regexp* re) {
  switch (re->op) {
    case OpConcat: {
      if (re->sub.size() < 2) return false;
      int wc = 0;
      for (size_t i = 0; i + 1 < re->sub.size(); i++)
        if (isWildcardSubexpression(re->sub[i])) wc++;
      if (wc == 0) return false;
      for (size_t i = 1; i + 1 < re->sub.size(); i++)
        if (containsAnchor(re->sub[i])) return false;
      return true;
    }
    case OpCapture: return !re->sub.empty() && isSafeForReverseSuffix(re->sub[0]);
    default: return false;
  }
}

bool isSafeForReverseInner(const Regexp* re) {
  switch (re->op) {
    case OpConcat: {
      if (re->sub.size() < 2) return false;
      const Regexp* first = re->sub[0];
      if ((first->op == OpStar || first->op == OpPlus) && !first->sub.empty() &&
          (first->sub[0]->op == OpAnyChar || first->sub[0]->op == OpAnyCharNotNL))
        return true;
      if (first->op == OpPlus && !first->sub.empty() && first->sub[0]->op == OpCharClass) return true;
      return false;
    }
    case OpCapture: return !re->sub.empty() && isSafeForReverseInner(re->sub[0]);
    default: return false;
  }
}

// reference literal/extractor.go:693-760 (extractInner)
Seq extractInner(const Regexp* re, int depth) {
  ExtractorConfig cfg;
  if (depth > 100) return Seq();
  switch (re->op) {
    case OpLiteral: {
      if (re->flags & FoldCase) {
        // case-fold expansion, all marked incomplete: reuse prefix extraction of the literal
        Seq s = ExtractPrefixes(re, cfg);
        for (auto& l : s.lits) l.complete = false;
        return s;
      }
      Seq s = ExtractPrefixes(re, cfg);
      for (auto& l : s.lits) l.complete = false;
      return s;
    }
    case OpConcat:
      for (auto* sub : re->sub) {
        Seq s = extractInner(sub, depth + 1);
        if (!s.empty()) return s;
      }
      return Seq();
    case OpAlternate: {
      Seq all;
      for (auto* sub : re->sub) {
        Seq s = extractInner(sub, depth + 1);
        if (s.empty()) return Seq();
        for (auto& l : s.lits) {
          all.lits.push_back(l);
          if ((int)all.len() >= cfg.max_literals) return all;
        }
      }
      return all;
    }
    case OpCharClass: {
      // expandCharClass == prefix extraction of a bare class
      return ExtractPrefixes(re, cfg);
    }
    case OpCapture:
      if (re->sub.empty()) return Seq();
      return extractInner(re->sub[0], depth + 1);
    default:
      return Seq();
  }
}

struct InnerInfo {
  Seq literals;
  size_t idx = 0;
  bool ok = false;
};

InnerInfo extractInnerForReverseSearch(const Regexp* re) {
  InnerInfo info;
  if (re->op != OpConcat || re->sub.size() < 3) return info;
  for (size_t i = 1; i + 1 < re->sub.size(); i++) {
    Seq lits = extractInner(re->sub[i], 0);
    if (lits.empty()) continue;
    bool before = false, after = false;
    for (size_t j = 0; j < i; j++)
      if (isWildcardOrRepetition(re->sub[j])) before = true;
    for (size_t j = i + 1; j < re->sub.size(); j++)
      if (isWildcardOrRepetition(re->sub[j])) after = true;
    if (before && after) {
      info.literals = lits;
      info.idx = i;
      info.ok = true;
      return info;
    }
  }
  return info;
}

bool hasFastPrefixPrefilter(const Seq& lits) {
  if (lits.empty()) return false;
  if (lits.longest_common_prefix().size() >= 1) return true;
  // prefilter.WouldBeFast
  if (lits.len() == 1) return true;
  return lits.min_len() >= 3;
}

}  // namespace

int SelectReverseStrategy(const Regexp* re, const NFA& n, const Seq& prefix_literals, bool& exact) {
  // HasImpossibleEndAnchor: an EndText that is not in tail position (approximation: any EndText
  // inside a non-tail concat element)
  if (anyOp(re, [](Op o) { return o == OpWordBoundary || o == OpNoWordBoundary; })) return 0;
  if (n.anchored) return 0;
  if (anyOp(re, [](Op o) { return o == OpEndText; })) {
    // end-anchored patterns are routed earlier (UseReverseAnchored) or are "impossible" anchors
    return 0;
  }
  // multiline reverse suffix ((?m)^ ... wildcard ... suffix): engine not restated
  if (anyOp(re, [](Op o) { return o == OpBeginLine; }) &&
      anyOp(re, [](Op o) { return o == OpAnyChar || o == OpAnyCharNotNL; })) {
    exact = false;
  }
  if (hasFastPrefixPrefilter(prefix_literals)) return 0;

  Seq suf = ExtractSuffixes(re);
  if (!suf.empty()) {
    if (suf.longest_common_suffix().size() >= 1) {
      if (!isSafeForReverseSuffix(re)) return 0;
      exact = false;
      return UseReverseSuffix;
    }
  }
  // shouldUseReverseSuffixSet
  if (isSafeForReverseSuffix(re) && !suf.empty()) {
    bool exactAlt = !prefix_literals.empty() && prefix_literals.all_complete() &&
                    prefix_literals.len() == suf.len();
    size_t cnt = suf.len();
    bool ok = !exactAlt && cnt >= 2 && cnt <= 32;
    if (ok)
      for (auto& l : suf.lits)
        if (l.bytes.size() < 2) ok = false;
    if (ok) {
      exact = false;
      return UseReverseSuffixSet;
    }
  }
  InnerInfo info = extractInnerForReverseSearch(re);
  if (info.ok) {
    std::string lcp = info.literals.longest_common_prefix();
    if (lcp.size() == 1 && isDigitLeadPattern(re)) return 0;
    if (lcp.size() >= 1) {
      if (!isSafeForReverseInner(re)) return 0;
      return UseReverseInner;
    }
  }
  return 0;
}

// reference meta/strategy.go:691-716
static bool hasNonGreedyQuantifier(const Regexp* re) {
  switch (re->op) {
    case OpStar: case OpPlus: case OpQuest: case OpRepeat:
      if (re->flags & NonGreedy) return true;
      for (auto* s : re->sub)
        if (hasNonGreedyQuantifier(s)) return true;
      return false;
    case OpConcat: case OpAlternate: case OpCapture:
      for (auto* s : re->sub)
        if (hasNonGreedyQuantifier(s)) return true;
      return false;
    default:
      return false;
  }
}

// reference meta/compile.go:176-219 buildReverseDFA, case UseDFA: forward DFA for the match end,
// reverse DFA (ReverseAnchored NFA, BreakAtMatch off) for the match start; non-greedy patterns get
// none and keep the PikeVM for the bounds.
bool BuildBidirectional(const Regexp* re, const NFA& fwd, NFA& rev_out) {
  if (hasNonGreedyQuantifier(re)) return false;
  ReverseNFA(fwd, /*anchored=*/true, rev_out);
  return true;
}

// ---- NOT RESTATED (see revsearch.h): these no-ops send the engine to the PikeVM restatement ----

std::unique_ptr<ReverseInner> BuildReverseInner(const Regexp*, const NFA&) { return nullptr; }

bool ReverseInnerFindAt(ReverseInner&, PikeVM& pikevm, const uint8_t* h, int64_t n, int64_t at,
                        int64_t& s, int64_t& e) {
  return pikevm.SearchAt(h, n, at, s, e);
}

void ReverseNFA(const NFA& fwd, bool anchored, NFA& out) { ReverseNFAStates(fwd, anchored, out); }

}  // namespace revsearch
}  // namespace oracle
