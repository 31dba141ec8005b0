// oracle/revsearch.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// What IS restated here: the SELECTION of the reference's reverse strategies
//   reference meta/strategy.go:974-1083 (selectReverseStrategy) and its predicates :565-960
//   reference literal/extractor.go:1010-1180 (ExtractInnerForReverseSearch, buildPrefix/SuffixAST)
// so that Oracle.strategy names what the reference would pick (tests/test_oracle_golden.py
// test_strategy_table, tests/test_host_compile.py test_reference_strategy_agrees_with_oracle).
//
// What is NOT restated: the reverse SEARCHERS themselves —
//   reference nfa/reverse.go:8-330 (Reverse / ReverseAnchored), meta/reverse_inner.go:95-190, :522-592,
//   meta/reverse_suffix.go, meta/compile.go:185-219 (buildReverseDFA for UseDFA/UseBoth).
// BuildBidirectional / BuildReverseInner / ReverseNFA below are deliberate NO-OPS that make the
// engine fall back to the PikeVM restatement (leftmost-first) and clear `strategy_exact`: for
// UseDFA / UseBoth / UseReverse* the oracle pins "what stdlib leftmost-first yields" — which the
// reference's own tests assert those searchers equal — not the searchers' code paths.  PARITY FOR
// THOSE STRATEGIES IS THEREFORE PINNED TO LEFTMOST-FIRST SEMANTICS, NOT TO THE REFERENCE ENGINES
// (DESIGN.md §3).
#pragma once
#include <memory>

#include "../syntax/syntax.h"
#include "lazydfa.h"
#include "literal.h"
#include "nfa.h"
#include "pikevm.h"

namespace oracle {

struct ReverseInner {
  gosyntax::Arena arena;
  NFA prefix_nfa, rev_nfa, suffix_nfa;
  std::unique_ptr<LazyDFA> rev_dfa, fwd_dfa;
  std::string inner;  // the single inner literal (LCP of the inner literal set)
  std::vector<std::string> inner_set;
  bool universal_prefix = false, universal_suffix = false;
};

namespace revsearch {

// 0 = no reverse strategy; otherwise a Strategy value.  `exact` is cleared when the selected
// strategy's search engine is not restated here.
int SelectReverseStrategy(const gosyntax::Regexp* re, const NFA& n, const Seq& prefix_literals,
                          bool& exact);

// reverse NFA for bidirectional UseDFA/UseBoth search; false when the reference builds none
bool BuildBidirectional(const gosyntax::Regexp* re, const NFA& fwd, NFA& rev_out);

std::unique_ptr<ReverseInner> BuildReverseInner(const gosyntax::Regexp* re, const NFA& full);
bool ReverseInnerFindAt(ReverseInner& ri, PikeVM& pikevm, const uint8_t* h, int64_t n, int64_t at,
                        int64_t& s, int64_t& e);

// reference nfa/reverse.go:38 (anchored=false) / :8 (anchored=true)
void ReverseNFA(const NFA& fwd, bool anchored, NFA& out);

}  // namespace revsearch
}  // namespace oracle
