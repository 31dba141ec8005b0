// oracle/revsearch.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// Restated here:
//   * the SELECTION of the reference's reverse strategies:
//     reference meta/strategy.go:974-1083 (selectReverseStrategy) and its predicates :565-960,
//     literal/extractor.go:1010-1180 (ExtractInnerForReverseSearch, buildPrefix/SuffixAST);
//   * the bidirectional search of UseDFA: reference nfa/reverse.go:8-634 (ReverseAnchored; the
//     states are built in oracle/nfa.cpp ReverseNFAStates), meta/compile.go:176-219 (buildReverseDFA:
//     only UseDFA, not for non-greedy patterns), meta/find_indices.go:686-705 (forward lazy DFA for
//     the end, reverse lazy DFA for the start; oracle/meta.cpp findDFAAt over lazydfa.cpp SearchAt /
//     SearchReverse).  Exercised by tests/test_oracle_golden.py::test_bidirectional_* against the
//     PikeVM restatement and Python `re`.  It is NOT the oracle's default path for UseDFA: restated
//     as written, the reverse automaton of a pattern that opens with a star loop loses the loop
//     (test_reverse_nfa_of_a_leading_star_quirk), which cannot be confirmed without running the
//     reference; the default stays leftmost-first through the PikeVM (Engine::set_bidirectional).
//
// NOT restated: the ReverseInner / ReverseSuffix / ReverseAnchored searchers and the adaptive
// UseBoth searcher (reference meta/reverse_inner.go:95-190, :522-592, meta/reverse_suffix.go,
// meta/find_indices.go:406-460).  BuildReverseInner / ReverseInnerFindAt below are deliberate
// no-ops that send the engine to the PikeVM restatement and clear `strategy_exact`: for those
// strategies the oracle pins "what leftmost-first yields" — which the reference's own tests assert
// those searchers equal — not their code paths.
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
