// oracle/revsearch.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// CPU restatement of the reference's reverse-search pieces:
//   reference meta/strategy.go:974-1083 (selectReverseStrategy) and its predicates :565-960
//   reference literal/extractor.go:1010-1180 (ExtractInnerForReverseSearch, buildPrefix/SuffixAST)
//   reference nfa/reverse.go:8-330 (Reverse / ReverseAnchored)
//   reference meta/reverse_inner.go:95-190 (NewReverseInnerSearcher), :522-592 (findIndicesAtImpl)
//   reference dfa/lazy/lazy.go:1947-2035 (SearchReverseLimited)
//   reference meta/compile.go:185-219 (buildReverseDFA for UseDFA/UseBoth)
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
