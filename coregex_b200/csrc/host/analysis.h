// analysis.h — host-side pattern analysis: which reference strategy a pattern would get, the
// prefix literal list (order matters: it fixes Teddy bucket assignment), digit-lead facts.
//
// Mirrors the DECISIONS of reference meta/strategy.go:1377-1546 (SelectStrategy) and
// meta/compile.go:440-654 for the strategies the GPU engines implement; everything else is
// reported under its reference name and executed with plain leftmost-first semantics (which is
// what every reference strategy is tested to equal).
#pragma once
#include <string>
#include <vector>

#include "../../../syntax/syntax.h"
#include "prog.h"

namespace cgx {

enum RefStrategy : int {
  RS_UseNFA = 0, RS_UseDFA, RS_UseBoth, RS_UseReverseAnchored, RS_UseReverseSuffix, RS_UseOnePass,
  RS_UseReverseInner, RS_UseBoundedBacktracker, RS_UseTeddy, RS_UseReverseSuffixSet,
  RS_UseCharClassSearcher, RS_UseCompositeSearcher, RS_UseBranchDispatch, RS_UseDigitPrefilter,
  RS_UseAhoCorasick, RS_UseAnchoredLiteral, RS_UseMultilineReverseSuffix,
};
const char* RefStrategyName(int s);

struct Lit {
  std::string bytes;
  bool complete = true;
};

struct Analysis {
  int strategy = RS_UseNFA;
  std::vector<Lit> prefixes;       // reference literal.ExtractPrefixes order
  bool prefixes_all_complete = false;
  bool digit_lead = false;
  bool digit_run_skip_safe = false;
  bool can_match_empty = false;
  bool has_anchors = false;
  // ReverseInner split (index of the inner literal in the top-level concat), -1 if none
  int inner_idx = -1;
  std::string inner_literal;
};

// the part of reference meta.Config (meta/config.go:31-113) that strategy selection reads
// (meta/strategy.go:515,963,976,1265,1392,1447; meta/compile.go:466)
struct AnalysisConfig {
  bool enable_dfa = true;
  bool enable_prefilter = true;
  int min_literal_len = 1;
};

Analysis Analyze(const gosyntax::Regexp* re, int prog_size_hint, bool anchored_start,
                 const AnalysisConfig& cfg = AnalysisConfig());

}  // namespace cgx
