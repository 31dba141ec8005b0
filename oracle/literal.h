// oracle/literal.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// CPU restatement of the reference's prefix-literal extraction (what decides literal ORDER,
// hence Teddy bucket assignment and tie-breaks):
//   reference literal/extractor.go:128-152 (ExtractPrefixes incl. >64 trimming cascade),
//     :155-238 (extractPrefixes op switch), :240-300 (alternate), :302-365 (concat cross product),
//     :376-440 (concatSubContribution), :450-485 (expandAlternateContribution),
//     :487-560 (overflow helpers), :836-960 (case-fold expansion), :963-1003 (expandCharClass)
//   reference literal/seq.go:206 (AllComplete), :343 (LongestCommonPrefix), :433 (CrossForward),
//     :470 (KeepFirstBytes), :491 (Dedup)
// Config as used by meta: MaxLiterals=256, MaxLiteralLen=64, MaxClassSize=10,
// CrossProductLimit=250 (reference meta/compile.go:467-471, literal/extractor.go:241-244).
// Suffix / inner extraction (ReverseSuffix/ReverseInner strategy selection) is restated only to
// the extent needed to classify the config patterns: see meta.cpp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../syntax/syntax.h"

namespace oracle {

struct Literal {
  std::string bytes;
  bool complete = true;
};

struct Seq {
  std::vector<Literal> lits;
  bool partial_coverage = false;
  bool empty() const { return lits.empty(); }
  size_t len() const { return lits.size(); }
  bool all_complete() const;
  std::string longest_common_prefix() const;
  std::string longest_common_suffix() const;
  void cross_forward(const Seq& other);
  void keep_first_bytes(size_t n);
  void dedup();
  size_t min_len() const;
};

struct ExtractorConfig {
  int max_literals = 256;
  int max_literal_len = 64;
  int max_class_size = 10;
  int cross_product_limit = 250;
};

Seq ExtractPrefixes(const gosyntax::Regexp* re, const ExtractorConfig& cfg = ExtractorConfig());
Seq ExtractSuffixes(const gosyntax::Regexp* re, const ExtractorConfig& cfg = ExtractorConfig());

}  // namespace oracle
