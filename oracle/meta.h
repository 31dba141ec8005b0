// oracle/meta.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// CPU restatement of the reference's meta engine for the bulk-scan hot path:
//   reference meta/compile.go:40-60,440-654 (Compile pipeline), :115-219 (buildStrategyEngines)
//   reference meta/strategy.go:303-560 (digit-lead analysis), :1131-1308 (literal analysis),
//     :1377-1546 (SelectStrategy)
//   reference meta/findall.go:155-290 (FindAllIndicesStreaming / findAllIndicesLoop), :297-380 (Count)
//   reference meta/find_indices.go:1050-1088 (DigitPrefilter loop), :925-951 (Teddy),
//     :1127-1170 (dispatcher), :1172-1212 (NFA path)
//   reference meta/ismatch.go:27,286-311 (IsMatch, isMatchDigitPrefilter)
//
// Strategy coverage: the strategies the BASELINE.json configs select are restated exactly
// (UseDigitPrefilter, UseTeddy incl. Fat Teddy, UseBoth/UseNFA via PikeVM, UseDFA bidirectional,
// UseReverseInner).  For any other strategy `strategy_exact` is false and the search runs the
// PikeVM restatement (leftmost-first semantics; the reference's own tests assert every strategy
// equals Go stdlib leftmost-first).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../syntax/syntax.h"
#include "lazydfa.h"
#include "nfa.h"

namespace oracle {

enum Strategy : int {
  UseNFA = 0,
  UseDFA,
  UseBoth,
  UseReverseAnchored,
  UseReverseSuffix,
  UseOnePass,
  UseReverseInner,
  UseBoundedBacktracker,
  UseTeddy,
  UseReverseSuffixSet,
  UseCharClassSearcher,
  UseCompositeSearcher,
  UseBranchDispatch,
  UseDigitPrefilter,
  UseAhoCorasick,
  UseAnchoredLiteral,
  UseMultilineReverseSuffix,
};
const char* StrategyName(int s);

class PikeVM;
class Teddy;
class FatTeddy;
struct ReverseInner;
class Literals;

class Engine {
 public:
  ~Engine();
  static std::unique_ptr<Engine> Compile(const std::string& pattern, std::string& err);

  bool IsMatch(const uint8_t* h, int64_t n);
  // appends (start,end) pairs; limit<=0 means all (reference n<0); returns count
  int64_t FindAll(const uint8_t* h, int64_t n, int64_t limit, std::vector<int64_t>& out);
  int64_t Count(const uint8_t* h, int64_t n, int64_t limit);
  // stride = 2*(NumSubexp+1); unmatched groups are -1,-1 (reference regex.go:1423)
  int64_t FindAllSubmatch(const uint8_t* h, int64_t n, int64_t limit, std::vector<int64_t>& out);
  // reference meta/find_indices.go:61
  bool FindIndicesAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e);

  int strategy() const { return strategy_; }
  bool strategy_exact() const { return strategy_exact_; }
  // UseDFA: run the restated bidirectional search (forward lazy DFA for the end, lazy DFA of the
  // ReverseAnchored NFA for the start: meta/find_indices.go:686-705, nfa/reverse.go) instead of the
  // PikeVM.  OFF by default — see DESIGN.md §3 "reverse NFA of a leading star".
  bool has_bidirectional() const { return dfa_ && rev_dfa_ && strategy_ == UseDFA; }
  void set_bidirectional(bool on) { use_bidir_ = on; }
  int num_captures() const { return nfa_.capture_count; }
  const NFA& nfa() const { return nfa_; }
  const gosyntax::Regexp* ast() const { return re_; }
  bool digit_run_skip_safe() const { return digit_run_skip_safe_; }

 private:
  Engine() = default;
  gosyntax::Arena arena_;
  const gosyntax::Regexp* re_ = nullptr;
  NFA nfa_;
  int strategy_ = UseNFA;
  bool strategy_exact_ = true;
  bool use_bidir_ = false;
  bool digit_run_skip_safe_ = false;
  bool can_match_empty_ = false;
  bool teddy_line_anchor_ = false;  // prefilter wrapped by WrapLineAnchor (meta/compile.go:663-686)
  int64_t teddy_uniform_len_ = 0;
  std::unique_ptr<LazyDFA> dfa_;      // forward
  std::unique_ptr<LazyDFA> rev_dfa_;  // reverse (UseDFA / UseBoth bidirectional)
  NFA rev_nfa_;
  std::unique_ptr<PikeVM> pikevm_;
  std::unique_ptr<Teddy> teddy_;
  std::unique_ptr<FatTeddy> fat_teddy_;
  std::unique_ptr<ReverseInner> rinner_;

  bool findDigitPrefilterAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e);
  bool findNFAAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e);
  bool findTeddyAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e);
  bool findDFAAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e);
  friend struct EngineBuilder;
};

// AST predicates shared with tests (reference meta/strategy.go)
bool isDigitLeadPattern(const gosyntax::Regexp* re);
bool isDigitRunSkipSafe(const gosyntax::Regexp* re);

}  // namespace oracle
