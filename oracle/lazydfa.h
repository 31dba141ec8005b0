// oracle/lazydfa.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// CPU restatement of the reference's lazy (hybrid) DFA:
//   reference dfa/lazy/builder.go:183-293 (moveWithWordContextBreak, epsilonClosureInto),
//     :300-431 (resolveWordBoundaries), :454-472 (CheckEOIMatch)
//   reference dfa/lazy/lazy.go:219-324 (SearchAtAnchored), :1102-1315 (searchAt),
//     :1336-1446 (determinize: 1-byte match delay, break-at-match, dead-end match state),
//     :1569-1620 (getStartState), :1769-1920 (SearchReverse)
//   reference dfa/lazy/start.go:17-37,96-110,205-256 (start kinds), dfa/lazy/look.go:88-112
//   reference dfa/lazy/state.go:329-373 (state key = SORTED nfa ids + isFromWord + isMatch)
//
// The construction stays LAZY on purpose: the reference keys cached states by the sorted NFA
// set while keeping the first-seen insertion order, so which ordering a state carries depends
// on the order inputs were seen.  Restating it lazily reproduces that.
// Cache clearing / NFA fallback (lazy.go:1428-1436,1472-1502) only changes speed, not results,
// and is not restated (the oracle's cache is unbounded).
#pragma once
#include <cstdint>
#include <map>
#include <vector>

#include "nfa.h"

namespace oracle {

enum LookSetBits : uint32_t {
  LS_StartText = 2,
  LS_EndText = 4,
  LS_StartLine = 8,
  LS_EndLine = 16,
  LS_WordBoundary = 32,
  LS_NoWordBoundary = 64,
};

enum StartKind : uint8_t { StartNonWord = 0, StartWord, StartText, StartLineLF, StartLineCR, kStartKinds };

struct DState {
  std::vector<StateID> nfa;  // insertion order
  bool is_match = false;     // delayed match tag
  bool from_word = false;
  bool match_at_wb = false, match_at_nwb = false;
  // transitions live in LazyDFA::flat_ (state*stride + class): -2 unknown, -1 dead, else state
  // index — the flat layout of reference dfa/lazy/cache.go:44-50 so the hot loop is two loads
};

struct LazyConfig {
  bool break_at_match = true;  // forward DFAs; reverse DFAs use false (meta/compile.go:191-205)
};

class LazyDFA {
 public:
  LazyDFA(const NFA* nfa, LazyConfig cfg);

  // reference dfa/lazy/lazy.go:219
  int64_t SearchAtAnchored(const uint8_t* h, int64_t n, int64_t at);
  // reference dfa/lazy/lazy.go:190 (SearchAt -> searchAt :1102), prefilter-free
  int64_t SearchAt(const uint8_t* h, int64_t n, int64_t at);
  // reference dfa/lazy/lazy.go:1769
  int64_t SearchReverse(const uint8_t* h, int64_t n, int64_t start, int64_t end);
  // reference dfa/lazy/lazy.go:530/561 (IsMatch -> searchEarliestMatch), semantic core
  bool IsMatch(const uint8_t* h, int64_t n);

  size_t NumStates() const { return states_.size(); }
  bool has_word_boundary() const { return has_wb_; }

 private:
  using Key = std::pair<std::vector<StateID>, int>;  // (sorted ids, flags)
  const NFA* nfa_;
  LazyConfig cfg_;
  bool has_wb_ = false, has_endline_ = false;
  std::vector<DState> states_;
  std::vector<int32_t> flat_;
  int stride_ = 1;
  std::map<Key, int> cache_;
  int newState(DState&& s);
  int start_[2][kStartKinds];

  void closureInto(std::vector<StateID>& set, std::vector<uint8_t>& in_set, StateID seed,
                   uint32_t look_have) const;
  std::vector<StateID> closure(const std::vector<StateID>& seeds, uint32_t look_have) const;
  std::vector<StateID> resolveWB(const std::vector<StateID>& states, bool satisfied) const;
  std::vector<StateID> move(const std::vector<StateID>& states, uint8_t b, bool from_word,
                            bool break_at_match) const;
  bool containsMatch(const std::vector<StateID>& s) const;
  int determinize(int cur, uint8_t b);  // returns next state index or -1 (dead)
  int getStart(const uint8_t* h, int64_t pos, bool anchored);
  bool checkEOI(int sid) const;
  bool matchesEmpty();
  bool wbFast(const DState& st, uint8_t b) const;
  static Key makeKey(const std::vector<StateID>& ids, bool from_word, bool is_match);
  int step(int sid, uint8_t b);
};

}  // namespace oracle
