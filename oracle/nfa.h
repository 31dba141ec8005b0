// oracle/nfa.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// CPU restatement of the reference's Thompson NFA model, builder, byte-class computation and
// compiler, following:
//   reference nfa/nfa.go:21-60 (state kinds), :120-146 (State), :287-325 (NFA)
//   reference nfa/builder.go:40-260 (Add*/Patch), nfa/alphabet.go:111-166 (ByteClassSet)
//   reference nfa/compile.go:99-165 (CompileRegexp), :167-233 (op switch), :237-361 (literals),
//     :384-437 (ASCII char class), :1225-1311 (concat/alternate/split chain),
//     :1313-1483 (star/plus/quest, canMatchEmpty), :1485-1565 (repeat), :1633-1689 (prefix, capture)
//
// Parity status: pinned by the reference's own known-answer vectors replayed in
// tests/test_oracle_golden.py.  The reference itself is Go and
// cannot be executed in this image (no Go toolchain), so vectors + Python `re` differential
// tests are the pin.
//
// Scope: ASCII classes only.  Non-ASCII classes / `.` (which the reference compiles into UTF-8
// automata, nfa/compile.go:440-1223) are rejected with "unsupported" (SURVEY.md §2.1 row 10).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../syntax/syntax.h"

namespace oracle {

using StateID = uint32_t;
constexpr StateID InvalidState = 0xFFFFFFFFu;

enum StateKind : uint8_t {
  StateMatch = 0,
  StateByteRange,
  StateSparse,
  StateSplit,
  StateEpsilon,
  StateCapture,
  StateFail,
  StateLook,
};

enum Look : uint8_t {
  LookStartText = 0,
  LookEndText,
  LookStartLine,
  LookEndLine,
  LookWordBoundary,
  LookNoWordBoundary,
};

struct Transition {
  uint8_t lo, hi;
  StateID next;
};

struct State {
  StateKind kind = StateFail;
  uint8_t lo = 0, hi = 0;                // ByteRange
  StateID next = InvalidState;           // ByteRange / Epsilon / Capture / Look
  StateID left = InvalidState, right = InvalidState;  // Split
  bool quantifier_split = false;
  std::vector<Transition> trans;         // Sparse
  uint32_t cap_index = 0;                // Capture
  bool cap_start = false;
  Look look = LookStartText;
};

struct NFA {
  std::vector<State> states;
  StateID start_anchored = InvalidState, start_unanchored = InvalidState;
  bool anchored = false;      // pattern is inherently ^-anchored (IsAlwaysAnchored)
  int capture_count = 1;      // groups incl. group 0
  std::vector<std::string> capture_names;
  uint8_t byte_classes[256];
  int alphabet_len = 1;

  bool is_match(StateID s) const { return s < states.size() && states[s].kind == StateMatch; }
};

// Returns "" on success, else the error text.
std::string CompileNFA(const gosyntax::Regexp* re, bool anchored_cfg, NFA& out);

std::string DumpNFA(const NFA& n);

// reference nfa/reverse.go:8-81 (ReverseAnchored / Reverse -> reverseWithOptions) and the helpers it
// calls (:83-634): the reverse NFA of `fwd` — transitions reversed, start and match swapped, look
// assertions treated as epsilon edges, the (?s:.)*? prefix left out when `anchored`.  State
// allocation order, placeholder kinds and byte-class registration follow the reference step by
// step (they decide the lazy DFA's alphabet).
void ReverseNFAStates(const NFA& fwd, bool anchored, NFA& out);

}  // namespace oracle
