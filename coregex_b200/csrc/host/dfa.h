// dfa.h — eager, anchored, leftmost-first DFA for the GPU table-walk kernels.
//
// What it replaces: the reference determinizes lazily at search time into a premultiplied,
// tagged u32 table with a 1-byte match delay (reference dfa/lazy/lazy.go:1336-1446,
// dfa/lazy/state.go:26-55).  A GPU kernel cannot call back into a determinizer, so the whole
// automaton is built on the host before launch and shipped as one dense table:
//
//   trans[state*256 + byte] : u16, low 15 bits = next state (0 = DEAD),
//                             bit 15 = "a match ends right before this byte"
//   eoi[state]              : u8, 1 = a match ends at end of input in this state
//   start[kind]             : u16 per look-behind kind (nonword, word, text, LF, CR) — the same
//                             five kinds as reference dfa/lazy/start.go:17-37
//
// Leftmost-first is obtained the same way the reference gets it (break-at-match over an
// insertion-ordered NFA set, dfa/lazy/builder.go:210-213): DFA states are ORDERED thread lists and
// everything after the first Match thread is cut.  Look-ahead assertions ($, \b, \B) stay pending
// inside a state and are resolved against the incoming byte, which is why the match flag sits on
// the transition rather than on the state.  States that can no longer reach a match are folded
// into DEAD so walks stop as early as possible.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "prog.h"

namespace cgx {

enum StartKindIdx { SK_NONWORD = 0, SK_WORD, SK_TEXT, SK_LF, SK_CR, SK_COUNT };

constexpr uint16_t DFA_MATCH_BIT = 0x8000;
constexpr uint16_t DFA_STATE_MASK = 0x7FFF;

struct DfaTables {
  int nstates = 0;               // including DEAD (0)
  std::vector<uint16_t> trans;   // nstates * 256
  std::vector<uint8_t> eoi;      // nstates
  uint16_t start[SK_COUNT] = {0, 0, 0, 0, 0};
  bool matches_empty = false;    // some start state matches before consuming anything
  ByteSet first_bytes{};         // bytes with a live transition out of some start state
};

// anchored: true -> match must begin at the walk start (the only mode the candidate kernels use)
//           false -> prepend the (?s:.)*? prefix (unanchored forward DFA, N1)
// Returns "" or an error ("dfa too large: ...").
// longest:  false -> leftmost-first (everything after the first Match thread is cut)
//           true  -> no cut: the flag says "some thread matches here"; walking on and keeping the
//                    last flag yields the longest match (reverse DFAs: the leftmost start,
//                    reference dfa/lazy/lazy.go:1769-1920 with BreakAtMatch=false)
std::string BuildDFA(const Prog& p, bool anchored, int max_states, DfaTables& out, bool longest = false);

// true when no match can contain byte d (every live state goes DEAD on d)
bool DelimiterSafe(const DfaTables& t, uint8_t d);

}  // namespace cgx
