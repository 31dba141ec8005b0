// prog.h — host-side regex program for the B200 engines.
//
// The reference turns the AST into a Thompson NFA (reference nfa/compile.go:99-233) whose DFS
// order encodes leftmost-first priority.  The GPU build does not keep that state layout: it
// lowers the same AST into a flat Pike-style instruction array (byte-set instructions carry a
// 256-bit membership set instead of ByteRange/Sparse/ε-join triples), which is what both the
// eager DFA builder (dfa.cpp) and the PikeVM kernel (pikevm_kernel.cu) consume.
// Priority rule kept from the reference: Split prefers `out` over `out1`;
//   greedy loops put the loop body on `out`, lazy ones the exit (nfa/compile.go:1313-1483);
//   x* with nullable x is compiled as (x+)? (nfa/compile.go:1353-1387).
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "../../../syntax/syntax.h"

namespace cgx {

enum InstOp : uint8_t { I_FAIL = 0, I_SET, I_SPLIT, I_SAVE, I_ASSERT, I_NOP, I_MATCH };

// look kinds — same six the reference has (nfa/nfa.go look kinds; nfa/pikevm.go:1646-1675)
enum LookKind : uint8_t { L_START_TEXT = 0, L_END_TEXT, L_START_LINE, L_END_LINE, L_WORD, L_NOT_WORD };

struct Inst {
  InstOp op = I_FAIL;
  uint8_t look = 0;     // I_ASSERT
  uint16_t set = 0;     // I_SET: index into Prog::sets
  int32_t out = -1;     // primary successor
  int32_t out1 = -1;    // I_SPLIT: lower-priority successor
  int32_t slot = 0;     // I_SAVE: capture slot (2*group + isEnd)
};

using ByteSet = std::array<uint64_t, 4>;
inline bool set_has(const ByteSet& s, unsigned b) { return (s[b >> 6] >> (b & 63)) & 1; }
inline void set_add(ByteSet& s, unsigned lo, unsigned hi) {
  for (unsigned b = lo; b <= hi; b++) s[b >> 6] |= 1ull << (b & 63);
}

struct Prog {
  std::vector<Inst> inst;
  std::vector<ByteSet> sets;
  int start = 0;          // anchored entry
  int num_captures = 1;   // groups incl. group 0
  std::vector<std::string> cap_names;  // num_captures entries; "" for group 0 and unnamed groups
  bool anchored_start = false;  // pattern begins with \A (reference nfa/compile.go:1755-1775)
  bool has_looks = false;
  bool has_word_looks = false;
};

// Returns "" on success.  ASCII classes only (same scope as SURVEY.md §2.1 row 10).
std::string CompileProg(const gosyntax::Regexp* re, Prog& out);

// reverse program: matches the reversed language, used for reverse DFAs (N1)
std::string CompileReverseProg(const gosyntax::Regexp* re, Prog& out);

}  // namespace cgx
