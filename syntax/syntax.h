// syntax.h — restatement of Go's standard-library regexp/syntax parser (Perl flags).
//
// Why this exists: the reference (coregx/coregex) calls `syntax.Parse(pattern, syntax.Perl)`
// (reference meta/compile.go:58, nfa/compile.go:87, regex.go:483) and consumes the RAW AST
// (no Simplify()).  The parser lives in the Go toolchain (go.mod:3 pins go 1.25.4) and is NOT
// under /root/reference, and there is no Go toolchain in this image.  This directory restates
// the published algorithm of regexp/syntax/parse.go: literal coalescing (maybeConcat), the
// vertical-bar char-class merge (swapVerticalBar), alternation factoring rounds 1-4 (factor),
// Perl flags/groups, repeat validation, escapes, Perl + POSIX classes, ASCII case folding.
//
// It is a neutral dependency: both the product host compiler (coregex_b200/csrc/host) and the
// parity oracle (oracle/) consume the same AST, exactly as both coregex and its tests consume
// the same Go stdlib.  Parity is pinned at the RESULT level (match offsets vs the reference's
// own known-answer vectors and vs Python `re`), as SURVEY.md §8(c) item 1 describes.
//
// Limits (documented, tested): \p{..}/\P{..} Unicode groups are rejected with
// "unsupported: Unicode class"; case folding covers ASCII plus the two non-ASCII orbits that
// touch ASCII (K/k/U+212A, S/s/U+017F).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace gosyntax {

// Numeric order matters: parse.go compares Op values (swapVerticalBar, factor round 3).
enum Op : uint8_t {
  OpNoMatch = 1,
  OpEmptyMatch,
  OpLiteral,
  OpCharClass,
  OpAnyCharNotNL,
  OpAnyChar,
  OpBeginLine,
  OpEndLine,
  OpBeginText,
  OpEndText,
  OpWordBoundary,
  OpNoWordBoundary,
  OpCapture,
  OpStar,
  OpPlus,
  OpQuest,
  OpRepeat,
  OpConcat,
  OpAlternate,
  opPseudo = 128,
  opLeftParen,
  opVerticalBar,
};

enum Flags : uint16_t {
  FoldCase = 1,
  Literal = 2,
  ClassNL = 4,
  DotNL = 8,
  OneLine = 16,
  NonGreedy = 32,
  PerlX = 64,
  UnicodeGroups = 128,
  WasDollar = 256,
  Simple = 512,
  Perl = ClassNL | OneLine | PerlX | UnicodeGroups,
};

constexpr int32_t kMaxRune = 0x10FFFF;

struct Regexp {
  Op op = OpNoMatch;
  uint16_t flags = 0;
  std::vector<Regexp*> sub;   // subexpressions
  std::vector<int32_t> rune;  // literal runes, or class range pairs lo0,hi0,lo1,hi1,...
  int min = 0, max = 0;       // OpRepeat
  int cap = 0;                // OpCapture index
  std::string name;           // OpCapture name

  bool equal(const Regexp* y) const;
};

// Owns every node it hands out.
struct Arena {
  std::vector<std::unique_ptr<Regexp>> nodes;
  Regexp* make(Op op) {
    nodes.emplace_back(new Regexp());
    nodes.back()->op = op;
    return nodes.back().get();
  }
};

struct ParseResult {
  Regexp* re = nullptr;  // null on error
  std::string err;       // Go-formatted: "error parsing regexp: <code>: `<expr>`"
  int num_cap = 0;
};

ParseResult Parse(const std::string& pattern, uint16_t flags, Arena& arena);

// Deterministic s-expression dump used by tests to pin the AST shape, e.g.
//   cat{plus{cc{0x30-0x39}}lit{.}...}
std::string Dump(const Regexp* re);

// max capture index in the tree (Regexp.MaxCap)
int MaxCap(const Regexp* re);

}  // namespace gosyntax
