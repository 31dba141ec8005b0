// engine.h — compiled pattern: AST, program, reference-strategy classification, and the tables
// the GPU kernels consume.  Pure host C++ (no CUDA types); device residency lives in capi.cu.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "../scan_params.h"
#include "analysis.h"
#include "dfa.h"
#include "prog.h"
#include "pike_pack.h"
#include "teddy_tables.h"

namespace cgx {

enum EngineKind : int {
  ENG_DFA = 0,     // candidate filter + anchored DFA walk (scan_dfa.cu)
  ENG_TEDDY,       // nibble-fingerprint filter + ordered literal verify (scan_teddy.cu)
  ENG_PIKEVM,      // PikeVM search kernel: automata too large for the table kernels (pike_search.cu)
  ENG_LINE,        // one lane per record: unanchored forward DFA + reverse DFA (scan_dfa.cu)
};

// Capture groups of a flat pattern as item boundaries (flat_caps.cu): group g opens before item
// at[2g] and closes before item at[2g+1] of FlatDev::fwd_ops.  ok == false: some group sits under a
// quantifier (its offsets are those of the last iteration: the Pike captures kernel's business).
struct FlatCaps {
  bool ok = true;
  bool quantified = false;
  int nslots = 2;
  uint8_t at[34] = {};
};

struct Compiled {
  std::string pattern;
  gosyntax::Arena arena;
  const gosyntax::Regexp* re = nullptr;
  Prog prog;
  Analysis an;
  EngineKind kind = ENG_DFA;
  std::string engine_name;

  // ENG_DFA
  DfaTables dfa;
  int filter_kind = F_LUT;
  int nranges = 0;
  uint8_t rlo[4] = {0, 0, 0, 0}, rhi[4] = {0, 0, 0, 0};
  uint8_t lut[256];
  bool skip_safe = false;
  bool kind_lut_needed = false;
  FlatDev flat;  // nops == 0 when the pattern is not flat
  FlatCaps flat_caps;
  uint8_t delim = '\n';
  bool has_delim = true;  // false: matches may contain every byte value (PikeVM engine, one record)

  // ENG_LINE: unanchored forward DFA (leftmost-first end) and reverse DFA (longest = leftmost start)
  DfaTables udfa, rdfa;

  // ENG_TEDDY
  TeddyTables teddy;

  // ENG_PIKEVM: the whole program, packed for pike_search.cu
  PikePacked pike_search;

  // captures (FindAllSubmatchIndex)
  bool has_pike = false;
  PikePacked pike;       // packed program for pikevm_kernel.cu
  std::string pike_err;  // why captures are unavailable (when !has_pike)
};

// Source of "cgx_jit_prog.h" for the bitstream kernel specialised to this flat program: the class
// tests with their constants as immediates and the two marker passes as straight-line macro lists
// (scan_bits.cu, CGX_JIT).  Also the cache key of the compiled kernel.
std::string JitHeader(const FlatDev& f);

enum CompileStatus { COMPILE_OK = 0, COMPILE_SYNTAX = -1, COMPILE_UNSUPPORTED = -2 };

// err receives the Go-formatted syntax error or an "unsupported: ..." explanation
int CompilePattern(const std::string& pattern, std::unique_ptr<Compiled>& out, std::string& err,
                   const AnalysisConfig& cfg = AnalysisConfig());

}  // namespace cgx
