// pike_pack.h — packs a Prog into the flat form pikevm_kernel.cu interprets.
//
// code[2*pc]   = op | arg << 8     (I_SET: set index; I_SAVE: slot; I_ASSERT: look kind)
// code[2*pc+1] = out | out1 << 16  (0xFFFF = none)
// sets[8*k .. 8*k+7] = 256-bit membership of byte set k
// Limits of the captures kernel (per-lane state lives in registers/local memory):
//   small form <= 64 instructions and <= 32 live threads, large form <= 512 instructions and <= 64
//   simultaneously live threads (proved by a subset construction, pike_pack.cpp); <= 8 groups.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "prog.h"

namespace cgx {

struct PikePacked {
  std::vector<uint32_t> code;
  std::vector<uint32_t> sets;
  int ninst = 0, start = 0, nslots = 0, nthreads = 0;
  bool large = false;  // captures kernel: the 512-instruction form
};

// returns "" or the reason the program does not fit the kernel
std::string PackPike(const Prog& p, PikePacked& out);
// the same layout for the search kernel (pike_search.cu): up to 2048 instructions, 1024 live threads
std::string PackPikeSearch(const Prog& p, PikePacked& out);

}  // namespace cgx
