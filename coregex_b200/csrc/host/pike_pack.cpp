#include "pike_pack.h"

#include <algorithm>
#include <set>

namespace cgx {

static std::string Pack(const Prog& p, PikePacked& out, size_t max_inst, int max_groups, int max_threads);

// The most threads one generation of the anchored simulation can hold: the widest state of the
// subset construction over the program (a state = the set of byte-consuming/match instructions
// alive after some input; look-around assertions are taken as passing, which only widens it).
// The static count of consuming instructions bounds it too, but badly for UTF-8 automata, where a
// lead byte leaves only its own continuation states alive.  -1: construction abandoned (too many
// states) — the caller falls back to the static count.
static int MaxLiveThreads(const Prog& p) {
  const int n = (int)p.inst.size();
  auto closure = [&](const std::vector<int>& seeds) {
    std::vector<char> seen(n, 0);
    std::vector<int> stack(seeds.rbegin(), seeds.rend()), live;
    while (!stack.empty()) {
      const int pc = stack.back();
      stack.pop_back();
      if (pc < 0 || pc >= n || seen[pc]) continue;
      seen[pc] = 1;
      const Inst& in = p.inst[pc];
      switch (in.op) {
        case I_SET:
        case I_MATCH: live.push_back(pc); break;
        case I_SPLIT:
          stack.push_back(in.out1);
          stack.push_back(in.out);
          break;
        case I_SAVE:
        case I_ASSERT:
        case I_NOP: stack.push_back(in.out); break;
        default: break;
      }
    }
    std::sort(live.begin(), live.end());
    return live;
  };
  std::set<std::vector<int>> seen;
  std::vector<std::vector<int>> work{closure({p.start})};
  seen.insert(work[0]);
  size_t widest = work[0].size();
  while (!work.empty()) {
    const std::vector<int> st = work.back();
    work.pop_back();
    for (int b = 0; b < 256; b++) {
      std::vector<int> seeds;
      for (int pc : st)
        if (p.inst[pc].op == I_SET && set_has(p.sets[p.inst[pc].set], (unsigned)b)) seeds.push_back(p.inst[pc].out);
      if (seeds.empty()) continue;
      std::vector<int> nx = closure(seeds);
      if (seen.insert(nx).second) {
        widest = std::max(widest, nx.size());
        if (seen.size() > 3000) return -1;
        work.push_back(std::move(nx));
      }
    }
  }
  return (int)widest;
}

// captures kernel (pikevm_kernel.cu), small form: 64 instructions, 32 live threads; large form: 512
// instructions, 64 live threads.  The closure stack of the kernel grows by at most one frame per
// visited non-consuming instruction: 96 frames hold every 64-instruction program, 520 every
// 512-instruction one.
std::string PackPike(const Prog& p, PikePacked& out) {
  std::string err = Pack(p, out, 64, 8, 32);
  if (err.empty()) return err;
  const int live = MaxLiveThreads(p);
  err = Pack(p, out, 512, 8, live >= 0 ? (int)p.inst.size() : 64);
  if (!err.empty()) return err;
  if (live >= 0) {
    if (live > 64) return "more than 64 simultaneously live threads";
    out.nthreads = live;
  }
  out.large = p.inst.size() > 64 || out.nthreads > 32;
  return "";
}
std::string PackPikeSearch(const Prog& p, PikePacked& out) { return Pack(p, out, 2048, 1 << 20, 1024); }

static std::string Pack(const Prog& p, PikePacked& out, size_t max_inst, int max_groups, int max_threads) {
  out = PikePacked();
  if (p.inst.size() > max_inst) return "program has more than " + std::to_string(max_inst) + " instructions";
  if (p.num_captures > max_groups) return "more than " + std::to_string(max_groups) + " capture groups";
  int consuming = 0;
  for (auto& in : p.inst)
    if (in.op == I_SET || in.op == I_MATCH) consuming++;
  if (consuming > max_threads) return "more than " + std::to_string(max_threads) + " byte-consuming instructions";
  out.ninst = (int)p.inst.size();
  out.start = p.start;
  out.nslots = 2 * p.num_captures;
  out.nthreads = consuming;
  out.code.resize(2 * p.inst.size());
  for (size_t pc = 0; pc < p.inst.size(); pc++) {
    const Inst& in = p.inst[pc];
    uint32_t arg = 0;
    if (in.op == I_SET) arg = in.set;
    if (in.op == I_SAVE) arg = (uint32_t)in.slot;
    if (in.op == I_ASSERT) arg = in.look;
    uint32_t o = in.out < 0 ? 0xFFFFu : (uint32_t)in.out;
    uint32_t o1 = in.out1 < 0 ? 0xFFFFu : (uint32_t)in.out1;
    out.code[2 * pc] = (uint32_t)in.op | (arg << 8);
    out.code[2 * pc + 1] = o | (o1 << 16);
  }
  out.sets.resize(8 * p.sets.size());
  for (size_t k = 0; k < p.sets.size(); k++)
    for (int w = 0; w < 4; w++) {
      out.sets[8 * k + 2 * w] = (uint32_t)(p.sets[k][w] & 0xFFFFFFFFu);
      out.sets[8 * k + 2 * w + 1] = (uint32_t)(p.sets[k][w] >> 32);
    }
  return "";
}

}  // namespace cgx
