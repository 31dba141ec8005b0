#include "pike_pack.h"

namespace cgx {

static std::string Pack(const Prog& p, PikePacked& out, size_t max_inst, int max_groups, int max_threads);

std::string PackPike(const Prog& p, PikePacked& out) { return Pack(p, out, 64, 8, 32); }
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
