// engine.cpp — pattern -> GPU engine selection and table building.
#include "engine.h"

#include <cstdio>
#include <cstring>

namespace cgx {

using namespace gosyntax;

bool BuildTeddyTables(const std::vector<std::string>& patterns, TeddyTables& t, size_t max_patterns) {
  t = TeddyTables();
  const size_t n = patterns.size();
  if (n < 2 || n > max_patterns) return false;
  size_t mn = patterns[0].size(), mx = 0;
  for (auto& p : patterns) {
    if (p.size() < 3) return false;
    mn = std::min(mn, p.size());
    mx = std::max(mx, p.size());
  }
  t.npat = (int)n;
  t.min_len = (int)mn;
  t.max_len = (int)mx;
  t.fp_len = 2;  // DefaultTeddyConfig / DefaultFatTeddyConfig FingerprintLen=2, min pattern len 3
  // slim: min(8, n) buckets (teddy.go:277-281); fat (33..64 patterns): always 16 (teddy_fat.go:205)
  t.nbuckets = n <= 32 ? (int)std::min<size_t>(8, n) : 16;
  t.offs.push_back(0);
  for (auto& p : patterns) {
    t.bytes.insert(t.bytes.end(), p.begin(), p.end());
    t.offs.push_back((int32_t)t.bytes.size());
  }
  uint16_t exact[2][256] = {};
  t.bucket_of.resize(n);
  std::vector<std::vector<uint16_t>> buckets(t.nbuckets);
  for (size_t id = 0; id < n; id++) {
    int b = (int)(id % t.nbuckets);
    t.bucket_of[id] = (uint8_t)b;
    buckets[b].push_back((uint16_t)id);
    for (int pos = 0; pos < 2; pos++) exact[pos][(uint8_t)patterns[id][pos]] |= (uint16_t)(1u << b);
  }
  t.fp0.assign(exact[0], exact[0] + 256);
  t.fp1.assign(exact[1], exact[1] + 256);
  t.bucket_off.push_back(0);
  for (auto& b : buckets) {
    for (uint16_t id : b) t.order_simd.push_back(id);
    t.bucket_off.push_back((uint16_t)t.order_simd.size());
  }
  t.fp_packed.resize(256);
  for (int c = 0; c < 256; c++) t.fp_packed[c] = (uint32_t)t.fp0[c] | ((uint32_t)t.fp1[c] << 16);
  return true;
}

static bool setToRanges(const ByteSet& s, int& n, uint8_t* lo, uint8_t* hi) {
  n = 0;
  int b = 0;
  while (b < 256) {
    if (!set_has(s, b)) {
      b++;
      continue;
    }
    int e = b;
    while (e + 1 < 256 && set_has(s, e + 1)) e++;
    if (n == 4 || e > 0x7F) return false;  // SWAR path: <= 4 ASCII ranges
    lo[n] = (uint8_t)b;
    hi[n] = (uint8_t)e;
    n++;
    b = e + 1;
  }
  return n > 0;
}

// ---- flat-pattern extraction for the bit-parallel start filter ---------------------------------
// Accepts Concat/Capture/Literal/CharClass/Plus/Star/Quest/Repeat{m,n<=4} over ASCII classes with
// at most 4 ranges each, at most 4 distinct classes and 24 items.  Greedy vs lazy is irrelevant
// (the filter answers "does some match start here").
namespace {
struct FlatBuilder {
  FlatDev f;
  uint8_t lazy[24] = {};  // item carries the non-greedy flag (matters only to the bitstream engine)
  bool ok = true;
  // capture groups as item boundaries: slot 2g = items before group g opens, slot 2g+1 = items
  // before it closes.  A group under a quantifier (`(\d)+`, `(a)?`) has no fixed boundary.
  FlatCaps caps;
  int classOf(const ByteSet& s) {
    uint8_t lo[4], hi[4];
    int n;
    if (!setToRanges(s, n, lo, hi)) { ok = false; return 0; }
    for (int c = 0; c < f.nclasses; c++) {
      if (f.cls_nranges[c] != n) continue;
      bool same = true;
      for (int k = 0; k < n; k++) same &= f.cls_lo[c][k] == lo[k] && f.cls_hi[c][k] == hi[k];
      if (same) return c;
    }
    if (f.nclasses == 4) { ok = false; return 0; }
    int c = f.nclasses++;
    f.cls_nranges[c] = (uint8_t)n;
    for (int k = 0; k < n; k++) { f.cls_lo[c][k] = lo[k]; f.cls_hi[c][k] = hi[k]; }
    return c;
  }
  void push(int kind, const ByteSet& s, bool nongreedy = false) {
    if (!ok) return;
    if (f.nops == 24) { ok = false; return; }
    int c = classOf(s);
    if (!ok) return;
    lazy[f.nops] = nongreedy ? 1 : 0;
    f.op_kind[f.nops] = (uint8_t)kind;
    f.op_class[f.nops] = (uint8_t)c;
    f.nops++;
  }
  static bool classSet(const Regexp* re, ByteSet& s, bool* quantified_cap = nullptr) {
    s = ByteSet{};
    if (re->op == OpCharClass) {
      if (re->rune.empty()) return false;
      for (size_t k = 0; k + 1 < re->rune.size(); k += 2) {
        if (re->rune[k + 1] > 127) return false;
        set_add(s, (unsigned)re->rune[k], (unsigned)re->rune[k + 1]);
      }
      return true;
    }
    if (re->op == OpLiteral && re->rune.size() == 1 && re->rune[0] < 128) {
      unsigned r = (unsigned)re->rune[0];
      set_add(s, r, r);
      bool letter = (r >= 'a' && r <= 'z') || (r >= 'A' && r <= 'Z');
      if ((re->flags & FoldCase) && letter) set_add(s, r ^ 0x20, r ^ 0x20);
      return true;
    }
    if (re->op == OpCapture && re->sub.size() == 1) {
      if (quantified_cap) *quantified_cap = true;
      return classSet(re->sub[0], s, quantified_cap);
    }
    return false;
  }
  void walk(const Regexp* re) {
    if (!ok) return;
    ByteSet s;
    switch (re->op) {
      case OpConcat:
        for (auto* x : re->sub) walk(x);
        return;
      case OpCapture:
        if (re->sub.size() == 1) {
          const int g = re->cap;
          if (g <= 0 || 2 * g + 1 >= (int)sizeof caps.at) caps.ok = false;
          else caps.at[2 * g] = (uint8_t)f.nops;
          walk(re->sub[0]);
          if (caps.ok && g > 0) {
            caps.at[2 * g + 1] = (uint8_t)f.nops;
            if (2 * g + 2 > caps.nslots) caps.nslots = 2 * g + 2;
          }
          return;
        }
        ok = false;
        return;
      case OpLiteral:
        for (int32_t r : re->rune) {
          if (r > 127) { ok = false; return; }
          ByteSet b{};
          set_add(b, (unsigned)r, (unsigned)r);
          bool letter = (r >= 'a' && r <= 'z') || (r >= 'A' && r <= 'Z');
          if ((re->flags & FoldCase) && letter) set_add(b, (unsigned)r ^ 0x20, (unsigned)r ^ 0x20);
          push(0, b);
        }
        return;
      case OpCharClass:
        if (!classSet(re, s)) { ok = false; return; }
        push(0, s);
        return;
      case OpPlus: case OpStar: case OpQuest:
        if (re->sub.size() != 1 || !classSet(re->sub[0], s, &caps.quantified)) { ok = false; return; }
        push(re->op == OpPlus ? 1 : re->op == OpStar ? 2 : 3, s, (re->flags & NonGreedy) != 0);
        return;
      case OpRepeat: {
        if (re->sub.size() != 1 || !classSet(re->sub[0], s, &caps.quantified)) { ok = false; return; }
        int mx = re->max == -1 ? re->min : re->max;
        if (mx > 8) { ok = false; return; }
        const bool ng = (re->flags & NonGreedy) != 0;
        for (int i = 0; i < re->min; i++) push(0, s);
        if (re->max == -1) push(2, s, ng);
        else for (int i = re->min; i < re->max; i++) push(3, s, ng);
        return;
      }
      default:
        ok = false;
    }
  }
};
}  // namespace

static void BuildFlat(const Regexp* re, FlatDev& out, uint8_t (&lazy)[24], FlatCaps& caps) {
  FlatBuilder b;
  memset(&b.f, 0, sizeof b.f);
  b.walk(re);
  memset(&out, 0, sizeof out);
  memcpy(lazy, b.lazy, sizeof lazy);
  caps = b.caps;
  if (!b.ok || b.caps.quantified) caps.ok = false;
  // worth it only when the pattern is longer than its first item (otherwise the first-level
  // filter already is the whole pattern) — except a lone `C+` (the reference's CharClassSearcher
  // patterns, nfa/charclass_searcher.go), where "flat" unlocks the run-start filter: one
  // candidate per run instead of one per byte
  if (!(b.ok && (b.f.nops >= 2 || (b.f.nops == 1 && b.f.op_kind[0] == 1)))) return;
  for (int c = 0; c < b.f.nclasses; c++)
    for (int r = 0; r < b.f.cls_nranges[c]; r++) {
      uint32_t lo = b.f.cls_lo[c][r], hi = b.f.cls_hi[c][r], w = hi - lo, m = 0;
      while (m < w) m = m * 2 + 1;
      if ((lo & m) == 0) {  // x in [lo,hi]  <=>  (x ^ lo) <= w
        b.f.cls_mode[c][r] = 0;
        b.f.cls_k1[c][r] = lo * 0x01010101u;
        b.f.cls_k2[c][r] = (0x7Fu - w) * 0x01010101u;
      } else {
        b.f.cls_mode[c][r] = 1;
        b.f.cls_k1[c][r] = 0x80808080u - lo * 0x01010101u;
        b.f.cls_k2[c][r] = (hi * 0x01010101u) | 0x80808080u;
      }
    }
  // right-to-left program: skip trailing items that can be empty (they leave M = everything),
  // the first ONE/PLUS item initialises M with its class
  int k = b.f.nops - 1;
  while (k >= 0 && b.f.op_kind[k] >= 2) k--;
  if (k < 0) return;  // nullable pattern: not filtered
  b.f.rev_init_class = b.f.op_class[k];
  b.f.rev_nops = 0;
  for (k--; k >= 0; k--) b.f.rev_ops[b.f.rev_nops++] = (uint8_t)(b.f.op_kind[k] | (b.f.op_class[k] << 2));
  out = b.f;
}

// Is the flat pattern deterministic enough for the bitstream engine (scan_flat.cu)?  That engine
// finds match ENDS with a forward marker pass that always takes the greedy choice; this equals
// leftmost-first (reference nfa/pikevm.go thread priority, dfa/lazy break-at-match) when no choice
// ever exists: the class of every quantified item is disjoint from everything that can follow it
// up to and including the next mandatory item (later optional items of the SAME class are one
// counted group and fine).  A lazy quantifier only matters when nothing mandatory follows it.
static void DecideBitstream(Compiled& c, const uint8_t (&lazy)[24]) {
  FlatDev& f = c.flat;
  f.bs_ok = f.bs_runstart = f.bs_midrun_check = f.fwd_nops = 0;
  if (!f.nops || c.kind_lut_needed || f.rev_nops + 1 > f.nops) return;
  ByteSet cs[4] = {};
  ByteSet any{};
  for (int k = 0; k < f.nclasses; k++)
    for (int r = 0; r < f.cls_nranges[k]; r++) {
      set_add(cs[k], f.cls_lo[k][r], f.cls_hi[k][r]);
      set_add(any, f.cls_lo[k][r], f.cls_hi[k][r]);
    }
  auto inter = [&](int a, int b) {
    for (int x = 0; x < 256; x++)
      if (set_has(cs[a], x) && set_has(cs[b], x)) return true;
    return false;
  };
  auto nullable = [&](int i) { return f.op_kind[i] >= 2; };
  if (nullable(0)) return;  // every byte of the first class would start its own overlapping match
  // `a a+` / `a a?`: matches from neighbouring starts share their end, which the engine detects and
  // sends to the serial path every time — such patterns stay on the candidate/DFA kernel
  if (f.op_kind[0] == 0 && f.nops > 1 && f.op_kind[1] != 0 && inter(f.op_class[0], f.op_class[1])) return;
  for (int i = 0; i < f.nops; i++) {
    if (f.op_kind[i] == 0) continue;
    bool tail_nullable = true;
    for (int j = i + 1; j < f.nops; j++) {
      const bool same_group = f.op_class[j] == f.op_class[i] && nullable(j);
      if (!same_group && inter(f.op_class[i], f.op_class[j])) return;
      if (!nullable(j)) {
        tail_nullable = false;
        break;
      }
    }
    if (lazy[i] && tail_nullable) return;
  }
  if (c.filter_kind == F_RUNSTART) {
    if (!(f.first_is_filter && f.op_kind[0] == 1)) return;
    f.bs_runstart = 1;
    f.bs_midrun_check = !(f.op_kind[f.nops - 1] == 1 && f.op_class[f.nops - 1] == f.op_class[0]);
  }
  for (int i = 0; i < f.nops; i++) f.fwd_ops[i] = (uint8_t)(f.op_kind[i] | (f.op_class[i] << 2));
  f.fwd_nops = f.nops;
  for (int x = 0; x < 256; x++)
    if (!set_has(any, x)) f.sync_lut[x >> 5] |= 1u << (x & 31);
  f.bs_ok = 1;
}

std::string JitHeader(const FlatDev& f) {
  char buf[256];
  std::string o = "// generated by host/engine.cpp JitHeader\n";
  auto add = [&](const char* fmt, auto... args) {
    snprintf(buf, sizeof buf, fmt, args...);
    o += buf;
  };
  add("#define CGX_JIT_NCLASSES %d\n", f.nclasses);
  add("#define CGX_JIT_RUNSTART %d\n", f.bs_runstart);
  add("#define CGX_JIT_MIDRUN %d\n", f.bs_midrun_check);
  add("#define CGX_JIT_REV_INIT %d\n", f.rev_init_class);
  // a match can be one or two bytes long (`\d+`, `\w+`, `[a-z]+=`): chunks hold far more matches than
  // the staging buffer, the kernel keeps a second set of bitmaps instead (scan_bits.cu CGX_PARK)
  {
    int mandatory = 0;
    for (int i = 0; i < f.fwd_nops; i++) mandatory += (f.fwd_ops[i] & 3) <= 1 ? 1 : 0;
    add("#define CGX_JIT_PARK %d\n", mandatory <= 2 && f.nclasses <= 2 ? 1 : 0);
  }
  // classes that cover much of ordinary text (`[a-z]+/\d+`, `\w+@...`) leave long stretches without
  // a sync byte: open segments get the bit-parallel replay (scan_bits.cu replay_bits).  Narrow
  // classes (digits and dots) hardly ever do, and the mere presence of that code costs the hot loop
  // 2 % on B200, so they keep the plain replay.
  {
    int nsync = 0;
    for (int b = 0; b < 256; b++) nsync += (f.sync_lut[b >> 5] >> (b & 31)) & 1u;
    add("#define CGX_RBITS %d\n", 256 - nsync >= 20 ? 1 : 0);
  }
  // Every step of a pass owns one slot I of per-lane state (the carry and the shifted-out bits that
  // travel from one word to the next, scan_bits.cu PassState).  Right to left, a single byte of
  // class A followed by a run of class B (`B+ A` in the pattern) is one fused step over the
  // precomputed bitmap "A right after B": one 2-position shift, no intermediate marker set
  // (scan_bits.cu rev_fused); the bitmap itself takes a state slot too (DEF).
  std::string fused_defs, fused_keys;
  int slot = 0, rev_slots = 0;
  std::string rev = "#define CGX_JIT_REV_PASS(STEP, FUSE)";
  for (int k = 0; k < f.rev_nops; k++) {
    const int kind = f.rev_ops[k] & 3, cls = f.rev_ops[k] >> 2;
    if (kind == 0 && k + 1 < f.rev_nops && (f.rev_ops[k + 1] & 3) == 1) {
      const int cls2 = f.rev_ops[k + 1] >> 2;
      snprintf(buf, sizeof buf, " FUSE(%d, %d, %d)", cls, cls2, slot++);
      rev += buf;
      snprintf(buf, sizeof buf, "(%d,%d)", cls, cls2);
      if (fused_keys.find(buf) == std::string::npos) {
        fused_keys += buf;
        snprintf(buf, sizeof buf, " DEF(%d, %d, @%zu@)", cls, cls2, fused_keys.size());
        fused_defs += buf;
      }
      k++;
    } else {
      snprintf(buf, sizeof buf, " STEP(%d, %d, %d)", kind, cls, slot++);
      rev += buf;
    }
  }
  // the DEFs take the slots after the steps'
  {
    std::string d;
    size_t pos = 0;
    while (pos < fused_defs.size()) {
      const size_t at = fused_defs.find('@', pos);
      if (at == std::string::npos) {
        d += fused_defs.substr(pos);
        break;
      }
      const size_t at2 = fused_defs.find('@', at + 1);
      d += fused_defs.substr(pos, at - pos) + std::to_string(slot++);
      pos = at2 + 1;
    }
    fused_defs = d;
  }
  rev_slots = slot;
  o += rev;
  o += "\n#define CGX_JIT_REV_FUSED(DEF)" + fused_defs;
  o += "\n#define CGX_JIT_FWD_PASS(STEP)";
  for (int k = 0; k < f.fwd_nops; k++) add(" STEP(%d, %d, %d)", f.fwd_ops[k] & 3, f.fwd_ops[k] >> 2, k);
  const int nstate = rev_slots > f.fwd_nops ? rev_slots : f.fwd_nops;
  add("\n#define CGX_JIT_NSTATE %d\n", nstate > 0 ? nstate : 1);
  o += "template <int C> __device__ __forceinline__ uint32_t cgx_jit_flags(uint32_t w, uint32_t one) { return 0u; }\n";
  for (int c = 0; c < f.nclasses; c++) {
    add("template <> __device__ __forceinline__ uint32_t cgx_jit_flags<%d>(uint32_t w, uint32_t one) {\n  uint32_t fl = 0u;\n", c);
    for (int r = 0; r < f.cls_nranges[c]; r++) {
      if (f.cls_mode[c][r] == 0)
        // 2 ALU-pipe + 1 FMA-pipe instruction per word (scan_bits.cu "pipe-aware primitives")
        add("  { const uint32_t z = mad_fma(xor_and(w, 0x%08Xu, 0x7F7F7F7Fu), one, 0x%08Xu); fl %s nor_and(z, w, 0x80808080u); }\n",
            f.cls_k1[c][r], f.cls_k2[c][r], r == 0 ? "=" : "|=");
      else
        add("  fl |= swar_in_range(w, 0x%08Xu, 0x%08Xu);\n", f.cls_k1[c][r], f.cls_k2[c][r]);
    }
    o += "  return fl;\n}\n";
  }
  return o;
}

int CompilePattern(const std::string& pattern, std::unique_ptr<Compiled>& out, std::string& err,
                   const AnalysisConfig& cfg) {
  std::unique_ptr<Compiled> c(new Compiled());
  c->pattern = pattern;
  memset(&c->flat, 0, sizeof c->flat);
  ParseResult pr = Parse(pattern, Perl, c->arena);
  if (!pr.re) {
    err = pr.err;
    return pr.err.find("unsupported:") != std::string::npos ? COMPILE_UNSUPPORTED : COMPILE_SYNTAX;
  }
  c->re = pr.re;
  std::string e = CompileProg(pr.re, c->prog);
  if (!e.empty()) {
    err = e;
    return COMPILE_UNSUPPORTED;
  }
  c->an = Analyze(pr.re, (int)c->prog.inst.size(), c->prog.anchored_start, cfg);

  // (?m)^-anchored literal sets use the reference's line-anchor wrapper (meta/compile.go:663-686),
  // whose results equal plain leftmost-first: they run on the DFA engine here.
  if (c->an.strategy == RS_UseTeddy && !c->an.has_anchors) {
    std::vector<std::string> pats;
    for (auto& l : c->an.prefixes) pats.push_back(l.bytes);
    bool has_nl = false;
    for (auto& p : pats)
      if (p.find('\n') != std::string::npos) has_nl = true;
    if (!has_nl && BuildTeddyTables(pats, c->teddy)) {
      c->kind = ENG_TEDDY;
      c->engine_name = c->teddy.nbuckets == 16 ? "fat-teddy" : "teddy";
    }
  }
  // More than 64 complete literals: the reference's Aho-Corasick strategy (meta/strategy.go:1165; the
  // automaton is the external module github.com/coregx/ahocorasick, not in the reference tree).  What
  // it returns is the leftmost match, and among literals that start there the first of the
  // alternation.  When no literal is a prefix of another, at most one literal matches at a position,
  // the order of verification cannot be observed, and the set runs on the multi-literal engine with
  // 16 buckets (fingerprint filter + byte-for-byte verification) instead of a DFA that does not fit
  // the table kernels.  Sets with prefix-related literals keep the generic engines.
  if (c->an.strategy == RS_UseAhoCorasick && !c->an.has_anchors) {
    std::vector<std::string> pats;
    for (auto& l : c->an.prefixes) pats.push_back(l.bytes);
    bool ok = true;
    for (size_t i = 0; i < pats.size() && ok; i++) {
      if (pats[i].find('\n') != std::string::npos) ok = false;
      for (size_t j = 0; j < pats.size() && ok; j++)
        if (i != j && pats[j].size() >= pats[i].size() && pats[j].compare(0, pats[i].size(), pats[i]) == 0) ok = false;
    }
    if (ok && BuildTeddyTables(pats, c->teddy, 1024)) {
      c->kind = ENG_TEDDY;
      c->engine_name = "teddy-large";
    }
  }

  // The anchored DFA is built for every pattern: ENG_DFA needs it, and it is also how the
  // record-delimiter safety of the pattern is proven.
  std::string de = BuildDFA(c->prog, /*anchored=*/true, /*max_states=*/160, c->dfa);
  const bool nullable_pat = c->an.can_match_empty || (de.empty() && c->dfa.matches_empty);
  // The PikeVM search kernel takes what the table kernels cannot: automata over 160 DFA states,
  // patterns that can match the empty string (the empty-match rules of meta/findall.go:247-279 are
  // sequential within a record) and patterns whose matches may contain every byte value (no record
  // delimiter: the haystack is one record, scanned by one lane) — as the reference's PikeVM does
  // when it selects UseNFA or its lazy DFA gives up (meta/find_indices.go:1172, nfa/pikevm.go:1711).
  auto to_pikevm = [&](const std::string& why) -> bool {
    const std::string pe = PackPikeSearch(c->prog, c->pike_search);
    if (!pe.empty()) {
      err = "unsupported: " + why + "; PikeVM search kernel: " + pe;
      return false;
    }
    // records are cut at a byte no instruction can consume
    ByteSet any{};
    for (auto& in : c->prog.inst)
      if (in.op == I_SET)
        for (int w = 0; w < 4; w++) any[w] |= c->prog.sets[in.set][w];
    static const char pref[] = "\n etaoinsrhldcumfpgwybvkxjqz\t,.;:/-_=0123456789ETAOINSRHLDCUMFPGWYBVKXJQZ";
    int chosen = -1;
    for (const char* q = pref; *q && chosen < 0; q++)
      if (!set_has(any, (uint8_t)*q)) chosen = (uint8_t)*q;
    for (int d = 0; d < 256 && chosen < 0; d++)
      if (!set_has(any, (unsigned)d)) chosen = d;
    c->has_delim = chosen >= 0;
    c->delim = (uint8_t)(chosen >= 0 ? chosen : 0);
    c->kind = ENG_PIKEVM;
    c->engine_name = c->has_delim ? "pikevm" : "pikevm-serial";
    return true;
  };
  // (the multi-literal engine needs no DFA — 64 literals do not fit 160 states — and a set of
  // literals of three bytes and more never matches the empty string)
  if (c->kind == ENG_DFA && (!de.empty() || nullable_pat)) {
    if (!to_pikevm(de.empty() ? std::string("pattern can match the empty string") : de)) return COMPILE_UNSUPPORTED;
  }
  if (c->kind == ENG_DFA) {
    // Records are cut at a byte no match can contain.  '\n' whenever possible (lines are what
    // callers think in); otherwise the most frequent text byte the pattern cannot consume, e.g.
    // `\s+` is scanned as records between letters, `[^a]+` as records between 'a's.
    {
      static const char pref[] = "\n etaoinsrhldcumfpgwybvkxjqz\t,.;:/-_=0123456789ETAOINSRHLDCUMFPGWYBVKXJQZ";
      int chosen = -1;
      for (const char* q = pref; *q && chosen < 0; q++)
        if (DelimiterSafe(c->dfa, (uint8_t)*q)) chosen = (uint8_t)*q;
      for (int d = 0; d < 256 && chosen < 0; d++)
        if (DelimiterSafe(c->dfa, (uint8_t)d)) chosen = d;
      if (chosen < 0) {
        if (!to_pikevm("a match can contain every byte value (no record delimiter)")) return COMPILE_UNSUPPORTED;
      } else {
        c->delim = (uint8_t)chosen;
      }
    }
  }
  if (c->kind == ENG_DFA) {
    for (int k = 1; k < SK_COUNT; k++)
      if (c->dfa.start[k] != c->dfa.start[0]) c->kind_lut_needed = true;
    memset(c->lut, 0, sizeof c->lut);
    uint8_t flat_lazy[24];
    BuildFlat(pr.re, c->flat, flat_lazy, c->flat_caps);
    auto flat_first = [&]() {
      FlatDev& f = c->flat;
      f.first_is_filter = f.nops && f.cls_nranges[0] == c->nranges;
      for (int k = 0; k < c->nranges && f.first_is_filter; k++)
        if (f.cls_lo[0][k] != c->rlo[k] || f.cls_hi[0][k] != c->rhi[k]) f.first_is_filter = 0;
    };
    if (c->an.strategy == RS_UseDigitPrefilter) {
      // reference meta/find_indices.go:1050-1088: candidates are ASCII digits; with
      // digitRunSkipSafe a failed candidate skips the rest of its digit run.
      c->nranges = 1;
      c->rlo[0] = '0';
      c->rhi[0] = '9';
      c->skip_safe = c->an.digit_run_skip_safe;
      c->filter_kind = c->skip_safe ? F_RUNSTART : F_BYTESET;
      c->engine_name = c->skip_safe ? "dfa-runstart" : "dfa-byteset";
      flat_first();
    } else {
      for (int b = 0; b < 256; b++) c->lut[b] = set_has(c->dfa.first_bytes, b) ? 1 : 0;
      if (setToRanges(c->dfa.first_bytes, c->nranges, c->rlo, c->rhi)) {
        c->filter_kind = F_BYTESET;
        c->engine_name = "dfa-byteset";
      } else {
        c->nranges = 0;
        c->filter_kind = F_LUT;
        c->engine_name = "dfa-lut";
      }
      flat_first();
      // A flat pattern that opens with C+ and whose possible first bytes are exactly C: a match
      // starting inside a run of C implies one starting at the run's first byte, so only run
      // starts (and resume positions after a match, handled by the replay) can begin a match.
      if (c->flat.nops && c->flat.first_is_filter && c->flat.op_kind[0] == 1 &&
          c->filter_kind == F_BYTESET) {
        c->filter_kind = F_RUNSTART;
        c->skip_safe = true;
        c->engine_name = "dfa-runstart";
      }
    }
    DecideBitstream(*c, flat_lazy);
    if (c->flat.nops) c->engine_name += c->flat.bs_ok ? "+bitstream" : "+flat";
    // A pattern that can begin with (almost) any byte has no useful candidate filter: every
    // position would start an anchored walk, which is quadratic per record.  Such patterns run
    // the way the reference runs UseDFA patterns (meta/find_indices.go:686-705): one unanchored
    // forward DFA pass for the match end, one reverse DFA pass for its start — per record, linear.
    int first = 0;
    for (int b = 0; b < 256; b++) first += set_has(c->dfa.first_bytes, b) ? 1 : 0;
    // (the record engine marks delimiters with the SWAR byte test of phase A, which holds for bytes
    // below 0x80 only: a pattern whose only safe delimiters are high bytes — `\P{Han}+`, cut at
    // 0xC0 — stays on the candidate + anchored-walk engine)
    if (!c->flat.nops && first > 64 && c->delim < 0x80) {
      Prog rprog;
      if (CompileReverseProg(pr.re, rprog).empty() &&
          BuildDFA(c->prog, /*anchored=*/false, 160, c->udfa).empty() &&
          BuildDFA(rprog, /*anchored=*/true, 160, c->rdfa, /*longest=*/true).empty() &&
          c->udfa.nstates + c->rdfa.nstates <= 200 && !c->rdfa.matches_empty) {
        c->kind = ENG_LINE;
        c->engine_name = "line-dfa";
      }
    }
  }
  c->pike_err = PackPike(c->prog, c->pike);
  c->has_pike = c->pike_err.empty();
  out = std::move(c);
  return COMPILE_OK;
}

}  // namespace cgx
