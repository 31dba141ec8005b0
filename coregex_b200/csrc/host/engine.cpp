// engine.cpp — pattern -> GPU engine selection and table building.
#include "engine.h"

#include <cstring>

namespace cgx {

using namespace gosyntax;

bool BuildTeddyTables(const std::vector<std::string>& patterns, TeddyTables& t) {
  t = TeddyTables();
  const size_t n = patterns.size();
  if (n < 2 || n > 64) return false;
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
  uint16_t lo[2][16] = {}, hi[2][16] = {};
  t.bucket_of.resize(n);
  std::vector<std::vector<uint16_t>> buckets(t.nbuckets);
  for (size_t id = 0; id < n; id++) {
    int b = (int)(id % t.nbuckets);
    t.bucket_of[id] = (uint8_t)b;
    buckets[b].push_back((uint16_t)id);
    for (int pos = 0; pos < 2; pos++) {
      uint8_t c = (uint8_t)patterns[id][pos];
      lo[pos][c & 15] |= (uint16_t)(1u << b);
      hi[pos][c >> 4] |= (uint16_t)(1u << b);
    }
  }
  t.fp0.resize(256);
  t.fp1.resize(256);
  for (int c = 0; c < 256; c++) {
    t.fp0[c] = lo[0][c & 15] & hi[0][c >> 4];
    t.fp1[c] = lo[1][c & 15] & hi[1][c >> 4];
  }
  for (auto& b : buckets)
    for (uint16_t id : b) t.order_simd.push_back(id);
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

int CompilePattern(const std::string& pattern, std::unique_ptr<Compiled>& out, std::string& err) {
  std::unique_ptr<Compiled> c(new Compiled());
  c->pattern = pattern;
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
  c->an = Analyze(pr.re, (int)c->prog.inst.size(), c->prog.anchored_start);

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

  // The anchored DFA is built for every pattern: ENG_DFA needs it, and it is also how the
  // record-delimiter safety of the pattern is proven.
  std::string de = BuildDFA(c->prog, /*anchored=*/true, /*max_states=*/160, c->dfa);
  if (c->kind == ENG_DFA) {
    if (!de.empty()) {
      err = "unsupported: " + de + " (PikeVM-kernel fallback for large automata is not built yet)";
      return COMPILE_UNSUPPORTED;
    }
    if (c->dfa.matches_empty || c->an.can_match_empty) {
      err = "unsupported: pattern can match the empty string (needs the sequential empty-match "
            "rules of meta/findall.go:247-279; not record-parallel)";
      return COMPILE_UNSUPPORTED;
    }
    if (!DelimiterSafe(c->dfa, '\n')) {
      err = "unsupported: a match can span '\\n' (record-parallel scan needs a delimiter no match "
            "can contain)";
      return COMPILE_UNSUPPORTED;
    }
    c->delim = '\n';
    for (int k = 1; k < SK_COUNT; k++)
      if (c->dfa.start[k] != c->dfa.start[0]) c->kind_lut_needed = true;
    memset(c->lut, 0, sizeof c->lut);
    if (c->an.strategy == RS_UseDigitPrefilter) {
      // reference meta/find_indices.go:1050-1088: candidates are ASCII digits; with
      // digitRunSkipSafe a failed candidate skips the rest of its digit run.
      c->nranges = 1;
      c->rlo[0] = '0';
      c->rhi[0] = '9';
      c->skip_safe = c->an.digit_run_skip_safe;
      c->filter_kind = c->skip_safe ? F_RUNSTART : F_BYTESET;
      c->engine_name = c->skip_safe ? "dfa-runstart" : "dfa-byteset";
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
    }
  }
  out = std::move(c);
  return COMPILE_OK;
}

}  // namespace cgx
