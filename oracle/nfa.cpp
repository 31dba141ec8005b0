// oracle/nfa.cpp — TEST INFRASTRUCTURE ONLY.  See nfa.h for the reference citations.
#include "nfa.h"

#include <cstdio>
#include <cstring>

namespace oracle {

using namespace gosyntax;

namespace {

// reference nfa/alphabet.go:111-166 — boundary bitset, classes by walking 0..255
struct ByteClassSet {
  bool bits[256] = {false};
  void set_range(uint8_t lo, uint8_t hi) {
    if (lo > 0) bits[lo - 1] = true;
    bits[hi] = true;
  }
};

struct Builder {
  std::vector<State> states;
  ByteClassSet bcs;

  StateID add(State s) {
    states.push_back(std::move(s));
    return (StateID)states.size() - 1;
  }
  StateID AddMatch() {
    State s;
    s.kind = StateMatch;
    return add(s);
  }
  StateID AddByteRange(uint8_t lo, uint8_t hi, StateID next) {
    bcs.set_range(lo, hi);
    State s;
    s.kind = StateByteRange;
    s.lo = lo;
    s.hi = hi;
    s.next = next;
    return add(s);
  }
  StateID AddSparse(const std::vector<Transition>& tr) {
    for (auto& t : tr) bcs.set_range(t.lo, t.hi);
    State s;
    s.kind = StateSparse;
    s.trans = tr;
    return add(s);
  }
  StateID AddSplit(StateID l, StateID r, bool quant = false) {
    State s;
    s.kind = StateSplit;
    s.left = l;
    s.right = r;
    s.quantifier_split = quant;
    return add(s);
  }
  StateID AddEpsilon(StateID next) {
    State s;
    s.kind = StateEpsilon;
    s.next = next;
    return add(s);
  }
  StateID AddCapture(uint32_t idx, bool is_start, StateID next) {
    State s;
    s.kind = StateCapture;
    s.cap_index = idx;
    s.cap_start = is_start;
    s.next = next;
    return add(s);
  }
  StateID AddLook(Look l, StateID next) {
    State s;
    s.kind = StateLook;
    s.look = l;
    s.next = next;
    return add(s);
  }
  // reference nfa/builder.go:189-210: only single-target kinds are patchable
  bool Patch(StateID id, StateID target) {
    if (id >= states.size()) return false;
    State& s = states[id];
    switch (s.kind) {
      case StateByteRange:
      case StateEpsilon:
      case StateCapture:
      case StateLook:
        s.next = target;
        return true;
      default:
        return false;
    }
  }
};

struct Frag {
  StateID start = InvalidState, end = InvalidState;
};

struct Compiler {
  Builder b;
  Arena scratch;  // synthetic literal nodes of small Unicode classes
  int depth = 0;
  int max_depth = 100;
  int capture_count = 0;
  std::vector<std::string> names;
  std::string err;
  std::vector<std::unique_ptr<Regexp>> synth;  // synthetic star/quest nodes for repeats

  bool fail(const std::string& e) {
    if (err.empty()) err = e;
    return false;
  }

  void countCaps(const Regexp* re) {
    switch (re->op) {
      case OpCapture:
        if (re->cap > capture_count) capture_count = re->cap;
        for (auto* s : re->sub) countCaps(s);
        break;
      case OpConcat:
      case OpAlternate:
        for (auto* s : re->sub) countCaps(s);
        break;
      case OpStar:
      case OpPlus:
      case OpQuest:
      case OpRepeat:
        if (!re->sub.empty()) countCaps(re->sub[0]);
        break;
      default:
        break;
    }
  }
  void collectNames(const Regexp* re) {
    switch (re->op) {
      case OpCapture:
        if (re->cap >= 0 && re->cap < (int)names.size()) names[re->cap] = re->name;
        for (auto* s : re->sub) collectNames(s);
        break;
      case OpConcat:
      case OpAlternate:
        for (auto* s : re->sub) collectNames(s);
        break;
      case OpStar:
      case OpPlus:
      case OpQuest:
      case OpRepeat:
        if (!re->sub.empty()) collectNames(re->sub[0]);
        break;
      default:
        break;
    }
  }

  static bool isPatternAnchored(const Regexp* re) {
    switch (re->op) {
      case OpBeginText:
        return true;
      case OpConcat:
      case OpCapture:
        if (!re->sub.empty()) return isPatternAnchored(re->sub[0]);
        return false;
      default:
        return false;
    }
  }

  static int encodeRune(uint8_t* buf, int32_t r) {
    if (r < 0x80) {
      buf[0] = (uint8_t)r;
      return 1;
    }
    if (r < 0x800) {
      buf[0] = 0xC0 | (r >> 6);
      buf[1] = 0x80 | (r & 0x3F);
      return 2;
    }
    if (r < 0x10000) {
      buf[0] = 0xE0 | (r >> 12);
      buf[1] = 0x80 | ((r >> 6) & 0x3F);
      buf[2] = 0x80 | (r & 0x3F);
      return 3;
    }
    buf[0] = 0xF0 | (r >> 18);
    buf[1] = 0x80 | ((r >> 12) & 0x3F);
    buf[2] = 0x80 | ((r >> 6) & 0x3F);
    buf[3] = 0x80 | (r & 0x3F);
    return 4;
  }

  bool compileSingleRune(int32_t r, Frag& f) {
    uint8_t buf[4];
    int n = encodeRune(buf, r);
    StateID prev = InvalidState, first = InvalidState;
    for (int i = 0; i < n; i++) {
      StateID id = b.AddByteRange(buf[i], buf[i], InvalidState);
      if (first == InvalidState) first = id;
      if (prev != InvalidState) b.Patch(prev, id);
      prev = id;
    }
    f = {first, prev};
    return true;
  }

  bool compileEmptyMatch(Frag& f) {
    StateID id = b.AddEpsilon(InvalidState);
    f = {id, id};
    return true;
  }
  bool compileNoMatch(Frag& f) {
    f.start = b.AddEpsilon(InvalidState);
    f.end = b.AddEpsilon(InvalidState);
    return true;
  }

  // reference nfa/compile.go:237-361
  bool compileLiteral(const Regexp* re, Frag& f) {
    if (re->rune.empty()) return compileEmptyMatch(f);
    bool fold = re->flags & FoldCase;
    StateID prev = InvalidState, first = InvalidState;
    for (int32_t r : re->rune) {
      bool letter = (r >= 'a' && r <= 'z') || (r >= 'A' && r <= 'Z');
      if (fold && letter) {
        int32_t upper = (r >= 'a' && r <= 'z') ? r - 32 : r;
        int32_t lower = (r >= 'A' && r <= 'Z') ? r + 32 : r;
        Frag u, l;
        compileSingleRune(upper, u);
        compileSingleRune(lower, l);
        StateID next = b.AddEpsilon(InvalidState);
        b.Patch(u.end, next);
        b.Patch(l.end, next);
        StateID split = b.AddSplit(u.start, l.start);
        if (prev == InvalidState)
          first = split;
        else
          b.Patch(prev, split);
        prev = next;
      } else {
        uint8_t buf[4];
        int n = encodeRune(buf, r);
        for (int i = 0; i < n; i++) {
          StateID id = b.AddByteRange(buf[i], buf[i], InvalidState);
          if (first == InvalidState) first = id;
          if (prev != InvalidState) b.Patch(prev, id);
          prev = id;
        }
      }
    }
    f = {first, prev};
    return true;
  }

  // reference nfa/compile.go:384-437 (ASCII only; Unicode branch is out of scope)
  bool compileCharClass(const std::vector<int32_t>& ranges, Frag& f) {
    if (ranges.empty()) return compileNoMatch(f);
    for (int32_t r : ranges)
      if (r > 127) return compileUnicodeClass(ranges, f);
    std::vector<Transition> tr;
    for (size_t i = 0; i + 1 < ranges.size(); i += 2)
      tr.push_back({(uint8_t)ranges[i], (uint8_t)ranges[i + 1], InvalidState});
    if (tr.size() == 1) {
      StateID id = b.AddByteRange(tr[0].lo, tr[0].hi, InvalidState);
      f = {id, id};
      return true;
    }
    StateID target = b.AddEpsilon(InvalidState);
    for (auto& t : tr) t.next = target;
    StateID id = b.AddSparse(tr);
    f = {id, target};
    return true;
  }

  // ---- UTF-8 automata ------------------------------------------------------------------------
  // reference nfa/compile.go:440-487: classes of at most 256 code points become an alternation of
  // single-rune literals, larger ones go through the range compiler
  bool compileUnicodeClass(const std::vector<int32_t>& ranges, Frag& f) {
    int64_t total = 0;
    for (size_t i = 0; i + 1 < ranges.size(); i += 2) {
      total += (int64_t)ranges[i + 1] - ranges[i] + 1;
      if (total > 256) return compileUnicodeClassLarge(ranges, f);
    }
    std::vector<Regexp*> alts;
    for (size_t i = 0; i + 1 < ranges.size(); i += 2)
      for (int32_t r = ranges[i]; r <= ranges[i + 1]; r++) {
        Regexp* lit = scratch.make(OpLiteral);
        lit->rune.push_back(r);
        alts.push_back(lit);
      }
    if (alts.size() == 1) return compile(alts[0], f);
    return compileAlternate(alts, f);
  }

  // reference nfa/compile.go:491-596
  bool compileUnicodeClassLarge(const std::vector<int32_t>& ranges, Frag& f) {
    std::vector<Transition> ascii;
    std::vector<std::pair<int32_t, int32_t>> wide;
    for (size_t i = 0; i + 1 < ranges.size(); i += 2) {
      const int32_t lo = ranges[i], hi = ranges[i + 1];
      if (hi < 0x80) {
        ascii.push_back({(uint8_t)lo, (uint8_t)hi, InvalidState});
      } else if (lo >= 0x80) {
        wide.push_back({lo, hi});
      } else {
        ascii.push_back({(uint8_t)lo, 0x7F, InvalidState});
        wide.push_back({0x80, hi});
      }
    }
    const bool all_wide = wide.size() == 1 && wide[0].first <= 0x80 && wide[0].second >= 0x10FFFF;
    StateID target = b.AddEpsilon(InvalidState);
    std::vector<StateID> starts;
    if (!ascii.empty()) {
      for (auto& t : ascii) t.next = target;
      starts.push_back(ascii.size() == 1 ? b.AddByteRange(ascii[0].lo, ascii[0].hi, target) : b.AddSparse(ascii));
    }
    if (!wide.empty()) {
      if (all_wide) {
        buildUTF8NonASCIIBranches(target, starts);
        starts.push_back(b.AddByteRange(0x80, 0xFF, target));  // any stray high byte, lowest priority
      } else {
        for (auto& w : wide) compileUTF8Range(w.first, w.second, target, starts);
      }
    }
    if (starts.empty()) return compileNoMatch(f);
    if (starts.size() == 1) {
      f = {starts[0], target};
      return true;
    }
    f = {buildSplitChain(starts), target};
    return true;
  }

  // reference nfa/compile.go:600-654: split [lo,hi] by encoded length
  void compileUTF8Range(int32_t lo, int32_t hi, StateID end, std::vector<StateID>& out) {
    if (lo <= 0x7F) {
      out.push_back(b.AddByteRange((uint8_t)lo, (uint8_t)(hi > 0x7F ? 0x7F : hi), end));
      lo = 0x80;
    }
    if (lo > hi) return;
    if (lo <= 0x7FF) {
      utf8Range2(lo, hi > 0x7FF ? 0x7FF : hi, end, out);
      lo = 0x800;
    }
    if (lo > hi) return;
    if (lo <= 0xFFFF) {
      utf8Range3(lo, hi > 0xFFFF ? 0xFFFF : hi, end, out);
      lo = 0x10000;
    }
    if (lo > hi) return;
    utf8Range4(lo, hi, end, out);
  }

  // reference nfa/compile.go:663-703
  void utf8Range2(int32_t lo, int32_t hi, StateID end, std::vector<StateID>& out) {
    const uint8_t l0 = 0xC0 | (lo >> 6), l1 = 0x80 | (lo & 0x3F), h0 = 0xC0 | (hi >> 6), h1 = 0x80 | (hi & 0x3F);
    if (l0 == h0) {
      StateID c = b.AddByteRange(l1, h1, end);
      out.push_back(b.AddByteRange(l0, l0, c));
      return;
    }
    StateID c1 = b.AddByteRange(l1, 0xBF, end);
    out.push_back(b.AddByteRange(l0, l0, c1));
    if (h0 > l0 + 1) {
      StateID cm = b.AddByteRange(0x80, 0xBF, end);
      out.push_back(b.AddByteRange(l0 + 1, h0 - 1, cm));
    }
    StateID c2 = b.AddByteRange(0x80, h1, end);
    out.push_back(b.AddByteRange(h0, h0, c2));
  }

  // reference nfa/compile.go:706-737: surrogates are cut out of three-byte ranges
  void utf8Range3(int32_t lo, int32_t hi, StateID end, std::vector<StateID>& out) {
    if (lo <= 0xD7FF && hi >= 0xE000) {
      utf8Range3Simple(lo, 0xD7FF, end, out);
      utf8Range3Simple(0xE000, hi, end, out);
      return;
    }
    if (lo >= 0xD800 && hi <= 0xDFFF) return;
    if (lo >= 0xD800 && lo <= 0xDFFF) lo = 0xE000;
    if (hi >= 0xD800 && hi <= 0xDFFF) hi = 0xD7FF;
    if (lo > hi) return;
    utf8Range3Simple(lo, hi, end, out);
  }

  // reference nfa/compile.go:740-793 (+ the bound helpers :922-973): one lead/cont1 pair per branch
  void utf8Range3Simple(int32_t lo, int32_t hi, StateID end, std::vector<StateID>& out) {
    const int l0 = 0xE0 | (lo >> 12), l1 = 0x80 | ((lo >> 6) & 0x3F), l2 = 0x80 | (lo & 0x3F);
    const int h0 = 0xE0 | (hi >> 12), h1 = 0x80 | ((hi >> 6) & 0x3F), h2 = 0x80 | (hi & 0x3F);
    auto branch = [&](int lead, int c1, int c2lo, int c2hi) {
      StateID s2 = b.AddByteRange((uint8_t)c2lo, (uint8_t)c2hi, end);
      StateID s1 = b.AddByteRange((uint8_t)c1, (uint8_t)c1, s2);
      out.push_back(b.AddByteRange((uint8_t)lead, (uint8_t)lead, s1));
    };
    if (l0 == h0 && l1 == h1) {
      branch(l0, l1, l2, h2);
    } else if (l0 == h0) {
      for (int c1 = l1; c1 <= h1; c1++) branch(l0, c1, c1 == l1 ? l2 : 0x80, c1 == h1 ? h2 : 0xBF);
    } else {
      for (int lead = l0; lead <= h0; lead++) {
        const int c1lo = lead == l0 ? l1 : (lead == 0xE0 ? 0xA0 : 0x80);
        const int c1hi = lead == h0 ? h1 : (lead == 0xED ? 0x9F : 0xBF);
        for (int c1 = c1lo; c1 <= c1hi; c1++)
          branch(lead, c1, (lead == l0 && c1 == l1) ? l2 : 0x80, (lead == h0 && c1 == h1) ? h2 : 0xBF);
      }
    }
  }

  // reference nfa/compile.go:796-842: four-byte ranges are widened to whole lead bytes
  void utf8Range4(int32_t lo, int32_t hi, StateID end, std::vector<StateID>& out) {
    if (hi > 0x10FFFF) hi = 0x10FFFF;
    if (lo < 0x10000) lo = 0x10000;
    if (lo > hi) return;
    const int l0 = 0xF0 | (lo >> 18), h0 = 0xF0 | (hi >> 18);
    for (int lead = l0; lead <= h0; lead++) {
      StateID s3 = b.AddByteRange(0x80, 0xBF, end);
      StateID s2 = b.AddByteRange(0x80, 0xBF, s3);
      StateID s1 = b.AddByteRange(lead == 0xF0 ? 0x90 : 0x80, lead == 0xF4 ? 0x8F : 0xBF, s2);
      out.push_back(b.AddByteRange((uint8_t)lead, (uint8_t)lead, s1));
    }
  }

  // reference nfa/compile.go:845-917: all valid multi-byte sequences, no state sharing
  void buildUTF8NonASCIIBranches(StateID end, std::vector<StateID>& out) {
    struct Seq { uint8_t n, r[4][2]; };
    static const Seq seqs[8] = {
        {2, {{0xC2, 0xDF}, {0x80, 0xBF}}},
        {3, {{0xE0, 0xE0}, {0xA0, 0xBF}, {0x80, 0xBF}}},
        {3, {{0xE1, 0xEC}, {0x80, 0xBF}, {0x80, 0xBF}}},
        {3, {{0xED, 0xED}, {0x80, 0x9F}, {0x80, 0xBF}}},
        {3, {{0xEE, 0xEF}, {0x80, 0xBF}, {0x80, 0xBF}}},
        {4, {{0xF0, 0xF0}, {0x90, 0xBF}, {0x80, 0xBF}, {0x80, 0xBF}}},
        {4, {{0xF1, 0xF3}, {0x80, 0xBF}, {0x80, 0xBF}, {0x80, 0xBF}}},
        {4, {{0xF4, 0xF4}, {0x80, 0x8F}, {0x80, 0xBF}, {0x80, 0xBF}}},
    };
    for (const Seq& q : seqs) {
      StateID t = end;
      for (int i = q.n - 1; i >= 0; i--) t = b.AddByteRange(q.r[i][0], q.r[i][1], t);
      out.push_back(t);
    }
  }

  // reference nfa/utf8_suffix.go:30-123: direct-mapped 64-entry cache keyed by (target, lo, hi);
  // a colliding key simply overwrites, so which suffix states are shared depends on the hash
  struct SuffixCache {
    struct E { bool used = false; StateID from = 0; uint8_t lo = 0, hi = 0; StateID val = 0; };
    E e[64];
    static int slot(StateID from, uint8_t lo, uint8_t hi) {
      uint64_t h = 14695981039346656037ull;
      h = (h ^ (uint64_t)from) * 1099511628211ull;
      h = (h ^ (uint64_t)lo) * 1099511628211ull;
      h = (h ^ (uint64_t)hi) * 1099511628211ull;
      return (int)(h % 64);
    }
  };
  StateID suffixState(SuffixCache& c, StateID target, uint8_t lo, uint8_t hi) {
    auto& x = c.e[SuffixCache::slot(target, lo, hi)];
    if (x.used && x.from == target && x.lo == lo && x.hi == hi) return x.val;
    StateID s = b.AddByteRange(lo, hi, target);
    x.used = true; x.from = target; x.lo = lo; x.hi = hi; x.val = s;
    return s;
  }

  // reference nfa/compile.go:1142-1223: `.` / (?s:.) = ASCII | valid multi-byte sequences (suffix
  // states shared through the cache) | single stray bytes 80-BF, C0-C1, F5-FF
  bool compileUTF8Any(bool include_nl, Frag& f) {
    StateID end = b.AddEpsilon(InvalidState);
    SuffixCache cache;
    std::vector<StateID> br;
    if (include_nl) {
      br.push_back(b.AddByteRange(0x00, 0x7F, end));
    } else {
      br.push_back(b.AddSparse({{0x00, 0x09, end}, {0x0B, 0x7F, end}}));
    }
    static const uint8_t seqs[8][9] = {
        {2, 0xC2, 0xDF, 0x80, 0xBF}, {3, 0xE0, 0xE0, 0xA0, 0xBF, 0x80, 0xBF}, {3, 0xE1, 0xEC, 0x80, 0xBF, 0x80, 0xBF},
        {3, 0xED, 0xED, 0x80, 0x9F, 0x80, 0xBF}, {3, 0xEE, 0xEF, 0x80, 0xBF, 0x80, 0xBF},
        {4, 0xF0, 0xF0, 0x90, 0xBF, 0x80, 0xBF, 0x80, 0xBF}, {4, 0xF1, 0xF3, 0x80, 0xBF, 0x80, 0xBF, 0x80, 0xBF},
        {4, 0xF4, 0xF4, 0x80, 0x8F, 0x80, 0xBF, 0x80, 0xBF}};
    for (auto& q : seqs) {
      StateID t = end;
      for (int i = q[0] - 1; i >= 0; i--) t = suffixState(cache, t, q[1 + 2 * i], q[2 + 2 * i]);
      br.push_back(t);
    }
    br.push_back(b.AddSparse({{0x80, 0xBF, end}, {0xC0, 0xC1, end}, {0xF5, 0xFF, end}}));
    f = {buildSplitChain(br), end};
    return true;
  }

  bool compileConcat(const std::vector<Regexp*>& subs, Frag& f) {
    if (subs.empty()) return compileEmptyMatch(f);
    if (subs.size() == 1) return compile(subs[0], f);
    if (!compile(subs[0], f)) return false;
    for (size_t i = 1; i < subs.size(); i++) {
      Frag n;
      if (!compile(subs[i], n)) return false;
      if (!b.Patch(f.end, n.start)) {
        StateID eps = b.AddEpsilon(n.start);
        if (!b.Patch(f.end, eps)) return fail("concat patch failed");
      }
      f.end = n.end;
    }
    return true;
  }

  StateID buildSplitChain(const std::vector<StateID>& t, size_t from = 0) {
    size_t n = t.size() - from;
    if (n == 1) return t[from];
    if (n == 2) return b.AddSplit(t[from], t[from + 1]);
    StateID right = buildSplitChain(t, from + 1);
    return b.AddSplit(t[from], right);
  }

  bool compileAlternate(const std::vector<Regexp*>& subs, Frag& f) {
    if (subs.empty()) return compileEmptyMatch(f);
    if (subs.size() == 1) return compile(subs[0], f);
    std::vector<StateID> starts, ends;
    for (auto* s : subs) {
      Frag x;
      if (!compile(s, x)) return false;
      starts.push_back(x.start);
      ends.push_back(x.end);
    }
    StateID split = buildSplitChain(starts);
    StateID join = b.AddEpsilon(InvalidState);
    for (StateID e : ends) b.Patch(e, join);  // failures ignored, as in the reference
    f = {split, join};
    return true;
  }

  // reference nfa/compile.go:1390-1431
  static bool canMatchEmpty(const Regexp* re) {
    switch (re->op) {
      case OpEmptyMatch: return true;
      case OpLiteral: return re->rune.empty();
      case OpCharClass: case OpAnyCharNotNL: case OpAnyChar: return false;
      case OpCapture: return re->sub.empty() ? true : canMatchEmpty(re->sub[0]);
      case OpStar: case OpQuest: return true;
      case OpPlus: return !re->sub.empty() && canMatchEmpty(re->sub[0]);
      case OpRepeat: return re->min == 0 || (!re->sub.empty() && canMatchEmpty(re->sub[0]));
      case OpConcat:
        for (auto* s : re->sub) if (!canMatchEmpty(s)) return false;
        return true;
      case OpAlternate:
        for (auto* s : re->sub) if (canMatchEmpty(s)) return true;
        return false;
      case OpNoMatch: return false;
      case OpBeginLine: case OpEndLine: case OpBeginText: case OpEndText:
      case OpWordBoundary: case OpNoWordBoundary: return true;
      default: return false;
    }
  }

  bool loopBack(StateID subEnd, StateID split) {
    if (!b.Patch(subEnd, split)) {
      StateID eps = b.AddEpsilon(split);
      if (!b.Patch(subEnd, eps)) return fail("loop patch failed");
    }
    return true;
  }

  bool compileStar(const Regexp* sub, bool ng, Frag& f) {
    if (canMatchEmpty(sub)) {
      // (x+)? form — reference nfa/compile.go:1353-1387
      Frag s;
      if (!compile(sub, s)) return false;
      StateID end = b.AddEpsilon(InvalidState);
      StateID plus = ng ? b.AddSplit(end, s.start, true) : b.AddSplit(s.start, end, true);
      if (!loopBack(s.end, plus)) return false;
      StateID q = ng ? b.AddSplit(end, s.start, true) : b.AddSplit(s.start, end, true);
      f = {q, end};
      return true;
    }
    Frag s;
    if (!compile(sub, s)) return false;
    StateID end = b.AddEpsilon(InvalidState);
    StateID split = ng ? b.AddSplit(end, s.start, true) : b.AddSplit(s.start, end, true);
    if (!loopBack(s.end, split)) return false;
    f = {split, end};
    return true;
  }

  bool compilePlus(const Regexp* sub, bool ng, Frag& f) {
    Frag s;
    if (!compile(sub, s)) return false;
    StateID end = b.AddEpsilon(InvalidState);
    StateID split = ng ? b.AddSplit(end, s.start, true) : b.AddSplit(s.start, end, true);
    if (!loopBack(s.end, split)) return false;
    f = {s.start, end};
    return true;
  }

  bool compileQuest(const Regexp* sub, bool ng, Frag& f) {
    Frag s;
    if (!compile(sub, s)) return false;
    StateID end = b.AddEpsilon(InvalidState);
    StateID split = ng ? b.AddSplit(end, s.start, true) : b.AddSplit(s.start, end, true);
    if (!b.Patch(s.end, end)) {
      StateID eps = b.AddEpsilon(end);
      if (!b.Patch(s.end, eps)) return fail("quest patch failed");
    }
    f = {split, end};
    return true;
  }

  Regexp* synthNode(Op op, bool ng, Regexp* sub) {
    synth.emplace_back(new Regexp());
    Regexp* r = synth.back().get();
    r->op = op;
    r->flags = ng ? NonGreedy : 0;
    r->sub.assign(1, sub);
    return r;
  }

  // reference nfa/compile.go:1485-1565
  bool compileRepeat(Regexp* sub, int mn, int mx, bool ng, Frag& f) {
    std::vector<Regexp*> subs;
    if (mx == -1) {
      if (mn == 0) return compileStar(sub, ng, f);
      for (int i = 0; i < mn; i++) subs.push_back(sub);
      subs.push_back(synthNode(OpStar, ng, sub));
      return compileConcat(subs, f);
    }
    if (mn == mx) {
      if (mn == 0) return compileEmptyMatch(f);
      if (mn == 1) return compile(sub, f);
      for (int i = 0; i < mn; i++) subs.push_back(sub);
      return compileConcat(subs, f);
    }
    if (mn > mx) return fail("invalid repeat range");
    for (int i = 0; i < mn; i++) subs.push_back(sub);
    for (int i = 0; i < mx - mn; i++) subs.push_back(synthNode(OpQuest, ng, sub));
    return compileConcat(subs, f);
  }

  bool compileCapture(const Regexp* re, Frag& f) {
    if (re->sub.empty()) return compileEmptyMatch(f);
    Frag s;
    if (!compile(re->sub[0], s)) return false;
    StateID close = b.AddCapture((uint32_t)re->cap, false, InvalidState);
    if (!b.Patch(s.end, close)) {
      StateID eps = b.AddEpsilon(close);
      if (!b.Patch(s.end, eps)) return fail("capture patch failed");
    }
    StateID open = b.AddCapture((uint32_t)re->cap, true, s.start);
    f = {open, close};
    return true;
  }

  bool look(Look l, Frag& f) {
    StateID id = b.AddLook(l, InvalidState);
    f = {id, id};
    return true;
  }

  // reference nfa/compile.go:167-233
  bool compile(const Regexp* re, Frag& f) {
    depth++;
    struct D {
      int& d;
      ~D() { d--; }
    } guard{depth};
    if (depth > max_depth) return fail("regex too complex");
    bool ng = re->flags & NonGreedy;
    switch (re->op) {
      case OpLiteral: return compileLiteral(re, f);
      case OpCharClass: return compileCharClass(re->rune, f);
      case OpAnyChar: return compileUTF8Any(true, f);        // reference nfa/compile.go:977-985
      case OpAnyCharNotNL: return compileUTF8Any(false, f);  // reference nfa/compile.go:995-1003
      case OpConcat: return compileConcat(re->sub, f);
      case OpAlternate: return compileAlternate(re->sub, f);
      case OpStar: return compileStar(re->sub[0], ng, f);
      case OpPlus: return compilePlus(re->sub[0], ng, f);
      case OpQuest: return compileQuest(re->sub[0], ng, f);
      case OpRepeat: return compileRepeat(re->sub[0], re->min, re->max, ng, f);
      case OpCapture: return compileCapture(re, f);
      case OpBeginText: return look(LookStartText, f);
      case OpEndText: return look(LookEndText, f);
      case OpBeginLine: return look(LookStartLine, f);
      case OpEndLine: return look(LookEndLine, f);
      case OpWordBoundary: return look(LookWordBoundary, f);
      case OpNoWordBoundary: return look(LookNoWordBoundary, f);
      case OpEmptyMatch: return compileEmptyMatch(f);
      default:
        return fail("unsupported regex operation");
    }
  }
};

}  // namespace

std::string CompileNFA(const Regexp* re, bool anchored_cfg, NFA& out) {
  Compiler c;
  c.countCaps(re);
  c.names.assign(c.capture_count + 1, "");
  c.collectNames(re);
  bool allAnchored = Compiler::isPatternAnchored(re);

  Frag p;
  if (!c.compile(re, p)) return c.err.empty() ? "compile failed" : c.err;
  StateID match = c.b.AddMatch();
  if (!c.b.Patch(p.end, match)) {
    StateID eps = c.b.AddEpsilon(match);
    if (!c.b.Patch(p.end, eps)) return "failed to connect to match state";
  }
  StateID anchoredStart = p.start;
  StateID unanchoredStart;
  if (anchored_cfg || allAnchored) {
    unanchoredStart = anchoredStart;
  } else {
    // reference nfa/compile.go:1633-1650: (?s:.)*? prefix
    StateID any = c.b.AddByteRange(0x00, 0xFF, InvalidState);
    StateID split = c.b.AddSplit(p.start, any);
    c.b.Patch(any, split);
    unanchoredStart = split;
  }
  out.states = std::move(c.b.states);
  out.start_anchored = anchoredStart;
  out.start_unanchored = unanchoredStart;
  out.anchored = anchored_cfg || allAnchored;
  out.capture_count = c.capture_count + 1;
  out.capture_names = c.names;
  uint8_t cls = 0;
  for (int i = 0; i < 256; i++) {
    out.byte_classes[i] = cls;
    if (c.b.bcs.bits[i]) cls++;
  }
  int mx = 0;
  for (int i = 0; i < 256; i++)
    if (out.byte_classes[i] > mx) mx = out.byte_classes[i];
  out.alphabet_len = mx + 1;
  return "";
}

// ---- reference nfa/reverse.go ------------------------------------------------------------------------
namespace {
enum EdgeKind : uint8_t { edgeByteRange = 0, edgeSparse, edgeEpsilon };
struct RevEdge {
  StateID from;
  EdgeKind kind;
  uint8_t lo, hi;
};
struct Reverser {
  Builder b;
  std::vector<std::vector<RevEdge>> edges;  // reverseEdges[to]
  std::vector<StateID> map;                 // revStateMap (InvalidState = no entry)
  std::vector<bool> has;

  StateID AddFail() {
    State s;
    s.kind = StateFail;
    return b.add(s);
  }
  // Go's map read of a missing key yields 0 (the reverse match state): reference code that does
  // not check `ok` gets exactly that
  StateID get0(StateID f) const { return has[f] ? map[f] : 0; }
  void put(StateID f, StateID r) {
    map[f] = r;
    has[f] = true;
  }
  bool PatchSplit(StateID id, StateID l, StateID r) {
    if (id >= b.states.size() || b.states[id].kind != StateSplit) return false;
    b.states[id].left = l;
    b.states[id].right = r;
    return true;
  }
  // :598-614
  StateID buildSplitChain(const std::vector<StateID>& t, size_t from = 0) {
    const size_t n = t.size() - from;
    if (n == 0) return AddFail();
    if (n == 1) return t[from];
    if (n == 2) return b.AddSplit(t[from], t[from + 1]);
    StateID right = buildSplitChain(t, from + 1);
    return b.AddSplit(t[from], right);
  }
  // :329-367
  StateID allocatePlaceholder(const std::vector<RevEdge>& e) {
    if (e.empty()) return AddFail();
    int br = 0, eps = 0;
    for (auto& x : e) (x.kind == edgeEpsilon ? eps : br)++;
    if (br == 0 && eps > 0) return eps == 1 ? b.AddEpsilon(InvalidState) : b.AddSplit(InvalidState, InvalidState);
    if (br == 1 && eps == 0) return b.AddByteRange(0, 0, InvalidState);
    if (br > 0 && eps > 0) return b.AddSplit(InvalidState, InvalidState);
    return b.AddSparse({Transition{0, 0, InvalidState}});
  }
  // :553-574
  void updateByteRangeState(StateID id, uint8_t lo, uint8_t hi, StateID next) {
    if (id >= b.states.size()) return;
    b.bcs.set_range(lo, hi);
    State& s = b.states[id];
    if (s.kind == StateByteRange || s.kind == StateEpsilon) {
      s.kind = StateByteRange;
      s.lo = lo;
      s.hi = hi;
      s.next = next;
    }
  }
  // :576-596
  void updateSparseState(StateID id, const std::vector<Transition>& tr) {
    if (id >= b.states.size()) return;
    for (auto& t : tr) b.bcs.set_range(t.lo, t.hi);
    State& s = b.states[id];
    if (s.kind == StateSparse || s.kind == StateByteRange || s.kind == StateEpsilon) {
      s.kind = StateSparse;
      s.trans = tr;
    }
  }
  // :488-514
  void fillSparseState(StateID id, const std::vector<RevEdge>& br) {
    std::vector<Transition> tr;
    for (auto& e : br) tr.push_back(Transition{e.lo, e.hi, get0(e.from)});
    updateSparseState(id, tr);
  }
  // :446-486
  void fillEpsilonState(StateID id, const std::vector<RevEdge>& eps) {
    if (eps.size() == 1) {
      b.Patch(id, get0(eps[0].from));
      return;
    }
    std::vector<StateID> t;
    for (auto& e : eps)
      if (has[e.from]) t.push_back(map[e.from]);
    if (t.empty()) return;
    if (t.size() == 1) {
      b.Patch(id, t[0]);
      return;
    }
    if (t.size() == 2) {
      PatchSplit(id, t[0], t[1]);
      return;
    }
    StateID right = buildSplitChain(t, 1);
    PatchSplit(id, t[0], right);
  }
  // :516-551
  void fillMixedState(StateID id, const std::vector<RevEdge>& br, const std::vector<RevEdge>& eps) {
    StateID sparse = b.AddSparse({Transition{0, 0, InvalidState}});
    fillSparseState(sparse, br);
    std::vector<StateID> t;
    for (auto& e : eps)
      if (has[e.from]) t.push_back(map[e.from]);
    StateID right;
    if (t.empty()) {
      updateSparseState(id, {});
      fillSparseState(id, br);
      return;
    } else if (t.size() == 1) {
      right = t[0];
    } else {
      right = buildSplitChain(t);
    }
    PatchSplit(id, sparse, right);
  }
  // :369-410
  void fillReverseState(StateID id, const std::vector<RevEdge>& e) {
    if (e.empty()) return;
    std::vector<RevEdge> br, eps;
    for (auto& x : e) (x.kind == edgeEpsilon ? eps : br).push_back(x);
    if (br.empty()) return fillEpsilonState(id, eps);
    if (br.size() == 1 && eps.empty()) return updateByteRangeState(id, br[0].lo, br[0].hi, get0(br[0].from));
    if (!eps.empty()) return fillMixedState(id, br, eps);
    fillSparseState(id, br);
  }
  // :412-444
  void fillStartStateWithIncoming(StateID proxy, const std::vector<RevEdge>& e, StateID matchID) {
    std::vector<StateID> t;
    for (auto& x : e)
      if (has[x.from]) t.push_back(map[x.from]);
    if (t.empty()) return;
    const StateID left = t.size() == 1 ? t[0] : buildSplitChain(t);
    State& s = b.states[proxy];
    s.kind = StateSplit;
    s.left = left;
    s.right = matchID;
    s.next = InvalidState;
  }
};
}  // namespace

void ReverseNFAStates(const NFA& fwd, bool anchored, NFA& out) {
  Reverser r;
  const size_t ns = fwd.states.size();
  r.edges.assign(ns, {});
  r.map.assign(ns, InvalidState);
  r.has.assign(ns, false);
  // :83-137 collectReverseEdges
  for (StateID from = 0; from < ns; from++) {
    const State& s = fwd.states[from];
    auto eps = [&](StateID to) {
      if (to != InvalidState) r.edges[to].push_back(RevEdge{from, edgeEpsilon, 0, 0});
    };
    switch (s.kind) {
      case StateByteRange:
        if (s.next != InvalidState) r.edges[s.next].push_back(RevEdge{from, edgeByteRange, s.lo, s.hi});
        break;
      case StateSparse:
        for (auto& t : s.trans)
          if (t.next != InvalidState) r.edges[t.next].push_back(RevEdge{from, edgeSparse, t.lo, t.hi});
        break;
      case StateSplit:
        eps(s.left);
        eps(s.right);
        break;
      case StateEpsilon:
      case StateCapture:
      case StateLook:  // zero-width: an epsilon edge (the reverse automaton checks no assertion)
        eps(s.next);
        break;
      default:
        break;
    }
  }
  const StateID matchID = r.b.AddMatch();
  const StateID fa = fwd.start_anchored, fu = fwd.start_unanchored;
  // :139-160 mapStartStates / mapSingleStartState
  r.put(fa, !r.edges[fa].empty() ? r.b.AddEpsilon(matchID) : matchID);
  if (!anchored && fu != fa) r.put(fu, !r.edges[fu].empty() ? r.b.AddEpsilon(matchID) : matchID);
  // :178-238 findUnanchoredPrefixStates / findLoopStates
  std::vector<bool> skip(ns, false);
  if (anchored && fu != fa) {
    skip[fu] = true;
    if (fwd.states[fu].kind == StateSplit) {
      const StateID start = fwd.states[fu].right;
      if (start != InvalidState && !skip[start]) {
        const State& s = fwd.states[start];
        if ((s.kind == StateByteRange && s.next == fu) || (s.kind == StateSplit && (s.left == fu || s.right == fu)) ||
            (s.kind == StateEpsilon && s.next == fu))
          skip[start] = true;
      }
    }
  }
  // :162-176 allocatePlaceholders
  for (StateID id = 0; id < ns; id++) {
    if (skip[id]) continue;
    if (!r.has[id]) r.put(id, r.allocatePlaceholder(r.edges[id]));
  }
  // :240-275 fillAllTransitions
  for (StateID id = 0; id < ns; id++) {
    if (skip[id]) continue;
    const bool isStart = id == fa || (!anchored && id == fu);
    const bool hasIncoming = !r.edges[id].empty();
    if (isStart && !hasIncoming) continue;
    if (!r.has[id]) continue;
    if (isStart && hasIncoming) r.fillStartStateWithIncoming(r.map[id], r.edges[id], matchID);
    else r.fillReverseState(r.map[id], r.edges[id]);
  }
  // :277-287 collectMatchStates, :616-634 buildReverseStarts
  std::vector<StateID> starts;
  size_t nmatch = 0;
  StateID single = InvalidState;
  for (StateID id = 0; id < ns; id++)
    if (fwd.states[id].kind == StateMatch) {
      nmatch++;
      single = id;
      if (r.has[id]) starts.push_back(r.map[id]);
    }
  StateID start;
  if (nmatch == 0) start = r.AddFail();
  else if (nmatch == 1) start = r.get0(single);
  else start = r.buildSplitChain(starts);
  // :289-308 buildFinalNFA
  out = NFA();
  out.states = std::move(r.b.states);
  out.start_anchored = out.start_unanchored = start;
  out.anchored = anchored || fwd.anchored;
  out.capture_count = 0;
  uint8_t cls = 0;
  for (int i = 0; i < 256; i++) {
    out.byte_classes[i] = cls;
    if (r.b.bcs.bits[i]) cls++;
  }
  int mx = 0;
  for (int i = 0; i < 256; i++)
    if (out.byte_classes[i] > mx) mx = out.byte_classes[i];
  out.alphabet_len = mx + 1;
}

std::string DumpNFA(const NFA& n) {
  std::string s;
  char buf[128];
  for (size_t i = 0; i < n.states.size(); i++) {
    const State& st = n.states[i];
    switch (st.kind) {
      case StateMatch: snprintf(buf, sizeof buf, "%zu Match\n", i); break;
      case StateByteRange:
        snprintf(buf, sizeof buf, "%zu BR[%02X-%02X]->%d\n", i, st.lo, st.hi, (int)st.next);
        break;
      case StateSparse: {
        snprintf(buf, sizeof buf, "%zu Sparse", i);
        s += buf;
        for (auto& t : st.trans) {
          snprintf(buf, sizeof buf, " [%02X-%02X]->%d", t.lo, t.hi, (int)t.next);
          s += buf;
        }
        snprintf(buf, sizeof buf, "\n");
        break;
      }
      case StateSplit:
        snprintf(buf, sizeof buf, "%zu %sSplit(%d,%d)\n", i, st.quantifier_split ? "Q" : "",
                 (int)st.left, (int)st.right);
        break;
      case StateEpsilon: snprintf(buf, sizeof buf, "%zu Eps->%d\n", i, (int)st.next); break;
      case StateCapture:
        snprintf(buf, sizeof buf, "%zu Cap(%u,%s)->%d\n", i, st.cap_index,
                 st.cap_start ? "open" : "close", (int)st.next);
        break;
      case StateLook: snprintf(buf, sizeof buf, "%zu Look(%d)->%d\n", i, st.look, (int)st.next); break;
      default: snprintf(buf, sizeof buf, "%zu Fail\n", i); break;
    }
    s += buf;
  }
  snprintf(buf, sizeof buf, "startAnchored=%d startUnanchored=%d classes=%d\n",
           (int)n.start_anchored, (int)n.start_unanchored, n.alphabet_len);
  s += buf;
  return s;
}

}  // namespace oracle
