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
      if (r > 127) return fail("unsupported: non-ASCII character class (UTF-8 automata out of scope)");
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
      case OpAnyChar:
      case OpAnyCharNotNL:
        return fail("unsupported: `.` (UTF-8 automata out of scope)");
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
