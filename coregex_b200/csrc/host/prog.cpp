// prog.cpp — AST -> Pike-style program (continuation-passing construction, built back to front).
#include "prog.h"

#include <functional>
#include <map>
#include <memory>

namespace cgx {

using namespace gosyntax;

namespace {

struct Builder {
  Prog& p;
  std::string err;
  bool reverse = false;
  int depth = 0;
  std::vector<std::unique_ptr<Regexp>> synth;

  int emit(const Inst& i) {
    p.inst.push_back(i);
    return (int)p.inst.size() - 1;
  }
  int emitSet(const ByteSet& s, int next) {
    // dedupe identical sets
    uint16_t idx = 0;
    for (; idx < p.sets.size(); idx++)
      if (p.sets[idx] == s) break;
    if (idx == p.sets.size()) p.sets.push_back(s);
    Inst i;
    i.op = I_SET;
    i.set = idx;
    i.out = next;
    return emit(i);
  }
  int emitSplit(int a, int b) {
    Inst i;
    i.op = I_SPLIT;
    i.out = a;
    i.out1 = b;
    return emit(i);
  }
  int fail(const std::string& e) {
    if (err.empty()) err = e;
    return -1;
  }

  static int encodeRune(uint8_t* buf, int32_t r) {
    if (r < 0x80) { buf[0] = (uint8_t)r; return 1; }
    if (r < 0x800) { buf[0] = 0xC0 | (r >> 6); buf[1] = 0x80 | (r & 0x3F); return 2; }
    if (r < 0x10000) {
      buf[0] = 0xE0 | (r >> 12); buf[1] = 0x80 | ((r >> 6) & 0x3F); buf[2] = 0x80 | (r & 0x3F);
      return 3;
    }
    buf[0] = 0xF0 | (r >> 18); buf[1] = 0x80 | ((r >> 12) & 0x3F);
    buf[2] = 0x80 | ((r >> 6) & 0x3F); buf[3] = 0x80 | (r & 0x3F);
    return 4;
  }

  static bool nullable(const Regexp* re) {  // reference nfa/compile.go:1390-1431
    switch (re->op) {
      case OpEmptyMatch: return true;
      case OpLiteral: return re->rune.empty();
      case OpCharClass: case OpAnyCharNotNL: case OpAnyChar: case OpNoMatch: return false;
      case OpCapture: return re->sub.empty() || nullable(re->sub[0]);
      case OpStar: case OpQuest: return true;
      case OpPlus: return !re->sub.empty() && nullable(re->sub[0]);
      case OpRepeat: return re->min == 0 || (!re->sub.empty() && nullable(re->sub[0]));
      case OpConcat:
        for (auto* s : re->sub) if (!nullable(s)) return false;
        return true;
      case OpAlternate:
        for (auto* s : re->sub) if (nullable(s)) return true;
        return false;
      default: return true;  // assertions
    }
  }

  int literal(const Regexp* re, int next) {
    // emit bytes back to front (front to back when building the reversed language)
    std::vector<std::pair<ByteSet, int>> bytes;  // per byte position: set
    std::vector<ByteSet> seq;
    for (int32_t r : re->rune) {
      bool letter = (r >= 'a' && r <= 'z') || (r >= 'A' && r <= 'Z');
      if ((re->flags & FoldCase) && letter) {
        ByteSet s{};
        unsigned u = (r >= 'a') ? r - 32 : r, l = (r <= 'Z') ? r + 32 : r;
        set_add(s, u, u);
        set_add(s, l, l);
        seq.push_back(s);
      } else {
        uint8_t buf[4];
        int n = encodeRune(buf, r);
        for (int i = 0; i < n; i++) {
          ByteSet s{};
          set_add(s, buf[i], buf[i]);
          seq.push_back(s);
        }
      }
    }
    if (!reverse) {
      for (size_t i = seq.size(); i-- > 0;) next = emitSet(seq[i], next);
    } else {
      for (size_t i = 0; i < seq.size(); i++) next = emitSet(seq[i], next);
    }
    return next;
  }

  int charClass(const Regexp* re, int next) {
    if (re->rune.empty()) {
      Inst i;
      i.op = I_FAIL;
      return emit(i);
    }
    bool ascii = true;
    for (int32_t r : re->rune) ascii = ascii && r <= 127;
    if (!ascii) return unicodeClass(re->rune, next);
    ByteSet s{};
    for (size_t k = 0; k + 1 < re->rune.size(); k += 2)
      set_add(s, (unsigned)re->rune[k], (unsigned)re->rune[k + 1]);
    return emitSet(s, next);
  }

  // ---- UTF-8: a class or `.` becomes an ordered alternation of byte-set sequences --------------
  // What has to equal the reference (nfa/compile.go:440-1223) is the byte LANGUAGE and the order
  // of alternatives that can overlap; the state shape is this program's own.
  using Seq = std::vector<ByteSet>;
  static ByteSet range(unsigned lo, unsigned hi) {
    ByteSet s{};
    set_add(s, lo, hi);
    return s;
  }
  int emitAlts(const std::vector<Seq>& alts, int next) {
    if (alts.empty()) {
      Inst i;
      i.op = I_FAIL;
      return emit(i);
    }
    std::vector<int> entries;
    for (const Seq& q : alts) {
      int nx = next;
      if (!reverse) {
        for (size_t i = q.size(); i-- > 0;) nx = emitSet(q[i], nx);
      } else {
        for (size_t i = 0; i < q.size(); i++) nx = emitSet(q[i], nx);
      }
      entries.push_back(nx);
    }
    int acc = entries.back();
    for (size_t i = entries.size() - 1; i-- > 0;) acc = emitSplit(entries[i], acc);
    return acc;
  }
  // Large classes (\pL: 650 ranges, 1500 sequences): the alternatives are UTF-8 sequences of
  // DISJOINT rune ranges, and the lead byte fixes the length, so whichever alternative takes a
  // position takes the same bytes — their order is not observable.  That allows sharing common
  // prefixes (a trie over byte sets) and common suffixes (equal subtrees emitted once): the same
  // byte language in a fraction of the instructions, which is what keeps such classes inside the
  // limits of the DFA builder and of the PikeVM search kernel.
  int emitShared(const std::vector<Seq>& alts, int next) {
    struct Node {
      std::vector<std::pair<ByteSet, int>> kids;  // child node, -1 = the sequence ends here
    };
    std::vector<Node> trie(1);
    for (const Seq& q0 : alts) {
      Seq q = q0;
      if (reverse) std::reverse(q.begin(), q.end());
      int at = 0;
      for (size_t i = 0; i < q.size(); i++) {
        const bool last = i + 1 == q.size();
        int hit = -2;
        for (auto& k : trie[at].kids)
          if (k.first == q[i]) hit = k.second;
        if (hit == -2) {
          hit = last ? -1 : (int)trie.size();
          trie[at].kids.push_back({q[i], hit});
          if (!last) trie.emplace_back();
        } else if ((hit == -1) != last) {
          return emitAlts(alts, next);  // one sequence a proper prefix of another: keep the plain form
        }
        at = hit;
      }
    }
    using Edge = std::pair<ByteSet, int>;             // byte set, target instruction
    std::map<Edge, int> edge_memo;                    // equal edges are one instruction
    std::map<std::vector<Edge>, int> node_memo;       // equal subtrees are one split chain
    std::function<int(int)> emitNode = [&](int n) -> int {
      std::vector<Edge> key;
      for (auto& k : trie[n].kids) key.push_back({k.first, k.second < 0 ? next : emitNode(k.second)});
      auto it = node_memo.find(key);
      if (it != node_memo.end()) return it->second;
      std::vector<int> entries;
      for (const Edge& e : key) {
        auto em = edge_memo.find(e);
        if (em == edge_memo.end()) em = edge_memo.emplace(e, emitSet(e.first, e.second)).first;
        entries.push_back(em->second);
      }
      int acc = entries.back();
      for (size_t i = entries.size() - 1; i-- > 0;) acc = emitSplit(entries[i], acc);
      node_memo[key] = acc;
      return acc;
    };
    return emitNode(0);
  }
  // every well-formed multi-byte sequence (the eight rows of the Unicode standard's table 3-7)
  static void validMultiByte(std::vector<Seq>& alts) {
    const ByteSet cont = range(0x80, 0xBF);
    alts.push_back({range(0xC2, 0xDF), cont});
    alts.push_back({range(0xE0, 0xE0), range(0xA0, 0xBF), cont});
    alts.push_back({range(0xE1, 0xEC), cont, cont});
    alts.push_back({range(0xED, 0xED), range(0x80, 0x9F), cont});
    alts.push_back({range(0xEE, 0xEF), cont, cont});
    alts.push_back({range(0xF0, 0xF0), range(0x90, 0xBF), cont, cont});
    alts.push_back({range(0xF1, 0xF3), cont, cont, cont});
    alts.push_back({range(0xF4, 0xF4), range(0x80, 0x8F), cont, cont});
  }
  // [lo,hi] inside one encoded length -> sequences of byte ranges (each a product of ranges)
  static void splitRange(int32_t lo, int32_t hi, int nbytes, std::vector<Seq>& alts) {
    for (int i = 1; i < nbytes; i++) {
      const int32_t m = (1 << (6 * i)) - 1;
      if ((lo & ~m) != (hi & ~m)) {
        if ((lo & m) != 0) {
          splitRange(lo, lo | m, nbytes, alts);
          splitRange((lo | m) + 1, hi, nbytes, alts);
          return;
        }
        if ((hi & m) != m) {
          splitRange(lo, (hi & ~m) - 1, nbytes, alts);
          splitRange(hi & ~m, hi, nbytes, alts);
          return;
        }
      }
    }
    uint8_t a[4], b[4];
    encodeRune(a, lo);
    encodeRune(b, hi);
    Seq q;
    for (int i = 0; i < nbytes; i++) q.push_back(range(a[i], b[i]));
    alts.push_back(q);
  }
  // reference nfa/compile.go:600-842 as a language: exact for 1-3 byte runes (surrogates cut
  // out, :706-737), four-byte ranges widened to whole lead bytes (:796-842)
  static void wideRange(int32_t lo, int32_t hi, std::vector<Seq>& alts) {
    if (lo <= 0x7F) {
      alts.push_back({range(lo, hi > 0x7F ? 0x7F : hi)});
      lo = 0x80;
    }
    if (lo > hi) return;
    if (lo <= 0x7FF) {
      splitRange(lo, hi > 0x7FF ? 0x7FF : hi, 2, alts);
      lo = 0x800;
    }
    if (lo > hi) return;
    if (lo <= 0xFFFF) {
      int32_t a = lo, b = hi > 0xFFFF ? 0xFFFF : hi;
      if (a <= 0xD7FF && b >= 0xE000) {
        splitRange(a, 0xD7FF, 3, alts);
        splitRange(0xE000, b, 3, alts);
      } else if (!(a >= 0xD800 && b <= 0xDFFF)) {
        if (a >= 0xD800 && a <= 0xDFFF) a = 0xE000;
        if (b >= 0xD800 && b <= 0xDFFF) b = 0xD7FF;
        if (a <= b) splitRange(a, b, 3, alts);
      }
      lo = 0x10000;
    }
    if (lo > hi) return;
    if (hi > 0x10FFFF) hi = 0x10FFFF;
    int32_t a = (lo >> 18) << 18, b = ((hi >> 18) << 18) | 0x3FFFF;
    if (a < 0x10000) a = 0x10000;
    if (b > 0x10FFFF) b = 0x10FFFF;
    splitRange(a, b, 4, alts);
  }
  int unicodeClass(const std::vector<int32_t>& r, int next) {
    int64_t total = 0;
    bool large = false;
    for (size_t k = 0; k + 1 < r.size() && !large; k += 2) {
      total += (int64_t)r[k + 1] - r[k] + 1;
      large = total > 256;
    }
    std::vector<Seq> alts;
    if (!large) {
      // reference :463-486: one literal per code point, in class order
      for (size_t k = 0; k + 1 < r.size(); k += 2)
        for (int32_t c = r[k]; c <= r[k + 1]; c++) {
          uint8_t buf[4];
          const int n = encodeRune(buf, c);
          Seq q;
          for (int i = 0; i < n; i++) q.push_back(range(buf[i], buf[i]));
          alts.push_back(q);
        }
      return emitAlts(alts, next);
    }
    // reference :491-596: ASCII part | multi-byte part (| any stray high byte when the class
    // covers everything above 0x7F)
    ByteSet ascii{};
    bool has_ascii = false;
    std::vector<std::pair<int32_t, int32_t>> wide;
    for (size_t k = 0; k + 1 < r.size(); k += 2) {
      const int32_t lo = r[k], hi = r[k + 1];
      if (lo < 0x80) {
        set_add(ascii, (unsigned)lo, (unsigned)(hi < 0x80 ? hi : 0x7F));
        has_ascii = true;
      }
      if (hi >= 0x80) wide.push_back({lo < 0x80 ? 0x80 : lo, hi});
    }
    if (has_ascii) alts.push_back({ascii});
    if (wide.size() == 1 && wide[0].first <= 0x80 && wide[0].second >= 0x10FFFF) {
      validMultiByte(alts);
      alts.push_back({range(0x80, 0xFF)});
    } else {
      for (auto& w : wide) wideRange(w.first, w.second, alts);
      if (alts.size() > 8) return emitShared(alts, next);
    }
    return emitAlts(alts, next);
  }
  // reference :1142-1223
  int anyChar(bool include_nl, int next) {
    std::vector<Seq> alts;
    ByteSet ascii = range(0x00, 0x7F);
    if (!include_nl) ascii[0] &= ~(1ull << '\n');
    alts.push_back({ascii});
    validMultiByte(alts);
    ByteSet stray = range(0x80, 0xBF);
    set_add(stray, 0xC0, 0xC1);
    set_add(stray, 0xF5, 0xFF);
    alts.push_back({stray});
    return emitAlts(alts, next);
  }

  // loop helpers: `greedy` decides which split arm is preferred
  int star(const Regexp* sub, bool ng, int next) {
    if (nullable(sub)) {
      // (x+)? : plus split P loops, quest split Q enters or skips
      int P = emitSplit(-1, -1);
      int body = compile(sub, P);
      if (body < 0) return -1;
      p.inst[P].out = ng ? next : body;
      p.inst[P].out1 = ng ? body : next;
      return ng ? emitSplit(next, body) : emitSplit(body, next);
    }
    int L = emitSplit(-1, -1);
    int body = compile(sub, L);
    if (body < 0) return -1;
    p.inst[L].out = ng ? next : body;
    p.inst[L].out1 = ng ? body : next;
    return L;
  }
  int plus(const Regexp* sub, bool ng, int next) {
    int P = emitSplit(-1, -1);
    int body = compile(sub, P);
    if (body < 0) return -1;
    p.inst[P].out = ng ? next : body;
    p.inst[P].out1 = ng ? body : next;
    return body;
  }
  int quest(const Regexp* sub, bool ng, int next) {
    int body = compile(sub, next);
    if (body < 0) return -1;
    return ng ? emitSplit(next, body) : emitSplit(body, next);
  }
  int repeat(const Regexp* re, int next) {
    const Regexp* sub = re->sub[0];
    bool ng = re->flags & NonGreedy;
    int mn = re->min, mx = re->max;
    // sequence of pieces in forward order: mn copies, then star or (mx-mn) optional copies
    enum Kind { COPY, STAR, QUEST };
    std::vector<Kind> pieces;
    if (mx == -1) {
      if (mn == 0) return star(sub, ng, next);
      for (int i = 0; i < mn; i++) pieces.push_back(COPY);
      pieces.push_back(STAR);
    } else if (mn == mx) {
      if (mn == 0) return next;
      for (int i = 0; i < mn; i++) pieces.push_back(COPY);
    } else {
      for (int i = 0; i < mn; i++) pieces.push_back(COPY);
      for (int i = 0; i < mx - mn; i++) pieces.push_back(QUEST);
    }
    auto one = [&](Kind k, int nx) {
      switch (k) {
        case COPY: return compile(sub, nx);
        case STAR: return star(sub, ng, nx);
        default: return quest(sub, ng, nx);
      }
    };
    if (!reverse) {
      for (size_t i = pieces.size(); i-- > 0;) {
        next = one(pieces[i], next);
        if (next < 0) return -1;
      }
    } else {
      for (size_t i = 0; i < pieces.size(); i++) {
        next = one(pieces[i], next);
        if (next < 0) return -1;
      }
    }
    return next;
  }

  int compile(const Regexp* re, int next) {
    if (++depth > 200) {
      depth--;
      return fail("regex too complex");
    }
    int r = compile1(re, next);
    depth--;
    return r;
  }

  int compile1(const Regexp* re, int next) {
    bool ng = re->flags & NonGreedy;
    switch (re->op) {
      case OpLiteral: return literal(re, next);
      case OpCharClass: return charClass(re, next);
      case OpAnyChar: return anyChar(true, next);
      case OpAnyCharNotNL: return anyChar(false, next);
      case OpConcat: {
        if (!reverse) {
          for (size_t i = re->sub.size(); i-- > 0;) {
            next = compile(re->sub[i], next);
            if (next < 0) return -1;
          }
        } else {
          for (size_t i = 0; i < re->sub.size(); i++) {
            next = compile(re->sub[i], next);
            if (next < 0) return -1;
          }
        }
        return next;
      }
      case OpAlternate: {
        std::vector<int> entries;
        for (auto* s : re->sub) {
          int e = compile(s, next);
          if (e < 0) return -1;
          entries.push_back(e);
        }
        int acc = entries.back();
        for (size_t i = entries.size() - 1; i-- > 0;) acc = emitSplit(entries[i], acc);
        return acc;
      }
      case OpStar: return star(re->sub[0], ng, next);
      case OpPlus: return plus(re->sub[0], ng, next);
      case OpQuest: return quest(re->sub[0], ng, next);
      case OpRepeat: return repeat(re, next);
      case OpCapture: {
        if (re->sub.empty()) return next;
        if (reverse) return compile(re->sub[0], next);  // reverse programs carry no captures
        Inst close;
        close.op = I_SAVE;
        close.slot = 2 * re->cap + 1;
        close.out = next;
        int c = emit(close);
        int body = compile(re->sub[0], c);
        if (body < 0) return -1;
        Inst open;
        open.op = I_SAVE;
        open.slot = 2 * re->cap;
        open.out = body;
        return emit(open);
      }
      case OpBeginText: return look(reverse ? L_END_TEXT : L_START_TEXT, next);
      case OpEndText: return look(reverse ? L_START_TEXT : L_END_TEXT, next);
      case OpBeginLine: return look(reverse ? L_END_LINE : L_START_LINE, next);
      case OpEndLine: return look(reverse ? L_START_LINE : L_END_LINE, next);
      case OpWordBoundary: return look(L_WORD, next);
      case OpNoWordBoundary: return look(L_NOT_WORD, next);
      case OpEmptyMatch: return next;
      case OpNoMatch: {
        Inst i;
        i.op = I_FAIL;
        return emit(i);
      }
      default:
        return fail("unsupported regex operation");
    }
  }

  int look(LookKind k, int next) {
    p.has_looks = true;
    if (k == L_WORD || k == L_NOT_WORD) p.has_word_looks = true;
    Inst i;
    i.op = I_ASSERT;
    i.look = k;
    i.out = next;
    return emit(i);
  }
};

bool patternAnchored(const Regexp* re) {
  switch (re->op) {
    case OpBeginText: return true;
    case OpConcat:
    case OpCapture: return !re->sub.empty() && patternAnchored(re->sub[0]);
    default: return false;
  }
}

std::string build(const Regexp* re, Prog& out, bool reverse) {
  out = Prog();
  Builder b{out};
  b.reverse = reverse;
  Inst m;
  m.op = I_MATCH;
  int match = b.emit(m);
  int start = b.compile(re, match);
  if (start < 0) return b.err.empty() ? "compile failed" : b.err;
  out.start = start;
  out.num_captures = MaxCap(re) + 1;
  out.cap_names.assign(out.num_captures, "");
  std::function<void(const Regexp*)> names = [&](const Regexp* x) {
    if (x->op == OpCapture && x->cap > 0 && x->cap < out.num_captures) out.cap_names[x->cap] = x->name;
    for (const Regexp* sub : x->sub) names(sub);
  };
  names(re);
  out.anchored_start = patternAnchored(re);
  return "";
}

}  // namespace

std::string CompileProg(const Regexp* re, Prog& out) { return build(re, out, false); }
std::string CompileReverseProg(const Regexp* re, Prog& out) { return build(re, out, true); }

}  // namespace cgx
