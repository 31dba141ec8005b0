// dfa.cpp — eager leftmost-first determinization (see dfa.h for the table contract).
#include "dfa.h"

#include <algorithm>
#include <map>

namespace cgx {

namespace {

inline bool isWord(unsigned b) {
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}

constexpr int RESTART = -2;  // pseudo-thread of the unanchored prefix (lowest priority)

// context carried by a DFA state: what is known about the byte BEFORE the current position
struct Ctx {
  bool at_text = false;  // position 0
  bool prev_lf = false;
  bool prev_word = false;
  int pack() const { return (at_text ? 1 : 0) | (prev_lf ? 2 : 0) | (prev_word ? 4 : 0); }
};

struct DState {
  std::vector<int> list;  // ordered: I_SET / I_MATCH / pending look-ahead I_ASSERT pcs / RESTART
  Ctx ctx;
};

struct Det {
  const Prog& p;
  bool unanchored;
  std::vector<DState> states;  // index 0 = DEAD
  std::map<std::pair<std::vector<int>, int>, int> index;
  std::vector<uint8_t> seen;

  explicit Det(const Prog& prog, bool un) : p(prog), unanchored(un) { seen.resize(prog.inst.size()); }

  // DFS closure from pc with "look-behind" knowledge ctx.  Look-ahead asserts stay pending
  // unless `ahead` is non-null, in which case they are resolved against it.
  struct Ahead {
    bool end_text, end_line, next_word;
  };
  void closure(int pc, const Ctx& ctx, const Ahead* ahead, std::vector<int>& out) {
    std::vector<int> stack{pc};
    while (!stack.empty()) {
      int x = stack.back();
      stack.pop_back();
      if (x < 0 || seen[x]) continue;
      seen[x] = 1;
      const Inst& in = p.inst[x];
      switch (in.op) {
        case I_SET:
        case I_MATCH:
          out.push_back(x);
          break;
        case I_SPLIT:
          stack.push_back(in.out1);
          stack.push_back(in.out);
          break;
        case I_SAVE:
        case I_NOP:
          stack.push_back(in.out);
          break;
        case I_ASSERT: {
          bool known = true, ok = false;
          switch (in.look) {
            case L_START_TEXT: ok = ctx.at_text; break;
            case L_START_LINE: ok = ctx.at_text || ctx.prev_lf; break;
            case L_END_TEXT:
              if (ahead) ok = ahead->end_text; else known = false;
              break;
            case L_END_LINE:
              if (ahead) ok = ahead->end_line; else known = false;
              break;
            case L_WORD:
              if (ahead) ok = ctx.prev_word != ahead->next_word; else known = false;
              break;
            case L_NOT_WORD:
              if (ahead) ok = ctx.prev_word == ahead->next_word; else known = false;
              break;
          }
          if (!known) {
            out.push_back(x);  // pending: resolved when the next byte (or EOI) is known
          } else if (ok) {
            stack.push_back(in.out);
          }
          break;
        }
        default:
          break;
      }
    }
  }

  // expand a state's list against the look-ahead; returns ordered SET/MATCH/RESTART entries
  std::vector<int> expand(const DState& s, const Ahead& ah) {
    std::fill(seen.begin(), seen.end(), 0);
    std::vector<int> out;
    for (int e : s.list) {
      if (e == RESTART) {
        out.push_back(RESTART);
        continue;
      }
      const Inst& in = p.inst[e];
      if (in.op == I_ASSERT) {
        // resolve this pending assert now (un-mark so closure processes it)
        closure(e, s.ctx, &ah, out);
      } else {
        if (!seen[e]) {
          seen[e] = 1;
          out.push_back(e);
        }
      }
    }
    return out;
  }

  int intern(DState&& s) {
    if (s.list.empty()) return 0;
    Ctx c = s.ctx;
    // context bits only matter when the program can observe them
    if (!p.has_looks) c = Ctx();
    else if (!p.has_word_looks) c.prev_word = false;
    s.ctx = c;
    auto key = std::make_pair(s.list, c.pack());
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    int id = (int)states.size();
    states.push_back(std::move(s));
    index[key] = id;
    return id;
  }

  int startState(StartKindIdx k) {
    Ctx c;
    c.at_text = k == SK_TEXT;
    c.prev_lf = k == SK_LF;
    c.prev_word = k == SK_WORD;
    std::fill(seen.begin(), seen.end(), 0);
    DState s;
    s.ctx = c;
    closure(p.start, c, nullptr, s.list);
    if (unanchored) s.list.push_back(RESTART);
    return intern(std::move(s));
  }
};

}  // namespace

std::string BuildDFA(const Prog& p, bool anchored, int max_states, DfaTables& out, bool longest) {
  Det d(p, !anchored);
  d.states.emplace_back();  // DEAD
  out = DfaTables();
  for (int k = 0; k < SK_COUNT; k++) out.start[k] = (uint16_t)d.startState((StartKindIdx)k);

  std::vector<uint16_t> trans;
  std::vector<uint8_t> eoi;
  for (size_t si = 0; si < d.states.size(); si++) {
    if ((int)d.states.size() > max_states)
      return "dfa too large: more than " + std::to_string(max_states) + " states";
    trans.resize((si + 1) * 256, 0);
    eoi.resize(si + 1, 0);
    if (si == 0) continue;
    {
      // EOI: every end look holds; the "next byte" is non-word
      DState cur = d.states[si];
      Det::Ahead ah{true, true, false};
      std::vector<int> ex = d.expand(cur, ah);
      for (int e : ex)
        if (e != RESTART && p.inst[e].op == I_MATCH) {
          eoi[si] = 1;
          break;
        }
    }
    for (unsigned b = 0; b < 256; b++) {
      DState cur = d.states[si];  // copy: d.states may reallocate below
      Det::Ahead ah{false, b == '\n', isWord(b)};
      std::vector<int> ex = d.expand(cur, ah);
      bool match_before = false;
      size_t cut = ex.size();
      for (size_t i = 0; i < ex.size(); i++)
        if (ex[i] != RESTART && p.inst[ex[i]].op == I_MATCH) {
          match_before = true;
          if (!longest) cut = i;
          break;
        }
      Ctx nc;
      nc.at_text = false;
      nc.prev_lf = b == '\n';
      nc.prev_word = isWord(b);
      DState nx;
      nx.ctx = nc;
      std::fill(d.seen.begin(), d.seen.end(), 0);
      for (size_t i = 0; i < cut; i++) {
        int e = ex[i];
        if (e == RESTART) {
          d.closure(p.start, nc, nullptr, nx.list);
          nx.list.push_back(RESTART);
          continue;
        }
        const Inst& in = p.inst[e];
        if (in.op == I_SET && set_has(p.sets[in.set], b)) d.closure(in.out, nc, nullptr, nx.list);
      }
      // a state holding only RESTART is still live (unanchored scan continues)
      int to = d.intern(std::move(nx));
      trans[si * 256 + b] = (uint16_t)to | (match_before ? DFA_MATCH_BIT : 0);
    }
  }
  int n = (int)d.states.size();
  if (n > DFA_STATE_MASK) return "dfa too large";

  // fold states that cannot reach a match into DEAD
  std::vector<uint8_t> live(n, 0);
  bool changed = true;
  for (int s = 1; s < n; s++) {
    if (eoi[s]) live[s] = 1;
    for (int b = 0; b < 256 && !live[s]; b++)
      if (trans[s * 256 + b] & DFA_MATCH_BIT) live[s] = 1;
  }
  while (changed) {
    changed = false;
    for (int s = 1; s < n; s++) {
      if (live[s]) continue;
      for (int b = 0; b < 256; b++) {
        int t = trans[s * 256 + b] & DFA_STATE_MASK;
        if (t && live[t]) {
          live[s] = 1;
          changed = true;
          break;
        }
      }
    }
  }
  std::vector<int> remap(n, 0);
  int m = 1;
  for (int s = 1; s < n; s++)
    if (live[s]) remap[s] = m++;
  out.nstates = m;
  out.trans.assign((size_t)m * 256, 0);
  out.eoi.assign(m, 0);
  for (int s = 1; s < n; s++) {
    if (!live[s]) continue;
    int ns = remap[s];
    out.eoi[ns] = eoi[s];
    for (int b = 0; b < 256; b++) {
      uint16_t e = trans[s * 256 + b];
      int t = remap[e & DFA_STATE_MASK];
      out.trans[(size_t)ns * 256 + b] = (uint16_t)t | (e & DFA_MATCH_BIT);
    }
  }
  for (int k = 0; k < SK_COUNT; k++) out.start[k] = (uint16_t)remap[out.start[k]];
  for (int k = 0; k < SK_COUNT; k++) {
    int s = out.start[k];
    if (!s) continue;
    if (out.eoi[s]) out.matches_empty = true;
    for (int b = 0; b < 256; b++) {
      uint16_t e = out.trans[(size_t)s * 256 + b];
      if (e & DFA_MATCH_BIT) out.matches_empty = true;
      if (e & DFA_STATE_MASK) set_add(out.first_bytes, b, b);
    }
  }
  return "";
}

bool DelimiterSafe(const DfaTables& t, uint8_t d) {
  for (int s = 1; s < t.nstates; s++)
    if (t.trans[(size_t)s * 256 + d] & DFA_STATE_MASK) return false;
  return true;
}

}  // namespace cgx
