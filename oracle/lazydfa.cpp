// oracle/lazydfa.cpp — TEST INFRASTRUCTURE ONLY.  See lazydfa.h for the reference citations.
#include "lazydfa.h"

#include <algorithm>

namespace oracle {

static inline bool isWordByte(uint8_t b) {
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}

static inline uint32_t lookBit(Look l) {
  switch (l) {
    case LookStartText: return LS_StartText;
    case LookEndText: return LS_EndText;
    case LookStartLine: return LS_StartLine;
    case LookEndLine: return LS_EndLine;
    case LookWordBoundary: return LS_WordBoundary;
    case LookNoWordBoundary: return LS_NoWordBoundary;
  }
  return 0;
}

// reference dfa/lazy/look.go:88-104
static uint32_t lookSetFromStartKind(StartKind k) {
  switch (k) {
    case StartText: return LS_StartText | LS_StartLine;
    case StartLineLF: return LS_StartLine;
    default: return 0;
  }
}

// reference dfa/lazy/start.go:96-110
static StartKind kindOfByte(uint8_t b) {
  if (b == '\n') return StartLineLF;
  if (b == '\r') return StartLineCR;
  if (isWordByte(b)) return StartWord;
  return StartNonWord;
}

LazyDFA::LazyDFA(const NFA* nfa, LazyConfig cfg) : nfa_(nfa), cfg_(cfg) {
  for (auto& s : nfa->states) {
    if (s.kind == StateLook) {
      if (s.look == LookWordBoundary || s.look == LookNoWordBoundary) has_wb_ = true;
      if (s.look == LookEndLine) has_endline_ = true;
    }
  }
  for (auto& row : start_)
    for (auto& x : row) x = -1;
  stride_ = nfa->alphabet_len;
}

int LazyDFA::newState(DState&& s) {
  int id = (int)states_.size();
  states_.push_back(std::move(s));
  flat_.resize((size_t)(id + 1) * stride_, -2);
  return id;
}

// reference dfa/lazy/builder.go:245-293: add-on-pop, push right then left
void LazyDFA::closureInto(std::vector<StateID>& set, std::vector<uint8_t>& in_set, StateID seed,
                          uint32_t look_have) const {
  std::vector<StateID> stack{seed};
  while (!stack.empty()) {
    StateID cur = stack.back();
    stack.pop_back();
    if (cur >= nfa_->states.size()) continue;  // State(id)==nil; also covers InvalidState
    if (in_set[cur]) continue;
    in_set[cur] = 1;
    set.push_back(cur);
    const State& st = nfa_->states[cur];
    switch (st.kind) {
      case StateEpsilon:
        if (st.next != InvalidState) stack.push_back(st.next);
        break;
      case StateSplit:
        if (st.right != InvalidState) stack.push_back(st.right);
        if (st.left != InvalidState) stack.push_back(st.left);
        break;
      case StateLook:
        if ((look_have & lookBit(st.look)) && st.next != InvalidState) stack.push_back(st.next);
        break;
      case StateCapture:
        if (st.next != InvalidState) stack.push_back(st.next);
        break;
      default:
        break;
    }
  }
}

std::vector<StateID> LazyDFA::closure(const std::vector<StateID>& seeds, uint32_t look_have) const {
  std::vector<StateID> set;
  std::vector<uint8_t> in(nfa_->states.size(), 0);
  for (StateID s : seeds) closureInto(set, in, s, look_have);
  return set;
}

// reference dfa/lazy/builder.go:323-431.  NOTE the final result comes from StateSet.ToSlice()
// (not insertion order): reference dfa/lazy/state.go ToSlice sorts ascending.
std::vector<StateID> LazyDFA::resolveWB(const std::vector<StateID>& states, bool sat) const {
  std::vector<uint8_t> crossed(nfa_->states.size(), 0);
  std::vector<StateID> stack;
  auto tryLook = [&](const State& st) {
    if (st.next == InvalidState) return;
    bool ok = (st.look == LookWordBoundary && sat) || (st.look == LookNoWordBoundary && !sat);
    if (ok && !crossed[st.next]) {
      crossed[st.next] = 1;
      stack.push_back(st.next);
    }
  };
  for (StateID sid : states) {
    if (sid >= nfa_->states.size()) continue;
    const State& st = nfa_->states[sid];
    if (st.kind == StateLook) tryLook(st);
  }
  if (stack.empty()) return states;
  while (!stack.empty()) {
    StateID cur = stack.back();
    stack.pop_back();
    if (cur >= nfa_->states.size()) continue;
    const State& st = nfa_->states[cur];
    auto add = [&](StateID n) {
      if (n != InvalidState && n < crossed.size() && !crossed[n]) {
        crossed[n] = 1;
        stack.push_back(n);
      }
    };
    switch (st.kind) {
      case StateLook: tryLook(st); break;
      case StateEpsilon: add(st.next); break;
      case StateSplit:
        add(st.left);
        add(st.right);
        break;
      case StateCapture: add(st.next); break;
      default: break;
    }
  }
  std::vector<uint8_t> in(nfa_->states.size(), 0);
  for (StateID s : states)
    if (s < in.size()) in[s] = 1;
  for (size_t i = 0; i < crossed.size(); i++)
    if (crossed[i]) in[i] = 1;
  std::vector<StateID> out;
  for (size_t i = 0; i < in.size(); i++)
    if (in[i]) out.push_back((StateID)i);
  return out;
}

bool LazyDFA::containsMatch(const std::vector<StateID>& s) const {
  for (StateID x : s)
    if (nfa_->is_match(x)) return true;
  return false;
}

// reference dfa/lazy/builder.go:183-243
std::vector<StateID> LazyDFA::move(const std::vector<StateID>& states, uint8_t b, bool from_word,
                                   bool break_at_match) const {
  std::vector<StateID> resolved_store;
  const std::vector<StateID>* resolved = &states;
  if (has_wb_) {
    bool sat = from_word != isWordByte(b);
    resolved_store = resolveWB(states, sat);
    resolved = &resolved_store;
  }
  uint32_t look_after = b == '\n' ? LS_StartLine : 0;
  std::vector<StateID> result;
  std::vector<uint8_t> in(nfa_->states.size(), 0);
  for (StateID sid : *resolved) {
    if (sid >= nfa_->states.size()) continue;
    const State& st = nfa_->states[sid];
    if (break_at_match && st.kind == StateMatch) break;
    if (st.kind == StateByteRange) {
      if (b >= st.lo && b <= st.hi) closureInto(result, in, st.next, look_after);
    } else if (st.kind == StateSparse) {
      for (auto& t : st.trans)
        if (b >= t.lo && b <= t.hi) closureInto(result, in, t.next, look_after);
    }
  }
  return result;
}

LazyDFA::Key LazyDFA::makeKey(const std::vector<StateID>& ids, bool from_word, bool is_match) {
  std::vector<StateID> s = ids;
  std::sort(s.begin(), s.end());
  return {std::move(s), (from_word ? 1 : 0) | (is_match ? 2 : 0)};
}

// reference dfa/lazy/lazy.go:1336-1446
int LazyDFA::determinize(int cur, uint8_t b) {
  uint8_t cls = nfa_->byte_classes[b];
  std::vector<StateID> cur_nfa = states_[cur].nfa;
  if (has_endline_ && b == '\n') cur_nfa = closure(cur_nfa, LS_EndLine);
  bool source_has_match = containsMatch(cur_nfa);
  bool bam = source_has_match && cfg_.break_at_match;
  std::vector<StateID> next = move(cur_nfa, b, states_[cur].from_word, bam);
  bool is_match = source_has_match;
  if (next.empty() && !is_match) {
    flat_[(size_t)cur * stride_ + cls] = -1;
    return -1;
  }
  bool next_from_word = isWordByte(b);
  Key key = makeKey(next, next_from_word, is_match);
  auto it = cache_.find(key);
  if (it != cache_.end()) {
    flat_[(size_t)cur * stride_ + cls] = it->second;
    return it->second;
  }
  DState ns;
  ns.nfa = next;
  ns.is_match = is_match;
  ns.from_word = next_from_word;
  if (has_wb_ && !is_match) {
    ns.match_at_wb = containsMatch(resolveWB(next, true));
    ns.match_at_nwb = containsMatch(resolveWB(next, false));
  }
  int id = newState(std::move(ns));
  cache_[key] = id;
  flat_[(size_t)cur * stride_ + cls] = id;
  return id;
}

int LazyDFA::step(int sid, uint8_t b) {
  int32_t t = flat_[(size_t)sid * stride_ + nfa_->byte_classes[b]];
  if (t == -2) return determinize(sid, b);
  return t;
}

// reference dfa/lazy/lazy.go:1569-1613 + dfa/lazy/start.go:205-256
int LazyDFA::getStart(const uint8_t* h, int64_t pos, bool anchored) {
  StartKind kind = pos == 0 ? StartText : kindOfByte(h[pos - 1]);
  int& slot = start_[anchored ? 1 : 0][kind];
  if (slot >= 0) return slot;
  StateID s0 = anchored ? nfa_->start_anchored : nfa_->start_unanchored;
  std::vector<StateID> set = closure({s0}, lookSetFromStartKind(kind));
  bool from_word = kind == StartWord;
  Key key = makeKey(set, from_word, false);
  auto it = cache_.find(key);
  if (it != cache_.end()) {
    slot = it->second;
    return slot;
  }
  DState ns;
  ns.nfa = set;
  ns.from_word = from_word;
  // NOTE: start states do not get matchAtWordBoundary flags in the reference
  // (ComputeStartStateWithStride does not set them) — keep false.
  int id = newState(std::move(ns));
  cache_[key] = id;
  slot = id;
  return id;
}

// reference dfa/lazy/builder.go:454-472
bool LazyDFA::checkEOI(int sid) const {
  const DState& st = states_[sid];
  std::vector<StateID> resolved = resolveWB(st.nfa, st.from_word);
  std::vector<StateID> fin = closure(resolved, LS_EndText | LS_EndLine);
  return containsMatch(fin);
}

// reference dfa/lazy/lazy.go:1636-1647: fresh caches have no StartState entry, so this is
// "does the NFA match the empty input" (PikeVM.Search([]byte{})).
bool LazyDFA::matchesEmpty() {
  std::vector<StateID> set = closure({nfa_->start_unanchored}, LS_StartText | LS_StartLine);
  // at EOI of empty input previous byte is "non-word": \B satisfied, \b not.
  std::vector<StateID> resolved = resolveWB(set, false);
  std::vector<StateID> fin = closure(resolved, LS_StartText | LS_StartLine | LS_EndText | LS_EndLine);
  return containsMatch(fin);
}

bool LazyDFA::wbFast(const DState& st, uint8_t b) const {
  if (st.is_match) return false;
  bool boundary = st.from_word != isWordByte(b);
  return boundary ? st.match_at_wb : st.match_at_nwb;
}

int64_t LazyDFA::SearchAtAnchored(const uint8_t* h, int64_t n, int64_t at) {
  if (at > n) return -1;
  if (at == n) return matchesEmpty() ? at : -1;
  int sid = getStart(h, at, true);
  int64_t last = -1;
  if (!has_wb_) {
    // hot loop (reference lazy.go:251-313 without the word-boundary branch)
    const uint8_t* cls = nfa_->byte_classes;
    for (int64_t pos = at; pos < n; pos++) {
      int32_t nx = flat_[(size_t)sid * stride_ + cls[h[pos]]];
      if (nx == -2) nx = determinize(sid, h[pos]);
      if (nx < 0) return last;
      sid = nx;
      if (states_[sid].is_match) last = pos;
    }
    if (checkEOI(sid)) return n;
    return last;
  }
  for (int64_t pos = at; pos < n; pos++) {
    uint8_t b = h[pos];
    if (has_wb_ && wbFast(states_[sid], b)) return pos;
    int nx = step(sid, b);
    if (nx < 0) return last;
    sid = nx;
    if (states_[sid].is_match) last = pos;
  }
  if (checkEOI(sid)) return n;
  return last;
}

int64_t LazyDFA::SearchAt(const uint8_t* h, int64_t n, int64_t at) {
  if (at > n) return -1;
  if (at == n) return matchesEmpty() ? at : -1;
  if (nfa_->anchored && at > 0) return -1;
  int sid = getStart(h, at, false);
  int64_t last = -1;
  for (int64_t pos = at; pos < n; pos++) {
    uint8_t b = h[pos];
    // reference lazy.go:1258: checkWordBoundaryMatch (state not already match)
    if (has_wb_ && !states_[sid].is_match) {
      bool sat = states_[sid].from_word != isWordByte(b);
      if (containsMatch(resolveWB(states_[sid].nfa, sat))) return pos;
    }
    int nx = step(sid, b);
    if (nx < 0) return last;
    sid = nx;
    if (states_[sid].is_match) last = pos;
  }
  if (checkEOI(sid)) return n;
  return last;
}

int64_t LazyDFA::SearchReverse(const uint8_t* h, int64_t n, int64_t start, int64_t end) {
  if (end <= start || end > n) return -1;
  // reference lazy.go:2123-2158: start kind from the byte AFTER the region
  StartKind kind = end >= n ? StartText : kindOfByte(h[end]);
  int& slot = start_[0][kind];
  int sid;
  if (slot >= 0) {
    sid = slot;
  } else {
    std::vector<StateID> set = closure({nfa_->start_unanchored}, lookSetFromStartKind(kind));
    bool from_word = kind == StartWord;
    Key key = makeKey(set, from_word, false);
    auto it = cache_.find(key);
    if (it != cache_.end()) {
      sid = it->second;
    } else {
      DState ns;
      ns.nfa = set;
      ns.from_word = from_word;
      sid = newState(std::move(ns));
      cache_[key] = sid;
    }
    slot = sid;
  }
  int64_t last = -1;
  for (int64_t at = end - 1; at >= start; at--) {
    int nx = step(sid, h[at]);
    if (nx < 0) return last;
    sid = nx;
    if (states_[sid].is_match) last = at + 1;
  }
  if (containsMatch(states_[sid].nfa)) last = start;
  return last;
}

bool LazyDFA::IsMatch(const uint8_t* h, int64_t n) { return SearchAt(h, n, 0) >= 0; }

}  // namespace oracle
