// oracle/pikevm.cpp — TEST INFRASTRUCTURE ONLY.  See pikevm.h for the reference citations.
#include "pikevm.h"

#include <algorithm>

namespace oracle {

static inline bool isWordByte(uint8_t b) {
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}

bool checkLookAssertion(Look look, const uint8_t* h, int64_t n, int64_t pos) {
  switch (look) {
    case LookStartText: return pos == 0;
    case LookEndText: return pos == n;
    case LookStartLine: return pos == 0 || (pos > 0 && h[pos - 1] == '\n');
    case LookEndLine: return pos == n || (pos < n && h[pos] == '\n');
    case LookWordBoundary: {
      bool wb = pos > 0 && isWordByte(h[pos - 1]);
      bool wa = pos < n && isWordByte(h[pos]);
      return wb != wa;
    }
    case LookNoWordBoundary: {
      bool wb = pos > 0 && isWordByte(h[pos - 1]);
      bool wa = pos < n && isWordByte(h[pos]);
      return wb == wa;
    }
  }
  return false;
}

PikeVM::PikeVM(const NFA* nfa) : nfa_(nfa) {
  nslots_ = nfa->capture_count * 2;
  visited_.assign(nfa->states.size(), 0);
  cur_slots_tab_.assign((nfa->states.size() + 1) * nslots_, -1);
  next_slots_tab_.assign((nfa->states.size() + 1) * nslots_, -1);
  curr_slots_.assign(nslots_, -1);
}

// reference nfa/pikevm.go:1895-2005 / :2066-2180 — the two functions differ only in which
// queue and which slot table they write, so they share this body.
void PikeVM::closure(StateID state, int64_t start, const uint8_t* h, int64_t n, int64_t pos,
                     std::vector<Thread>& queue, std::vector<int64_t>* tab) {
  struct Frame {
    StateID state;  // InvalidState => restore frame
    int64_t start;
    int slot;
    int64_t value;
  };
  std::vector<Frame> stack;
  stack.push_back({state, start, 0, 0});
  while (!stack.empty()) {
    Frame f = stack.back();
    stack.pop_back();
    if (f.state == InvalidState) {
      if (tab && f.slot < nslots_) curr_slots_[f.slot] = f.value;
      continue;
    }
    StateID sid = f.state;
    if (sid >= nfa_->states.size()) continue;
    if (visited_[sid]) continue;
    visited_[sid] = 1;
    const State& st = nfa_->states[sid];
    switch (st.kind) {
      case StateMatch:
      case StateByteRange:
      case StateSparse:
        if (tab) std::copy(curr_slots_.begin(), curr_slots_.end(), tab->begin() + (size_t)sid * nslots_);
        queue.push_back({sid, f.start});
        break;
      case StateEpsilon:
        if (st.next != InvalidState) stack.push_back({st.next, f.start, 0, 0});
        break;
      case StateSplit:
        if (st.right != InvalidState) stack.push_back({st.right, f.start, 0, 0});
        if (st.left != InvalidState) stack.push_back({st.left, f.start, 0, 0});
        break;
      case StateCapture:
        if (st.next != InvalidState) {
          if (tab) {
            int slot = (int)st.cap_index * 2 + (st.cap_start ? 0 : 1);
            if (slot < nslots_) {
              stack.push_back({InvalidState, 0, slot, curr_slots_[slot]});
              curr_slots_[slot] = pos;
            }
          }
          stack.push_back({st.next, f.start, 0, 0});
        }
        break;
      case StateLook:
        if (checkLookAssertion(st.look, h, n, pos) && st.next != InvalidState)
          stack.push_back({st.next, f.start, 0, 0});
        break;
      default:
        break;
    }
  }
}

void PikeVM::step(const Thread& t, uint8_t b, const uint8_t* h, int64_t n, int64_t next_pos,
                  bool caps) {
  const State& st = nfa_->states[t.state];
  auto go = [&](StateID next) {
    if (caps)
      std::copy(cur_slots_tab_.begin() + (size_t)t.state * nslots_,
                cur_slots_tab_.begin() + (size_t)(t.state + 1) * nslots_, curr_slots_.begin());
    closure(next, t.start, h, n, next_pos, next_, caps ? &next_slots_tab_ : nullptr);
  };
  if (st.kind == StateByteRange) {
    if (b >= st.lo && b <= st.hi) go(st.next);
  } else if (st.kind == StateSparse) {
    for (auto& tr : st.trans)
      if (b >= tr.lo && b <= tr.hi) go(tr.next);
  }
}

// reference nfa/pikevm.go:1569-1630
bool PikeVM::matchesEmptyAt(const uint8_t* h, int64_t n, int64_t pos) {
  clearVisited();
  std::vector<StateID> stack{nfa_->start_anchored};
  visited_[nfa_->start_anchored] = 1;
  while (!stack.empty()) {
    StateID id = stack.back();
    stack.pop_back();
    if (nfa_->is_match(id)) return true;
    if (id >= nfa_->states.size()) continue;
    const State& st = nfa_->states[id];
    auto push = [&](StateID x) {
      if (x != InvalidState && x < visited_.size() && !visited_[x]) {
        visited_[x] = 1;
        stack.push_back(x);
      }
    };
    switch (st.kind) {
      case StateEpsilon: push(st.next); break;
      case StateSplit:
        push(st.left);
        push(st.right);
        break;
      case StateLook:
        if (checkLookAssertion(st.look, h, n, pos)) push(st.next);
        break;
      case StateCapture: push(st.next); break;
      default: break;
    }
  }
  return false;
}

static inline bool better(int64_t bs, int64_t be, int64_t cs, int64_t ce) {
  if (bs == -1) return true;
  if (cs < bs) return true;
  if (cs > bs) return false;
  return ce > be;
}

bool PikeVM::SearchAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e) {
  if (at > n) return false;
  if (at == n) {
    if (matchesEmptyAt(h, n, at)) {
      s = e = at;
      return true;
    }
    return false;
  }
  cur_.clear();
  next_.clear();
  clearVisited();
  if (nfa_->anchored) {
    // reference :1832-1890
    closure(nfa_->start_anchored, at, h, n, at, cur_, nullptr);
    int64_t last = -1;
    for (int64_t pos = at; pos <= n; pos++) {
      if (pos < n) {
        uint8_t b = h[pos];
        clearVisited();
        for (auto& t : cur_) {
          if (nfa_->is_match(t.state)) {
            if (pos > last || last == -1) last = pos;
            break;
          }
          step(t, b, h, n, pos + 1, false);
        }
      } else {
        for (auto& t : cur_)
          if (nfa_->is_match(t.state)) {
            if (pos > last || last == -1) last = pos;
            break;
          }
      }
      if (next_.empty() && (pos >= n || last != -1)) break;
      if (pos >= n) break;
      cur_.swap(next_);
      next_.clear();
    }
    if (last != -1) {
      s = at;
      e = last;
      return true;
    }
    return false;
  }
  // reference :1747-1829
  int64_t bs = -1, be = -1;
  for (int64_t pos = at; pos <= n; pos++) {
    if (bs == -1) {
      clearVisited();
      closure(nfa_->start_anchored, pos, h, n, pos, cur_, nullptr);
    }
    if (pos < n) {
      uint8_t b = h[pos];
      clearVisited();
      for (auto& t : cur_) {
        if (nfa_->is_match(t.state)) {
          if (better(bs, be, t.start, pos)) {
            bs = t.start;
            be = pos;
          }
          break;
        }
        step(t, b, h, n, pos + 1, false);
      }
    } else {
      for (auto& t : cur_)
        if (nfa_->is_match(t.state)) {
          if (better(bs, be, t.start, pos)) {
            bs = t.start;
            be = pos;
          }
          break;
        }
    }
    if (pos >= n) break;
    if (bs != -1) {
      bool has = false;
      for (auto& t : next_)
        if (t.start <= bs) {
          has = true;
          break;
        }
      if (!has) break;
    }
    cur_.swap(next_);
    next_.clear();
  }
  if (bs != -1) {
    s = bs;
    e = be;
    return true;
  }
  return false;
}

bool PikeVM::SearchCapturesAt(const uint8_t* h, int64_t n, int64_t at, std::vector<int64_t>& slots) {
  auto finish = [&](const std::vector<int64_t>* best, int64_t ms, int64_t me) {
    // reference :2411-2432 buildCapturesFromSlots
    slots.assign(nslots_, -1);
    slots[0] = ms;
    slots[1] = me;
    if (best)
      for (int i = 1; i < nfa_->capture_count; i++) {
        int64_t a = (*best)[2 * i], b = (*best)[2 * i + 1];
        if (a >= 0 && b >= 0) {
          slots[2 * i] = a;
          slots[2 * i + 1] = b;
        }
      }
    return true;
  };
  if (at > n) return false;
  if (at == n) {
    if (matchesEmptyAt(h, n, at)) return finish(nullptr, at, at);
    return false;
  }
  cur_.clear();
  next_.clear();
  clearVisited();
  std::fill(cur_slots_tab_.begin(), cur_slots_tab_.end(), -1);
  std::fill(next_slots_tab_.begin(), next_slots_tab_.end(), -1);
  std::vector<int64_t> best;
  bool have_best = false;
  auto take = [&](StateID st) {
    best.assign(cur_slots_tab_.begin() + (size_t)st * nslots_,
                cur_slots_tab_.begin() + (size_t)(st + 1) * nslots_);
    have_best = true;
  };
  if (nfa_->anchored) {
    std::fill(curr_slots_.begin(), curr_slots_.end(), -1);
    closure(nfa_->start_anchored, at, h, n, at, cur_, &cur_slots_tab_);
    int64_t last = -1;
    for (int64_t pos = at; pos <= n; pos++) {
      if (pos < n) {
        uint8_t b = h[pos];
        clearVisited();
        for (auto& t : cur_) {
          if (nfa_->is_match(t.state)) {
            if (pos > last || last == -1) {
              last = pos;
              take(t.state);
            }
            break;
          }
          step(t, b, h, n, pos + 1, true);
        }
      } else {
        for (auto& t : cur_)
          if (nfa_->is_match(t.state)) {
            if (pos > last || last == -1) {
              last = pos;
              take(t.state);
            }
            break;
          }
      }
      if (next_.empty() && (pos >= n || last != -1)) break;
      if (pos >= n) break;
      cur_.swap(next_);
      next_.clear();
      cur_slots_tab_.swap(next_slots_tab_);
    }
    if (last == -1) return false;
    return finish(have_best ? &best : nullptr, at, last);
  }
  int64_t bs = -1, be = -1;
  for (int64_t pos = at; pos <= n; pos++) {
    if (bs == -1) {
      // Visited deliberately NOT cleared here (reference :2249-2251)
      std::fill(curr_slots_.begin(), curr_slots_.end(), -1);
      closure(nfa_->start_anchored, pos, h, n, pos, cur_, &cur_slots_tab_);
    }
    if (pos < n) {
      uint8_t b = h[pos];
      clearVisited();
      for (auto& t : cur_) {
        if (nfa_->is_match(t.state)) {
          if (better(bs, be, t.start, pos)) {
            bs = t.start;
            be = pos;
            take(t.state);
          }
          break;
        }
        step(t, b, h, n, pos + 1, true);
      }
    } else {
      for (auto& t : cur_)
        if (nfa_->is_match(t.state)) {
          if (better(bs, be, t.start, pos)) {
            bs = t.start;
            be = pos;
            take(t.state);
          }
          break;
        }
    }
    if (pos >= n) break;
    if (bs != -1) {
      bool has = false;
      for (auto& t : next_)
        if (t.start <= bs) {
          has = true;
          break;
        }
      if (!has) break;
    }
    cur_.swap(next_);
    next_.clear();
    cur_slots_tab_.swap(next_slots_tab_);
  }
  if (bs == -1) return false;
  return finish(have_best ? &best : nullptr, bs, be);
}

}  // namespace oracle
