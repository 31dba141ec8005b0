// oracle/pikevm.h — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// CPU restatement of the reference's PikeVM slot-table search:
//   reference nfa/pikevm.go:1646-1675 (checkLookAssertion)
//   reference nfa/pikevm.go:1711-1745 (SearchWithSlotTableAt), :1747-1829 (unanchored),
//     :1832-1890 (anchored), :1895-2005 (addSearchThread: DFS closure, RestoreCapture frames),
//     :2009-2060 (stepSearchThread), :2066-2180 (addSearchThreadToNext)
//   reference nfa/pikevm.go:2186-2330 (captures, unanchored), :2333-2409 (captures, anchored),
//     :2411-2432 (buildCapturesFromSlots), :147-173 (isBetterMatch), :1569-1630 (matchesEmptyAt)
//   reference nfa/slot_table.go:13-80 (per-state slot rows, two generations)
#pragma once
#include <cstdint>
#include <vector>

#include "nfa.h"

namespace oracle {

class PikeVM {
 public:
  explicit PikeVM(const NFA* nfa);
  // first match at or after `at`; returns false if none
  bool SearchAt(const uint8_t* h, int64_t n, int64_t at, int64_t& s, int64_t& e);
  // first match with capture slots (2*capture_count entries, -1 = unset); slots[0..1]=match
  bool SearchCapturesAt(const uint8_t* h, int64_t n, int64_t at, std::vector<int64_t>& slots);
  bool matchesEmptyAt(const uint8_t* h, int64_t n, int64_t pos);

 private:
  struct Thread {
    StateID state;
    int64_t start;
  };
  const NFA* nfa_;
  int nslots_;
  std::vector<Thread> cur_, next_;
  std::vector<uint8_t> visited_;
  std::vector<int64_t> cur_slots_tab_, next_slots_tab_;  // (states+1) x nslots
  std::vector<int64_t> curr_slots_;                      // working buffer

  void clearVisited() { std::fill(visited_.begin(), visited_.end(), 0); }
  // closure of t into `queue`, writing slot rows into `tab` when captures are on
  void closure(StateID state, int64_t start, const uint8_t* h, int64_t n, int64_t pos,
               std::vector<Thread>& queue, std::vector<int64_t>* tab);
  void step(const Thread& t, uint8_t b, const uint8_t* h, int64_t n, int64_t next_pos, bool caps);
};

bool checkLookAssertion(Look look, const uint8_t* h, int64_t n, int64_t pos);

}  // namespace oracle
