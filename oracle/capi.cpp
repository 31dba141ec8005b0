// oracle/capi.cpp — TEST INFRASTRUCTURE ONLY (parity oracle; never linked into the product).
//
// C entry points so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs can drive the CPU restatement through ctypes.  Nothing in coregex_b200/ may
// load this library.
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "literal.h"
#include "meta.h"
#include "simd.h"
#include "teddy.h"

using namespace oracle;

extern "C" {

void* orc_compile(const char* pat, size_t len, char* err, size_t errcap) {
  std::string e;
  auto eng = Engine::Compile(std::string(pat, len), e);
  if (!eng) {
    if (err && errcap) {
      strncpy(err, e.c_str(), errcap - 1);
      err[errcap - 1] = 0;
    }
    return nullptr;
  }
  return eng.release();
}

void orc_free(void* p) { delete (Engine*)p; }
int orc_strategy(void* p) { return ((Engine*)p)->strategy(); }
const char* orc_strategy_name(void* p) { return StrategyName(((Engine*)p)->strategy()); }
int orc_strategy_exact(void* p) { return ((Engine*)p)->strategy_exact() ? 1 : 0; }
int orc_has_bidirectional(void* p) { return ((Engine*)p)->has_bidirectional() ? 1 : 0; }
void orc_set_bidirectional(void* p, int on) { ((Engine*)p)->set_bidirectional(on != 0); }
int orc_num_captures(void* p) { return ((Engine*)p)->num_captures(); }
int orc_digit_run_skip_safe(void* p) { return ((Engine*)p)->digit_run_skip_safe() ? 1 : 0; }

int orc_is_match(void* p, const uint8_t* h, int64_t n) { return ((Engine*)p)->IsMatch(h, n) ? 1 : 0; }

// returns the total number of matches; writes at most cap pairs
int64_t orc_find_all(void* p, const uint8_t* h, int64_t n, int64_t limit, int64_t* out, int64_t cap) {
  std::vector<int64_t> v;
  int64_t c = ((Engine*)p)->FindAll(h, n, limit, v);
  int64_t w = c < cap ? c : cap;
  if (out && w > 0) memcpy(out, v.data(), (size_t)w * 2 * sizeof(int64_t));
  return c;
}

int64_t orc_count(void* p, const uint8_t* h, int64_t n, int64_t limit) {
  return ((Engine*)p)->Count(h, n, limit);
}

// returns number of matches; out receives count*stride ints (stride = 2*num_captures)
int64_t orc_find_all_submatch(void* p, const uint8_t* h, int64_t n, int64_t limit, int64_t* out,
                              int64_t cap_matches) {
  std::vector<int64_t> v;
  Engine* e = (Engine*)p;
  int64_t c = e->FindAllSubmatch(h, n, limit, v);
  int64_t stride = 2 * e->num_captures();
  int64_t w = c < cap_matches ? c : cap_matches;
  if (out && w > 0) memcpy(out, v.data(), (size_t)(w * stride) * sizeof(int64_t));
  return c;
}

int orc_find_at(void* p, const uint8_t* h, int64_t n, int64_t at, int64_t* s, int64_t* e) {
  return ((Engine*)p)->FindIndicesAt(h, n, at, *s, *e) ? 1 : 0;
}

// AST / NFA dumps for golden tests.  Returned buffers are thread-local statics.
const char* orc_dump_ast(const char* pat, size_t len) {
  static thread_local std::string s;
  gosyntax::Arena a;
  auto r = gosyntax::Parse(std::string(pat, len), gosyntax::Perl, a);
  s = r.re ? gosyntax::Dump(r.re) : ("ERR " + r.err);
  return s.c_str();
}
const char* orc_dump_nfa(void* p) {
  static thread_local std::string s;
  s = DumpNFA(((Engine*)p)->nfa());
  return s.c_str();
}
// prefix literals joined by '\n', each line "<C|I> <hex bytes>"
const char* orc_dump_prefixes(void* p) {
  static thread_local std::string s;
  s.clear();
  Seq q = ExtractPrefixes(((Engine*)p)->ast());
  char buf[4];
  for (auto& l : q.lits) {
    s += l.complete ? "C " : "I ";
    for (unsigned char c : l.bytes) {
      snprintf(buf, sizeof buf, "%02x", c);
      s += buf;
    }
    s += '\n';
  }
  return s.c_str();
}

int64_t orc_memchr_digit_at(const uint8_t* h, int64_t n, int64_t at) { return memchr_digit_at(h, n, at); }
int64_t orc_memchr(const uint8_t* h, int64_t n, uint8_t c) { return memchr1(h, n, c); }
int64_t orc_memchr2(const uint8_t* h, int64_t n, uint8_t a, uint8_t b) { return memchr2(h, n, a, b); }
int64_t orc_memchr3(const uint8_t* h, int64_t n, uint8_t a, uint8_t b, uint8_t c) { return memchr3(h, n, a, b, c); }

// Standalone Teddy (pattern list = '\n'-free byte strings given as offsets) for kernel-level tests.
// Returns (start,end) of first match at or after `at`, or 0.
int orc_teddy_find(const uint8_t* pats, const int32_t* offs, int npat, const uint8_t* h, int64_t n,
                   int64_t at, int64_t* s, int64_t* e) {
  std::vector<std::string> v;
  for (int i = 0; i < npat; i++) v.emplace_back((const char*)pats + offs[i], offs[i + 1] - offs[i]);
  if (npat <= 32) {
    Teddy t(v);
    if (!t.ok()) return -1;
    return t.FindMatch(h, n, at, *s, *e) ? 1 : 0;
  }
  FatTeddy t(v);
  if (!t.ok()) return -1;
  return t.FindMatch(h, n, at, *s, *e) ? 1 : 0;
}

// ---- multi-threaded CPU baseline --------------------------------------------------------------
// The reference is a single-haystack engine; its documented concurrency model is one immutable
// Regex shared by goroutines, each with its own SearchState (reference meta/engine.go:124-137,
// README.md:171-190).  The stand-in: split the buffer at '\n' boundaries into `threads` shards,
// each thread compiles its own Engine (== its own SearchState) and runs FindAll (mode 0) or Count
// (mode 1) on its shard.  Valid only for patterns that cannot match across '\n' (checked by the
// callers for the config patterns).  Returns total matches; *seconds = wall time of the scan only.
int64_t orc_scan_mt(const char* pat, size_t patlen, const uint8_t* h, int64_t n, int threads, int mode,
                    double* seconds) {
  if (threads < 1) threads = 1;
  std::vector<int64_t> bounds(threads + 1, 0);
  bounds[threads] = n;
  for (int t = 1; t < threads; t++) {
    int64_t b = n / threads * t;
    if (b < bounds[t - 1]) b = bounds[t - 1];
    while (b < n && b > 0 && h[b - 1] != '\n') b++;
    bounds[t] = b;
  }
  std::vector<std::unique_ptr<Engine>> engs(threads);
  for (int t = 0; t < threads; t++) {
    std::string e;
    engs[t] = Engine::Compile(std::string(pat, patlen), e);
    if (!engs[t]) return -1;
  }
  // warm each engine's lazy DFA on a small prefix so compile/determinize cost is not timed
  for (int t = 0; t < threads; t++) {
    std::vector<int64_t> tmp;
    int64_t w = n < 65536 ? n : 65536;
    engs[t]->FindAll(h, w, -1, tmp);
  }
  std::vector<int64_t> counts(threads, 0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) {
    th.emplace_back([&, t]() {
      const uint8_t* base = h + bounds[t];
      int64_t len = bounds[t + 1] - bounds[t];
      if (mode == 1) {
        counts[t] = engs[t]->Count(base, len, -1);
      } else {
        std::vector<int64_t> v;
        v.reserve((size_t)(len / 50));
        counts[t] = engs[t]->FindAll(base, len, -1, v);
      }
    });
  }
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  int64_t total = 0;
  for (auto c : counts) total += c;
  return total;
}

}  // extern "C"
