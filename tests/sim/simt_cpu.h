// simt_cpu.h — TEST HARNESS ONLY.  A tiny single-threaded SIMT emulator that lets the kernel
// SOURCE (coregex_b200/csrc/scan_bits.cu) be compiled with g++ and stepped on the CPU, one
// ucontext fiber per CUDA thread, so that warp-level logic (shuffles, ballots, look-back
// protocol, mbarrier phases) can be debugged in this GPU-less container before GPU minutes are
// spent.  It is never linked into libcoregex_b200.so; the product has no CPU path.
//
// Model: all fibers of a launch run on one OS thread, round-robin.  A warp collective blocks its
// fiber until all 32 lanes of the warp arrived (double-buffered by parity of a per-lane collective
// counter).  Spin loops in the kernel must call cgx_spin_yield().  TMA bulk copies complete at
// issue time (the earliest legal moment), which exposes write-after-read hazards on the staging
// buffers as wrong data.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

namespace sim {

struct Dim3 {
  unsigned x = 0, y = 0, z = 0;
};

struct WarpState {
  uint64_t slot[2][32];
  int arrived[2] = {0, 0};
  int left[2] = {0, 0};
};

struct BlockState {
  unsigned arrived = 0, gen = 0;
};

struct Fiber {
  BlockState* block = nullptr;
  ucontext_t ctx;
  std::vector<unsigned char> stack;
  Dim3 tid, bid, bdim, gdim;
  int lane = 0;
  WarpState* warp = nullptr;
  unsigned char* smem = nullptr;
  unsigned ncoll = 0;  // collectives executed so far
  bool done = false;
};

struct Sched {
  std::vector<Fiber*> fibers;
  size_t cur = 0;
  size_t live = 0;
  ucontext_t main_ctx;
  unsigned long long switches = 0;
};

inline Sched*& sched() {
  static Sched* s = nullptr;
  return s;
}
inline Fiber*& cur() {
  static Fiber* f = nullptr;
  return f;
}

inline void yield() {
  Sched* s = sched();
  Fiber* me = cur();
  size_t n = s->fibers.size();
  size_t k = s->cur;
  for (size_t step = 1; step <= n; step++) {
    size_t j = (k + step) % n;
    if (!s->fibers[j]->done) {
      if (s->fibers[j] == me) return;
      s->cur = j;
      cur() = s->fibers[j];
      if (++s->switches > 4000000000ull) {
        fprintf(stderr, "simt_cpu: livelock suspected\n");
        abort();
      }
      swapcontext(&me->ctx, &s->fibers[j]->ctx);
      return;
    }
  }
}

// every lane contributes v; returns a pointer to the 32 contributed values (valid until the next
// collective of this warp with the same parity, i.e. two collectives later)
inline const uint64_t* exchange(uint64_t v) {
  Fiber* f = cur();
  WarpState* w = f->warp;
  const int b = f->ncoll & 1;
  f->ncoll++;
  w->slot[b][f->lane] = v;
  w->arrived[b]++;
  while (w->arrived[b] < 32) yield();
  if (++w->left[b] == 32) {
    // last lane out: everybody has observed arrived==32 (each lane increments `left` only after
    // its own wait loop ended), so the counters of this parity can be re-armed
    w->left[b] = 0;
    w->arrived[b] = 0;
    // NOTE: slot contents stay readable; the next writer of this parity comes two collectives later
  } else {
    // lanes that leave early must not race ahead two collectives; the parity scheme guarantees it
  }
  return w->slot[b];
}

}  // namespace sim

#define threadIdx (sim::cur()->tid)
#define blockIdx (sim::cur()->bid)
#define blockDim (sim::cur()->bdim)
#define gridDim (sim::cur()->gdim)

// ---- warp collectives -----------------------------------------------------------------------------
template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 8, "");
  uint64_t x = 0;
  memcpy(&x, &v, sizeof(T));
  const int me = sim::cur()->lane;
  const uint64_t* s = sim::exchange(x);
  uint64_t r = s[(src & 31)];
  (void)me;
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}
template <class T>
inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
  const int me = sim::cur()->lane;
  const int src = me - (int)d;
  T r = __shfl_sync(m, v, src < 0 ? me : src);
  return r;
}
template <class T>
inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
  const int me = sim::cur()->lane;
  const int src = me + (int)d;
  return __shfl_sync(m, v, src > 31 ? me : src);
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int x) {
  return __shfl_sync(m, v, sim::cur()->lane ^ x);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  const uint64_t* s = sim::exchange(pred ? 1u : 0u);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(s[i] & 1u) << i;
  return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline void __syncwarp(unsigned = 0xffffffffu) { sim::exchange(0); }
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  const uint64_t* s = sim::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r += (unsigned)s[i];
  return r;
}

inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  const uint64_t* s = sim::exchange(v);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r = (unsigned)s[i] > r ? (unsigned)s[i] : r;
  return r;
}

inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  const uint64_t* s = sim::exchange(v);
  unsigned r = 0xffffffffu;
  for (int i = 0; i < 32; i++) r = (unsigned)s[i] < r ? (unsigned)s[i] : r;
  return r;
}

// ---- scalar intrinsics -----------------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline unsigned __brev(unsigned x) {
  x = (x >> 16) | (x << 16);
  x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
  x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
  x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
  return x;
}
inline unsigned long long __brevll(unsigned long long x) {
  return ((unsigned long long)__brev((unsigned)x) << 32) | __brev((unsigned)(x >> 32));
}
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
  return c;
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
  s &= 31;
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
  s &= 31;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
  const unsigned long long v = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 255u) << (8 * i);
  return r;
}
template <class T>
inline T __ldg(const T* p) {
  return *p;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) {
  unsigned o = *p;
  *p = o + v;
  return o;
}
inline unsigned atomicOr(unsigned* p, unsigned v) {
  unsigned o = *p;
  *p = o | v;
  return o;
}
inline int atomicCAS(int* p, int cmp, int val) {
  const int o = *p;
  if (o == cmp) *p = val;
  return o;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p;
  *p = o + v;
  return o;
}

struct uint2 {
  unsigned x, y;
};
struct uint4 {
  unsigned x, y, z, w;
};
struct longlong2 {
  long long x, y;
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }

// ---- the helpers scan_common.cuh implements with inline PTX on the device --------------------------
namespace cgx {
inline void cgx_spin_yield() { sim::yield(); }
inline void cgx_threadfence() {}
inline void cgx_fence_block() {}
inline void cgx_backoff() { sim::yield(); }
inline void cgx_idle() { sim::yield(); }
inline void cgx_syncthreads() {
  sim::Fiber* f = sim::cur();
  sim::BlockState* b = f->block;
  const unsigned gen = b->gen;
  if (++b->arrived == f->bdim.x) {
    b->arrived = 0;
    b->gen++;
  } else {
    while (b->gen == gen) sim::yield();
  }
}
// mbarrier model: the word counts completed phases
inline void mbar_init(uint64_t* bar, unsigned) { *bar = 0; }
inline void fence_mbar_init() {}
inline void fence_proxy_async() {}
inline void mbar_expect_tx(uint64_t*, uint32_t) {}
inline void mbar_arrive(uint64_t* bar) { (*bar)++; }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (((*bar) & 1u) == parity) sim::yield();
}
inline void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (bytes & 15) || !bytes) {
    fprintf(stderr, "simt_cpu: illegal bulk copy dst=%p src=%p bytes=%u\n", dst, src, bytes);
    abort();
  }
  memcpy(dst, src, bytes);
  (*bar)++;  // expect_tx + complete_tx of the whole transfer: phase done
}
inline void tma_prefetch_l2(const void*, uint32_t) {}
inline unsigned long long ld_status(const unsigned long long* p) { return *(volatile const unsigned long long*)p; }
inline void st_status(unsigned long long* p, unsigned long long v) { *(volatile unsigned long long*)p = v; }
}  // namespace cgx

// ---- launcher -----------------------------------------------------------------------------------------
namespace sim {

template <class Args>
struct Tramp {
  void (*kernel)(Args);
  const Args* args;
};

template <class Args>
void trampoline(unsigned lo, unsigned hi) {
  Tramp<Args>* t = reinterpret_cast<Tramp<Args>*>(((uintptr_t)hi << 32) | lo);
  t->kernel(*t->args);
  Fiber* me = cur();
  me->done = true;
  Sched* s = sched();
  s->live--;
  if (s->live == 0) {
    swapcontext(&me->ctx, &s->main_ctx);
  } else {
    yield();
  }
  fprintf(stderr, "simt_cpu: finished fiber resumed\n");
  abort();
}

// runs kernel(args) over grid x block threads (block a multiple of 32) with `smem` bytes of
// dynamic shared memory per block
template <class Args>
void launch(void (*kernel)(Args), unsigned grid, unsigned block, size_t smem, const Args& args) {
  Sched s;
  sched() = &s;
  Tramp<Args> tr{kernel, &args};
  std::vector<std::vector<unsigned char>> smems(grid);
  std::vector<WarpState> warps((size_t)grid * (block / 32));
  std::vector<BlockState> blocks(grid);
  std::vector<Fiber> fibers((size_t)grid * block);
  for (unsigned b = 0; b < grid; b++) {
    smems[b].assign(smem + 256, 0xCD);
    unsigned char* base = smems[b].data();
    base += (128 - ((uintptr_t)base & 127)) & 127;
    for (unsigned t = 0; t < block; t++) {
      Fiber& f = fibers[(size_t)b * block + t];
      f.tid.x = t;
      f.bid.x = b;
      f.bdim.x = block;
      f.gdim.x = grid;
      f.lane = t & 31;
      f.warp = &warps[(size_t)b * (block / 32) + t / 32];
      f.block = &blocks[b];
      f.smem = base;
      f.stack.resize(256 * 1024);
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data();
      f.ctx.uc_stack.ss_size = f.stack.size();
      f.ctx.uc_link = nullptr;
      const uintptr_t p = (uintptr_t)&tr;
      makecontext(&f.ctx, (void (*)())trampoline<Args>, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
      s.fibers.push_back(&f);
    }
  }
  s.live = s.fibers.size();
  s.cur = 0;
  cur() = s.fibers[0];
  swapcontext(&s.main_ctx, &s.fibers[0]->ctx);
  sched() = nullptr;
  cur() = nullptr;
}

}  // namespace sim

#define CGX_DYN_SMEM(name) unsigned char* name = sim::cur()->smem
