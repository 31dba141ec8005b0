// scan_common.cuh — device helpers for the sm_100a scan kernels: TMA 1-D bulk copy + mbarrier,
// SWAR byte-class tests, warp scans, decoupled look-back words.
#pragma once
#include "scan_params.h"  // fixed-width integer types (also under NVRTC, which has no host headers)
#ifdef CGX_CPU_SIM
// test harness: the kernel source compiled for the CPU SIMT emulator (tests/sim/simt_cpu.h), which
// supplies the mbarrier / bulk-copy / status-word helpers below as plain C++
#include "simt_cpu.h"
#else
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#define CGX_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

namespace cgx {

#ifndef CGX_CPU_SIM
__device__ __forceinline__ void cgx_spin_yield() {}
__device__ __forceinline__ void cgx_threadfence() { __threadfence(); }
__device__ __forceinline__ void cgx_fence_block() { __threadfence_block(); }
__device__ __forceinline__ void cgx_syncthreads() { __syncthreads(); }
// polite spin: the waiting warp leaves the issue slots to the warps that do the work
#ifndef CGX_BACKOFF_NS
#define CGX_BACKOFF_NS 200
#endif
#ifndef CGX_IDLE_NS
#define CGX_IDLE_NS 1000
#endif
__device__ __forceinline__ void cgx_backoff() { __nanosleep(CGX_BACKOFF_NS); }
__device__ __forceinline__ void cgx_idle() { __nanosleep(CGX_IDLE_NS); }
// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS: UBLKCP) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// asks L2 to fetch [src, src+bytes) (16-byte aligned, multiple of 16); no completion to wait for
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

#endif  // !CGX_CPU_SIM

// ---- SWAR byte-class tests ---------------------------------------------------------------------
// bit 7 of each byte of the result is set iff that byte of x lies in [lo,hi]; requires hi <= 0x7F.
__device__ __forceinline__ uint32_t swar_in_range(uint32_t x, uint32_t k_lo, uint32_t k_hi) {
  // k_lo = 0x80808080 - lo*0x01010101 ; k_hi = (hi*0x01010101) | 0x80808080
  uint32_t x7 = x & 0x7F7F7F7Fu;
  uint32_t ge = x7 + k_lo;  // bit7 set iff x7 >= lo
  uint32_t le = k_hi - x7;  // bit7 set iff x7 <= hi
  return ge & le & ~x & 0x80808080u;
}
__device__ __forceinline__ uint32_t swar_klo(uint32_t lo) { return 0x80808080u - lo * 0x01010101u; }
__device__ __forceinline__ uint32_t swar_khi(uint32_t hi) { return (hi * 0x01010101u) | 0x80808080u; }

// gather the four bit-7 flags of m (0x80 or 0x00 per byte) into bits 0..3
__device__ __forceinline__ uint32_t pack4(uint32_t m) {
  return ((m >> 7) * 0x00204081u >> 21) & 0xFu;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

// ---- decoupled look-back status words -----------------------------------------------------------
// bits 63..62: 0 = empty, 1 = aggregate of this chunk, 2 = inclusive prefix up to this chunk
constexpr unsigned long long LB_AGG = 1ull << 62;
constexpr unsigned long long LB_PREFIX = 2ull << 62;
constexpr unsigned long long LB_VALUE = (1ull << 62) - 1;

#ifndef CGX_CPU_SIM
// The status word carries flag and value together and guards no other data, so relaxed
// gpu-scope accesses are enough (an acquire load would add an L1 invalidate to every poll).
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// named barriers: producers arrive, consumers sync (count = total participating threads)
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

#endif  // !CGX_CPU_SIM

}  // namespace cgx
