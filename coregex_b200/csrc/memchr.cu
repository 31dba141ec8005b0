// memchr.cu — the byte-search family of the reference's prefilter layer as device entry points:
//   reference simd/memchr_amd64.go:67 Memchr, :114 Memchr2, :159 Memchr3, :202 MemchrPair
//   reference simd/memchr_digit_amd64.go:17 MemchrDigit, :34 MemchrDigitAt (memchr_digit_amd64.s:26)
//   reference simd/memchr_class_amd64.go:35 MemchrWord, :58 MemchrNotWord, :76 MemchrInTable,
//             :90 MemchrNotInTable
//   reference simd/memmem.go:53 Memmem
// Every one of them is "the FIRST position p whose byte is in a set (and, for the pair and the
// substring search, where a second test holds)".  One kernel: CTAs draw 32 KB chunks in ascending
// order from a ticket counter, every thread tests 16 bytes per step against a 256-bit set held in
// registers-as-shared memory, a hit lowers the global minimum with atomicMin, and a CTA whose chunk
// starts at or beyond the minimum stops drawing — the scan ends shortly after the first hit instead
// of reading the whole buffer.  Inside the scan kernels the same tests are fused into phase A
// (scan_bits.cu class bitmaps, scan_dfa.cu byte-set filter); these entry points exist for callers
// that use the prefilter layer on its own.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/coregex_b200.h"

namespace cgx {

namespace {

constexpr int MC_THREADS = 256;
constexpr int MC_STEPS = 8;
constexpr int64_t MC_CHUNK = (int64_t)MC_THREADS * 16 * MC_STEPS;  // 32 KB per ticket
constexpr int MC_MAXNEEDLE = 256;

struct McArgs {
  const uint8_t* h;
  int64_t n;
  int64_t from;         // positions before it are no hits (MemchrDigitAt); chunks start at its chunk
  uint32_t set[8];      // candidate bytes (for the pair / substring search: the first byte)
  int mode;             // 0 = set membership, 1 = pair (b2 at +offset), 2 = substring
  uint8_t b2;
  int64_t offset;       // pair: distance of the second byte
  int m;                // substring: needle length (<= MC_MAXNEEDLE)
  uint8_t needle[MC_MAXNEEDLE];
  unsigned long long* best;  // device: lowest hit so far (initialised to ~0)
  unsigned int* ticket;      // device: chunk tickets (initialised to 0)
};

__device__ __forceinline__ bool mc_confirm(const McArgs& a, int64_t p) {
  if (p < a.from) return false;
  if (a.mode == 0) return true;
  if (a.mode == 1) return p + a.offset < a.n && a.h[p + a.offset] == a.b2;
  if (p + a.m > a.n) return false;
  for (int k = 1; k < a.m; k++)
    if (a.h[p + k] != a.needle[k]) return false;
  return true;
}

__global__ void __launch_bounds__(MC_THREADS) memchr_kernel(const __grid_constant__ McArgs a) {
  __shared__ uint32_t s_set[8];
  __shared__ unsigned s_ticket;
  if (threadIdx.x < 8) s_set[threadIdx.x] = a.set[threadIdx.x];
  const int64_t first = a.from / MC_CHUNK;
  const int64_t nchunks = (a.n + MC_CHUNK - 1) / MC_CHUNK - first;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned t = atomicAdd(a.ticket, 1u);
      // tickets ascend: once a hit lies before this chunk no later chunk can hold the first one
      const bool stop = (int64_t)t >= nchunks ||
                        (unsigned long long)(((int64_t)t + first) * MC_CHUNK) >= *(volatile unsigned long long*)a.best;
      s_ticket = stop ? 0xFFFFFFFFu : t;
    }
    __syncthreads();
    if (s_ticket == 0xFFFFFFFFu) return;
    const int64_t c0 = ((int64_t)s_ticket + first) * MC_CHUNK;
    unsigned long long mine = ~0ull;
#pragma unroll 2
    for (int s = 0; s < MC_STEPS; s++) {
      const int64_t p0 = c0 + ((int64_t)s * MC_THREADS + threadIdx.x) * 16;
      if (p0 >= a.n || mine != ~0ull) continue;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (p0 + 16 <= a.n) {
        const uint4 v = *reinterpret_cast<const uint4*>(a.h + p0);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        for (int k = 0; k < 16; k++) {
          const uint32_t b = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
          if (((s_set[b >> 5] >> (b & 31)) & 1u) && mc_confirm(a, p0 + k)) {
            mine = (unsigned long long)(p0 + k);
            break;
          }
        }
      } else {
        for (int64_t p = p0; p < a.n; p++) {
          const uint32_t b = a.h[p];
          if (((s_set[b >> 5] >> (b & 31)) & 1u) && mc_confirm(a, p)) {
            mine = (unsigned long long)p;
            break;
          }
        }
      }
    }
    if (mine != ~0ull) atomicMin(a.best, mine);
  }
}


int mc_launch(McArgs& a, int64_t* d_result, cudaStream_t st) {
  // d_result[0]: the answer (-1 = all ones = "no hit yet" for the unsigned minimum); d_result[1]: tickets
  cudaError_t e = cudaMemsetAsync(d_result, 0xFF, 8, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_result + 1, 0, 8, st);
  if (e != cudaSuccess) return CGX_ERR_CUDA;
  if (a.n <= 0 || a.from >= a.n) return CGX_OK;
  a.best = reinterpret_cast<unsigned long long*>(d_result);
  a.ticket = reinterpret_cast<unsigned int*>(d_result + 1);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (int64_t)sms * 8;
  const int64_t nchunks = (a.n + MC_CHUNK - 1) / MC_CHUNK - a.from / MC_CHUNK;
  if (grid > nchunks) grid = nchunks;
  memchr_kernel<<<(unsigned)grid, MC_THREADS, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? CGX_OK : CGX_ERR_CUDA;
}

}  // namespace

}  // namespace cgx

using namespace cgx;

extern "C" {

int cgx_memchr_table_at_device(const uint8_t* d_h, size_t n, size_t at, const uint8_t* table256, int64_t* d_result,
                               void* stream) {
  if (!d_result || !table256 || (n && !d_h)) return CGX_ERR_ARGS;
  McArgs a{};
  a.h = d_h;
  a.n = (int64_t)n;
  a.from = (int64_t)at;
  for (int b = 0; b < 256; b++)
    if (table256[b]) a.set[b >> 5] |= 1u << (b & 31);
  a.mode = 0;
  return mc_launch(a, d_result, (cudaStream_t)stream);
}

int cgx_memchr_table_device(const uint8_t* d_h, size_t n, const uint8_t* table256, int64_t* d_result, void* stream) {
  return cgx_memchr_table_at_device(d_h, n, 0, table256, d_result, stream);
}

int cgx_memchr_pair_device(const uint8_t* d_h, size_t n, uint8_t byte1, uint8_t byte2, int64_t offset, int64_t* d_result,
                           void* stream) {
  if (!d_result || (n && !d_h)) return CGX_ERR_ARGS;
  McArgs a{};
  a.h = d_h;
  a.n = (int64_t)n;
  // reference simd/memchr_amd64.go:203-221: negative offset, a haystack no longer than the offset and
  // "same position, different bytes" find nothing
  if (offset < 0 || (int64_t)n <= offset || (offset == 0 && byte1 != byte2)) a.n = 0;
  a.set[byte1 >> 5] |= 1u << (byte1 & 31);
  a.mode = offset == 0 ? 0 : 1;
  a.b2 = byte2;
  a.offset = offset;
  return mc_launch(a, d_result, (cudaStream_t)stream);
}

int cgx_memmem_device(const uint8_t* d_h, size_t n, const uint8_t* needle, size_t m, int64_t* d_result, void* stream) {
  if (!d_result || (n && !d_h) || (m && !needle)) return CGX_ERR_ARGS;
  if (m > (size_t)MC_MAXNEEDLE) return CGX_ERR_UNSUPPORTED;
  McArgs a{};
  a.h = d_h;
  a.n = (int64_t)n;
  if (m == 0) {
    // reference simd/memmem.go:58-61: the empty needle matches at 0 (bytes.Index behaviour)
    const cudaError_t e = cudaMemsetAsync(d_result, 0, 16, (cudaStream_t)stream);
    return e == cudaSuccess ? CGX_OK : CGX_ERR_CUDA;
  }
  if (m > n) a.n = 0;
  a.set[needle[0] >> 5] |= 1u << (needle[0] & 31);
  a.mode = m == 1 ? 0 : 2;
  a.m = (int)m;
  for (size_t k = 0; k < m; k++) a.needle[k] = needle[k];
  return mc_launch(a, d_result, (cudaStream_t)stream);
}

}  // extern "C"
