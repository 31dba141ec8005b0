// wire.cu — compact wire format for gathering match offsets between GPUs (BASELINE
// config 5: "NCCL gather of match offsets").  A shard's matches are sorted int64 (start,end) pairs,
// 16 B each; on the wire a match is the low 32 bits of its shard-relative start plus a 16-bit
// length (6 B), and a small table says at which match index each 4 GiB segment of the shard
// begins, so the receiver can rebuild the high bits: start = base + (segment << 32) + lo.
// Layout of one shard's message (all little endian, `count` known to both sides from the count
// gather):  seg_first[nseg] u64 | lo[count] u32 | len[count] u16.
// A match longer than 65535 bytes does not fit: the packer counts those in *d_bad and the caller
// falls back to the plain int64 format.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/coregex_b200.h"

namespace {

__global__ void pack_kernel(const int64_t* __restrict__ pairs, uint64_t count, int64_t base, uint64_t* seg_first,
                            int nseg, uint32_t* __restrict__ lo, uint16_t* __restrict__ len,
                            unsigned long long* bad) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // segment table: first match whose shard-relative start is >= s << 32 (binary search, one thread each)
  if (tid < (uint64_t)nseg) {
    const int64_t key = base + ((int64_t)tid << 32);
    uint64_t a = 0, b = count;
    while (a < b) {
      const uint64_t m = (a + b) >> 1;
      if (pairs[2 * m] < key) a = m + 1;
      else b = m;
    }
    seg_first[tid] = a;
  }
  unsigned long long nbad = 0;
  for (uint64_t i = tid; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    const longlong2 p = reinterpret_cast<const longlong2*>(pairs)[i];
    const int64_t rel = p.x - base, l = p.y - p.x;
    lo[i] = (uint32_t)rel;
    len[i] = (uint16_t)l;
    if (l < 0 || l > 0xFFFF || rel < 0 || (rel >> 32) >= nseg) nbad++;
  }
  if (nbad) atomicAdd(bad, nbad);
}

__global__ void unpack_kernel(const uint64_t* __restrict__ seg_first, int nseg, const uint32_t* __restrict__ lo,
                              const uint16_t* __restrict__ len, uint64_t count, int64_t base,
                              int64_t* __restrict__ pairs) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x) {
    int seg = 0;
    for (int s = 1; s < nseg; s++) seg += seg_first[s] <= i ? 1 : 0;  // nseg is tiny (shard bytes >> 32)
    const int64_t start = base + ((int64_t)seg << 32) + (int64_t)lo[i];
    reinterpret_cast<longlong2*>(pairs)[i] = make_longlong2(start, start + (int64_t)len[i]);
  }
}

unsigned grid_for(uint64_t count) {
  uint64_t g = (count + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)g;
}

}  // namespace

extern "C" {

size_t cgx_wire_bytes(size_t count, int nseg) {
  return (size_t)nseg * 8 + count * 4 + ((count * 2 + 7) & ~(size_t)7);
}

int cgx_wire_segments(size_t shard_len) { return (int)((shard_len + (((size_t)1 << 32) - 1)) >> 32) + 1; }

int cgx_pack_offsets_device(const int64_t* d_pairs, size_t count, int64_t shard_base, size_t shard_len,
                            uint8_t* d_wire, uint64_t* d_bad, void* stream) {
  if (!d_wire || !d_bad || (count && !d_pairs) || ((uintptr_t)d_wire & 7) || ((uintptr_t)d_pairs & 15))
    return CGX_ERR_ARGS;
  const int nseg = cgx_wire_segments(shard_len);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(d_bad, 0, 8, st) != cudaSuccess) return CGX_ERR_CUDA;
  uint64_t* seg = reinterpret_cast<uint64_t*>(d_wire);
  uint32_t* lo = reinterpret_cast<uint32_t*>(d_wire + (size_t)nseg * 8);
  uint16_t* len = reinterpret_cast<uint16_t*>(d_wire + (size_t)nseg * 8 + count * 4);
  pack_kernel<<<grid_for(count), 256, 0, st>>>(d_pairs, count, shard_base, seg, nseg, lo, len,
                                               reinterpret_cast<unsigned long long*>(d_bad));
  return cudaGetLastError() == cudaSuccess ? CGX_OK : CGX_ERR_CUDA;
}

int cgx_unpack_offsets_device(const uint8_t* d_wire, size_t count, int64_t shard_base, size_t shard_len,
                              int64_t* d_pairs_out, void* stream) {
  if (!d_wire || (count && !d_pairs_out) || ((uintptr_t)d_wire & 7) || ((uintptr_t)d_pairs_out & 15))
    return CGX_ERR_ARGS;
  if (!count) return CGX_OK;
  const int nseg = cgx_wire_segments(shard_len);
  const uint64_t* seg = reinterpret_cast<const uint64_t*>(d_wire);
  const uint32_t* lo = reinterpret_cast<const uint32_t*>(d_wire + (size_t)nseg * 8);
  const uint16_t* len = reinterpret_cast<const uint16_t*>(d_wire + (size_t)nseg * 8 + count * 4);
  unpack_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(seg, nseg, lo, len, count, shard_base, d_pairs_out);
  return cudaGetLastError() == cudaSuccess ? CGX_OK : CGX_ERR_CUDA;
}

}  // extern "C"
