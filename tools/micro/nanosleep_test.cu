// micro-benchmark: how long does __nanosleep(t) really take on this GPU, idle and with busy neighbours?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(int ns, int busy_warps, long long* out, int iters) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < busy_warps) {  // busy neighbours: integer ALU work
    unsigned x = threadIdx.x;
    for (int i = 0; i < iters * 4000; i++) x = x * 1664525u + 1013904223u;
    if (x == 12345u) out[1000] = x;
    return;
  }
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) __nanosleep(ns);
  long long t1 = clock64();
  if (lane == 0 && warp == busy_warps) out[blockIdx.x] = (t1 - t0) / iters;
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 2048);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int busy : {0, 7, 15}) for (int ns : {20, 100, 200, 500, 2000}) {
    k<<<148 * 2, (busy + 1) * 32>>>(ns, busy, d, 200);
    cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("busy_warps=%2d nanosleep(%4d): %lld cycles = %.2f us\n", busy, ns, h[0], h[0] / (clk / 1e3));
  }
  return 0;
}
