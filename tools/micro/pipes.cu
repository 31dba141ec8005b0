// pipes.cu — issue rate of the integer instructions the bitstream kernel is made of, alone and in
// pairs, to see which share an execution pipe on B200 (the kernel is ALU-pipe bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipes tools/micro/pipes.cu && gpurun_out/pipes
// Output: warp-instructions per cycle per SM sub-partition (4 per SM) for each mix.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int MIX>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, uint32_t one) {
  uint32_t r[CHAINS], q[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    r[i] = seed * (threadIdx.x + 1) + i;
    q[i] = seed + i * 77;
  }
  const uint32_t m = seed | 0x7f7f7f7f, kk = seed ^ 0x30303030;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      if (MIX == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x28;" : "+r"(r[i]) : "r"(kk), "r"(m));
      if (MIX == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(one), "r"(kk));
      if (MIX == 2) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x28;" : "+r"(r[i]) : "r"(kk), "r"(m));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(one), "r"(kk));
      }
      if (MIX == 3) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(kk), "r"(m));
      if (MIX == 4) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x28;" : "+r"(r[i]) : "r"(kk), "r"(m));
        asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(kk), "r"(m));
      }
      if (MIX == 5) {
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(one), "r"(kk));
        asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(kk), "r"(m));
      }
      if (MIX == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(r[i]) : "r"(q[i]));
      if (MIX == 7) {  // 64-bit a*one + c: IMAD.WIDE
        uint64_t w = ((uint64_t)q[i] << 32) | r[i];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(r[i]), "r"(one));
        r[i] = (uint32_t)w;
        q[i] = (uint32_t)(w >> 32);
      }
      if (MIX == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(kk));
      if (MIX == 9) asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
      if (MIX == 10) asm volatile("brev.b32 %0, %0;" : "+r"(r[i]));
      if (MIX == 11) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(kk));
      if (MIX == 12) {
        asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %0, %1, p;}" : "+r"(r[i]) : "r"(q[i]));
      }
      if (MIX == 13) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(kk), "r"(q[i]));
      if (MIX == 14) {  // three-way: LOP3 + IMAD + IDP
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x28;" : "+r"(r[i]) : "r"(kk), "r"(m));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(one), "r"(kk));
        asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(q[(i + 1) % CHAINS]) : "r"(kk), "r"(m));
      }
    }
    if (MIX == 15) {
#pragma unroll
      for (int i = 0; i < CHAINS; i++) r[i] = __shfl_up_sync(0xffffffffu, r[i], 1);
    }
    if (MIX == 16) {
#pragma unroll
      for (int i = 0; i < CHAINS; i++) r[i] += __ballot_sync(0xffffffffu, r[i] & 1);
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc ^= r[i] ^ q[i];
  if (acc == 0x12345) out[threadIdx.x] = acc;
}

template <int MIX>
void run(const char* name, int per_iter, uint32_t* d, int sms, double mhz) {
  const int grid = sms * 4;
  k<MIX><<<grid, 256>>>(d, 12345u, 1u);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MIX><<<grid, 256>>>(d, 12345u, 1u);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  // per sub-partition: 4 CTAs x 8 warps / 4 = 8 warps, each issuing ITERS*CHAINS*per_iter instructions
  const double instr = 8.0 * ITERS * CHAINS * per_iter;
  const double cycles = ms * 1e-3 * mhz * 1e6;
  printf("%-28s %7.3f ms  %.3f warp-instr/clk/SMSP (at %.0f MHz)\n", name, ms, instr / cycles, mhz);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  uint32_t* d;
  cudaMalloc(&d, 4096);
  const int sms = p.multiProcessorCount;
  run<0>("LOP3", 1, d, sms, mhz);
  run<1>("IMAD", 1, d, sms, mhz);
  run<2>("LOP3+IMAD", 2, d, sms, mhz);
  run<3>("IDP4A", 1, d, sms, mhz);
  run<4>("LOP3+IDP4A", 2, d, sms, mhz);
  run<5>("IMAD+IDP4A", 2, d, sms, mhz);
  run<6>("SHF", 1, d, sms, mhz);
  run<7>("IMAD.WIDE", 1, d, sms, mhz);
  run<8>("IADD", 1, d, sms, mhz);
  run<9>("POPC", 1, d, sms, mhz);
  run<10>("BREV", 1, d, sms, mhz);
  run<11>("IMAD.HI", 1, d, sms, mhz);
  run<12>("ISETP+SEL", 2, d, sms, mhz);
  run<13>("PRMT", 1, d, sms, mhz);
  run<14>("LOP3+IMAD+IDP4A", 3, d, sms, mhz);
  run<15>("SHFL.UP", 1, d, sms, mhz);
  run<16>("LOP3.P+VOTE+IADD (3 instr)", 3, d, sms, mhz);
  return 0;
}
