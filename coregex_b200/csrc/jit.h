// jit.h — NVRTC specialisation of the bitstream kernel (see jit.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "scan_params.h"

namespace cgx {

struct JitKernel {
  void* func = nullptr;  // CUfunction
  int per_sm = 0;        // resident CTAs per SM
  int smem = 0, threads = 0, warps = 0;  // launch shape, read from the module (cgx_flat_jit_info)
  int mode = 0;          // the search mode the kernel was specialised for (ScanMode)
  int chunk_bytes = 0;   // bytes between chunk origins of this build (the host sizes the look-back arrays by it)
};

// NVRTC only: works without a device (used by the CPU test that the specialised source builds).
// One kernel per (pattern, search mode): FindAll, Count and IsMatch code never share a hot loop.
bool JitCompileCubin(const FlatDev& f, int mode, std::vector<char>& cubin, std::string& err);
// compiled + loaded kernel for the current device, cached per program and mode; nullptr (and err)
// when NVRTC or the driver entry points are unavailable
const JitKernel* GetJitKernel(const FlatDev& f, int mode, std::string& err);
// grid_out (optional): CTAs launched
cudaError_t launch_scan_flat_jit(const JitKernel* k, const ScanArgs& a, int sm_count, cudaStream_t stream,
                                 int* grid_out = nullptr);

}  // namespace cgx
