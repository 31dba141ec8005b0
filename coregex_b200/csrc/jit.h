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
  int tiles = 0;         // tiles evaluated jointly per iteration (CGX_TILES)
};

// tiles per iteration the specialised kernel is built with: $CGX_TILES (1 or 2), default kJitTilesDefault
constexpr int kJitTilesDefault = 1;
int JitTiles();

// NVRTC only: works without a device (used by the CPU test that the specialised source builds)
bool JitCompileCubin(const FlatDev& f, int tiles, std::vector<char>& cubin, std::string& err);
// compiled + loaded kernel for the current device, cached per program; nullptr (and err) when
// NVRTC or the driver entry points are unavailable
const JitKernel* GetJitKernel(const FlatDev& f, std::string& err);
cudaError_t launch_scan_flat_jit(const JitKernel* k, const ScanArgs& a, int sm_count, cudaStream_t stream);

}  // namespace cgx
