// jit.cu — per-pattern specialisation of the bitstream kernel (scan_bits.cu) with NVRTC.
//
// The reference compiles a pattern into data (NFA, lazy-DFA tables) that one generic loop
// interprets (reference meta/compile.go:40-219, dfa/lazy/lazy.go:219-324).  On the GPU the marker
// passes of a flat pattern are a dozen bit-vector operations: interpreting them costs more
// instructions (dispatch, class selection) than executing them, and a generic kernel that can run
// any program is ~17k SASS instructions — larger than the instruction cache.  So Compile() -> first
// device scan builds the SAME source (embedded at build time, jit_sources.inc) with -DCGX_JIT and a
// generated "cgx_jit_prog.h" (host/engine.cpp JitHeader): straight-line passes, class constants as
// immediates, ~4x less code.  NVRTC is dlopen'ed; the cubin is loaded and launched through driver
// entry points obtained from the runtime (no link-time dependency on libcuda/libnvrtc).  If NVRTC
// is not available the generic nvcc-built kernel runs instead (still on the GPU; cgx_engine says so).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "host/engine.h"
#include "jit.h"
#include "scan_params.h"

namespace cgx {

size_t scan_flat_smem_bytes();
int scan_flat_threads();
int scan_flat_warps();

namespace {

#include "jit_sources.inc"  // kSrcScanFlat, kSrcScanCommon, kSrcScanParams (build.py)

struct Nvrtc {
  void* h = nullptr;
  decltype(&nvrtcCreateProgram) create = nullptr;
  decltype(&nvrtcCompileProgram) compile = nullptr;
  decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
  decltype(&nvrtcGetCUBIN) cubin = nullptr;
  decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
  decltype(&nvrtcGetProgramLog) log = nullptr;
  decltype(&nvrtcDestroyProgram) destroy = nullptr;
  std::string err;
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
      n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.h) break;
    }
    if (!n.h) {
      n.err = "libnvrtc not found";
      return;
    }
#define SYM(field, name)                                   \
  n.field = (decltype(n.field))dlsym(n.h, name);           \
  if (!n.field) n.err = std::string("missing symbol ") + name;
    SYM(create, "nvrtcCreateProgram")
    SYM(compile, "nvrtcCompileProgram")
    SYM(cubin_size, "nvrtcGetCUBINSize")
    SYM(cubin, "nvrtcGetCUBIN")
    SYM(log_size, "nvrtcGetProgramLogSize")
    SYM(log, "nvrtcGetProgramLog")
    SYM(destroy, "nvrtcDestroyProgram")
#undef SYM
  });
  return n;
}

struct Driver {
  decltype(&cuModuleLoadData) load = nullptr;
  decltype(&cuModuleGetFunction) getfn = nullptr;
  decltype(&cuModuleGetGlobal) getglobal = nullptr;
  decltype(&cuFuncSetAttribute) setattr = nullptr;
  decltype(&cuOccupancyMaxActiveBlocksPerMultiprocessor) occ = nullptr;
  decltype(&cuLaunchKernel) launch = nullptr;
  std::string err;
};

Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [&](const char* name, void** fp) {
      cudaDriverEntryPointQueryResult st;
      if (cudaGetDriverEntryPoint(name, fp, cudaEnableDefault, &st) != cudaSuccess || !*fp)
        d.err = std::string("driver entry point ") + name;
    };
    get("cuModuleLoadData", (void**)&d.load);
    get("cuModuleGetFunction", (void**)&d.getfn);
    get("cuModuleGetGlobal", (void**)&d.getglobal);
    get("cuFuncSetAttribute", (void**)&d.setattr);
    get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&d.occ);
    get("cuLaunchKernel", (void**)&d.launch);
  });
  return d;
}

std::mutex g_mu;
std::map<std::string, std::vector<char>> g_cubins;          // header text -> cubin
std::map<std::pair<int, std::string>, JitKernel*> g_loaded;  // (device, header text) -> kernel

}  // namespace

bool JitCompileCubin(const FlatDev& f, int mode, std::vector<char>& cubin, std::string& err) {
  const std::string hdr = JitHeader(f);
  // experiments: extra -D options for the specialised build, e.g. CGX_JIT_DEFS="-DCGX_ROT=0 -DCGX_BACKOFF_NS=100"
  const char* extra = getenv("CGX_JIT_DEFS");
  std::vector<std::string> xo;
  if (extra) {
    std::string e(extra), cur;
    for (char ch : e + " ") {
      if (ch == ' ') {
        if (!cur.empty()) xo.push_back(cur);
        cur.clear();
      } else {
        cur += ch;
      }
    }
  }
  const std::string key = hdr + "#m" + std::to_string(mode) + (extra ? extra : "");
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cubins.find(key);
    if (it != g_cubins.end()) {
      cubin = it->second;
      return true;
    }
  }
  Nvrtc& n = nvrtc();
  if (!n.err.empty()) {
    err = n.err;
    return false;
  }
  const char* hsrc[] = {kSrcScanCommon, kSrcScanParams, hdr.c_str()};
  const char* hname[] = {"scan_common.cuh", "scan_params.h", "cgx_jit_prog.h"};
  nvrtcProgram prog;
  if (n.create(&prog, kSrcScanBits, "scan_bits.cu", 3, hsrc, hname) != NVRTC_SUCCESS) {
    err = "nvrtcCreateProgram failed";
    return false;
  }
  const std::string mode_def = "-DCGX_JIT_MODE=" + std::to_string(mode);
  // (-default-device: the generic lambda of the kernel's tile loop has no execution-space annotation)
  std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DCGX_JIT=1",
                                   "-default-device", mode_def.c_str()};
  for (auto& o : xo) opts.push_back(o.c_str());
  const nvrtcResult rc = n.compile(prog, (int)opts.size(), opts.data());
  if (rc != NVRTC_SUCCESS) {
    size_t ls = 0;
    n.log_size(prog, &ls);
    std::string log(ls, '\0');
    if (ls) n.log(prog, &log[0]);
    err = "nvrtc: " + log.substr(0, 2000);
    n.destroy(&prog);
    return false;
  }
  size_t cs = 0;
  n.cubin_size(prog, &cs);
  cubin.resize(cs);
  n.cubin(prog, cubin.data());
  n.destroy(&prog);
  std::lock_guard<std::mutex> lk(g_mu);
  g_cubins[key] = cubin;
  return true;
}

const JitKernel* GetJitKernel(const FlatDev& f, int mode, std::string& err) {
  static const bool off = [] {
    const char* e = getenv("CGX_JIT");
    return e && e[0] == '0';
  }();
  if (off) {
    err = "disabled by CGX_JIT=0";
    return nullptr;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    err = "cudaGetDevice failed";
    return nullptr;
  }
  const char* extra = getenv("CGX_JIT_DEFS");
  const std::string hdr = JitHeader(f) + "#m" + std::to_string(mode) + (extra ? extra : "");
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_loaded.find({dev, hdr});
    if (it != g_loaded.end()) return it->second;
  }
  std::vector<char> cubin;
  if (!JitCompileCubin(f, mode, cubin, err)) return nullptr;
  Driver& d = driver();
  if (!d.err.empty()) {
    err = d.err;
    return nullptr;
  }
  cudaFree(0);  // make sure the primary context is current
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  CUresult r = d.load(&mod, cubin.data());
  if (r == CUDA_SUCCESS) r = d.getfn(&fn, mod, "cgx_flat_jit");
  // the module describes its own launch shape: {dynamic smem bytes, threads, scanning warps, CTAs/SM, chunk bytes}
  int info[5] = {0, 0, 0, 0, 0};
  CUdeviceptr ip = 0;
  size_t isz = 0;
  if (r == CUDA_SUCCESS) r = d.getglobal(&ip, &isz, mod, "cgx_flat_jit_info");
  if (r == CUDA_SUCCESS && (isz != sizeof info ||
                            cudaMemcpy(info, (const void*)ip, sizeof info, cudaMemcpyDeviceToHost) != cudaSuccess))
    r = CUDA_ERROR_UNKNOWN;
  const int smem = info[0];
  if (r == CUDA_SUCCESS) r = d.setattr(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem);
  int per_sm = 0;
  if (r == CUDA_SUCCESS) r = d.occ(&per_sm, fn, info[1], (size_t)smem);
  if (r != CUDA_SUCCESS || per_sm < 1) {
    char b[96];
    snprintf(b, sizeof b, "loading the JIT cubin failed (CUresult %d, %d blocks/SM)", (int)r, per_sm);
    err = b;
    return nullptr;
  }
  JitKernel* k = new JitKernel();
  k->func = (void*)fn;
  k->per_sm = per_sm;
  k->smem = smem;
  k->threads = info[1];
  k->warps = info[2];
  k->mode = mode;
  k->chunk_bytes = info[4];
  std::lock_guard<std::mutex> lk(g_mu);
  g_loaded[{dev, hdr}] = k;
  return k;
}

cudaError_t launch_scan_flat_jit(const JitKernel* k, const ScanArgs& a, int sm_count, cudaStream_t stream,
                                 int* grid_out) {
  if (a.nchunks == 0) return cudaSuccess;
  Driver& d = driver();
  int64_t grid = (int64_t)sm_count * k->per_sm;
  const int64_t need = (a.nchunks + k->warps - 1) / k->warps;
  if (grid > need) grid = need;
  if (grid_out) *grid_out = (int)grid;
  void* params[] = {(void*)&a};
  const CUresult r = d.launch((CUfunction)k->func, (unsigned)grid, 1, 1, (unsigned)k->threads, 1, 1,
                              (unsigned)k->smem, (CUstream)stream, params, nullptr);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorLaunchFailure;
}

}  // namespace cgx
