// capi.cu — the C ABI (include/coregex_b200.h): device residency of compiled tables, scratch
// management, host-buffer wrappers (H2D -> kernel -> D2H) and the device-resident batch entry.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/coregex_b200.h"
#include "host/engine.h"
#include "jit.h"
#include "scan_params.h"
#include "synth.h"

namespace cgx {
size_t scan_dfa_smem_bytes(int nstates, int blob_bytes);
int64_t scan_dfa_chunks(int64_t n);
cudaError_t launch_scan_dfa(const ScanArgs& a, int sm_count, cudaStream_t stream);
int64_t scan_flat_chunks(int64_t n);
cudaError_t launch_scan_flat(const ScanArgs& a, int sm_count, cudaStream_t stream, int* grid_out);
int64_t scan_teddy_chunks(int64_t n);
cudaError_t launch_scan_teddy(const ScanArgs& a, int sm_count, cudaStream_t stream, int* grid_out);
int64_t pike_search_slices(int64_t n);
size_t pike_search_scratch_bytes(int64_t n);
cudaError_t launch_pike_search(const uint8_t* h, int64_t n, int64_t base, int64_t after, const uint32_t* code,
                               const uint32_t* sets, int ninst, int nthreads, int start_pc, int delim, int mode,
                               int64_t* out, int64_t cap, void* scratch, unsigned long long* total, cudaStream_t st,
                               int* launches);
cudaError_t launch_flat_captures(const uint8_t* h, int64_t n, int64_t base, const int64_t* matches,
                                 const unsigned long long* d_total, unsigned long long cap, const FlatDev& f,
                                 const uint8_t* at, int nslots, int64_t* out, cudaStream_t stream);
cudaError_t launch_pike_captures(const uint8_t* h, int64_t n, int64_t base, const int64_t* matches,
                                 const unsigned long long* d_total, unsigned long long cap,
                                 const uint32_t* code, const uint32_t* sets, int start_pc, int nslots,
                                 int64_t* out, cudaStream_t stream, bool large, bool text_end);
}  // namespace cgx

using namespace cgx;

static thread_local std::string g_last_error;

static int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? CGX_ERR_NO_DEVICE : CGX_ERR_CUDA;
}
#define CU(call)                                     \
  do {                                               \
    cudaError_t _e = (call);                         \
    if (_e != cudaSuccess) return cuda_fail(_e, #call); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return CGX_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, n);
      want = n;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cap = want;
    return CGX_OK;
  }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

struct cgx_regex {
  std::unique_ptr<Compiled> c;
  std::mutex mu;
  int device = -1;
  int sm_count = 0;
  // device copies of the tables
  DevBuf d_trans, d_eoi, d_lut, d_teddy, d_line, d_pike_search;
  TeddyDev teddy_dev;
  LineDev line_dev;
  // per-call scratch (serialised by mu)
  DevBuf d_ticket_total, d_status, d_hay, d_out, d_pairs, d_pike;
  // pipelined host path: two haystack/output slots, three streams, pinned per-slot results
  DevBuf d_hay2[2], d_out2[2];
  cudaStream_t s_h2d = nullptr, s_scan = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_scan[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  uint64_t* pinned_res = nullptr;  // [2][2]
  std::atomic<uint64_t> launches{0};
  // flat deterministic patterns run on the bitstream kernel (scan_bits.cu); CGX_BITSTREAM=0 or
  // cgx_debug_set_bitstream keep them on the candidate/DFA kernel (A/B runs, tests of both paths)
  bool bitstream = true;
  // literal sets: scan_teddy.cu (the bitstream kernel's skeleton) instead of the scan_dfa.cu engine.
  // Off by default: measured slower on B200 (C3 562 vs 891 GB/s, C5 385 vs 964 GB/s) — its per-lane
  // verification through global memory is latency-bound at 6 of 32 lanes.
  // CGX_TEDDY2=1 or cgx_debug_set_bitstream(re, 3) switch it on.
  bool teddy2 = [] {
    const char* e = getenv("CGX_TEDDY2");
    return e && e[0] == '1';
  }();
  // NVRTC-specialised build of that kernel for this pattern: 0 = not tried yet, 1 = in use,
  // -1 = unavailable (the generic nvcc-built kernel runs; jit_error says why)
  int jit_state = 0;
  const JitKernel* jit[3] = {nullptr, nullptr, nullptr};  // one specialised kernel per search mode
  std::string jit_error;
  // bitstream kernel, FindAll: a call is ONE launch.  The look-back words carry an epoch (stale words
  // read as empty), the group accumulators and the ticket counter are left at zero by the kernel.
  uint32_t epoch = 0;          // last epoch handed to a launch; 0 = the words must be cleared first
  size_t status_cap_seen = 0;  // capacity of d_status when its words were last cleared
  bool scratch_zero = false;   // ticket counter known to be zero (left so by the previous launch)
  bool diag = false;           // cgx_debug_scratch in use: clear the diagnostics before every launch
  bool longest = false;        // Regex.Longest(): leftmost-longest tables in use

  int ensure_pipeline() {
    if (s_h2d) return CGX_OK;
    CU(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s_scan, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CU(cudaEventCreateWithFlags(&ev_h2d[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_scan[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_d2h[i], cudaEventDisableTiming));
    }
    CU(cudaHostAlloc((void**)&pinned_res, 4 * sizeof(uint64_t), cudaHostAllocDefault));
    return CGX_OK;
  }
  ~cgx_regex() {
    if (s_h2d) {
      cudaStreamDestroy(s_h2d);
      cudaStreamDestroy(s_scan);
      cudaStreamDestroy(s_d2h);
      for (int i = 0; i < 2; i++) {
        cudaEventDestroy(ev_h2d[i]);
        cudaEventDestroy(ev_scan[i]);
        cudaEventDestroy(ev_d2h[i]);
      }
      cudaFreeHost(pinned_res);
    }
  }

  int ensure_device() {
    if (device >= 0) return CGX_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      g_last_error = "no CUDA device available (this library has no CPU fallback)";
      return CGX_ERR_NO_DEVICE;
    }
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    sm_count = prop.multiProcessorCount;
    if (c->kind == ENG_DFA) {
      int r;
      if ((r = d_trans.ensure(c->dfa.trans.size() * 2))) return r;
      if ((r = d_eoi.ensure(c->dfa.eoi.size()))) return r;
      if ((r = d_lut.ensure(256))) return r;
      CU(cudaMemcpy(d_trans.p, c->dfa.trans.data(), c->dfa.trans.size() * 2, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(d_eoi.p, c->dfa.eoi.data(), c->dfa.eoi.size(), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(d_lut.p, c->lut, 256, cudaMemcpyHostToDevice));
    }
    memset(&line_dev, 0, sizeof line_dev);
    if (c->kind == ENG_LINE) {
      const DfaTables &u = c->udfa, &rv = c->rdfa;
      std::vector<uint16_t> blob(u.trans.size() + rv.trans.size() + (u.nstates + rv.nstates + 3) / 2 + 8, 0);
      memcpy(blob.data(), u.trans.data(), u.trans.size() * 2);
      memcpy(blob.data() + u.trans.size(), rv.trans.data(), rv.trans.size() * 2);
      uint8_t* e = reinterpret_cast<uint8_t*>(blob.data() + u.trans.size() + rv.trans.size());
      memcpy(e, u.eoi.data(), u.nstates);
      memcpy(e + u.nstates, rv.eoi.data(), rv.nstates);
      int r;
      if ((r = d_line.ensure(blob.size() * 2))) return r;
      CU(cudaMemcpy(d_line.p, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice));
      line_dev.blob = (const uint16_t*)d_line.p;
      line_dev.un = u.nstates;
      line_dev.rn = rv.nstates;
      for (int k = 0; k < 5; k++) {
        line_dev.ustart[k] = u.start[k];
        line_dev.rstart[k] = rv.start[k];
        if (u.start[k] != u.start[0]) line_dev.ukinds = 1;
        if (rv.start[k] != rv.start[0]) line_dev.rkinds = 1;
      }
      line_dev.blob_bytes = (int)(blob.size() * 2);
    }
    memset(&teddy_dev, 0, sizeof teddy_dev);
    if (c->kind == ENG_TEDDY) {
      // one allocation: fp[256] u32 | lit8[npat] u64 | mask8[npat] u64 | offs[npat+1] i32 | order[npat] u16 |
      // bucket_off[nb+1] u16 | fp2[256] u16 | bytes
      const TeddyTables& t = c->teddy;
      size_t o_fp = 0, o_lit8 = o_fp + 1024, o_offs = o_lit8 + (size_t)t.npat * 16, o_order = o_offs + (t.npat + 1) * 4;
      size_t o_boff = o_order + ((t.npat * 2 + 3) & ~3), o_fp2 = o_boff + (((t.nbuckets + 1) * 2 + 3) & ~3);
      size_t o_bytes = o_fp2 + 512;
      size_t total = o_bytes + t.bytes.size();
      std::vector<uint8_t> blob((total + 3) & ~(size_t)3, 0);
      memcpy(&blob[o_fp], t.fp_packed.data(), 1024);
      for (int id = 0; id < t.npat; id++) {
        uint64_t v = 0, m = 0;
        const int len = t.offs[id + 1] - t.offs[id];
        for (int k = 0; k < 8 && k < len; k++) {
          v |= (uint64_t)t.bytes[t.offs[id] + k] << (8 * k);
          m |= 0xFFull << (8 * k);
        }
        memcpy(&blob[o_lit8 + 8 * (size_t)id], &v, 8);
        memcpy(&blob[o_lit8 + 8 * (size_t)(t.npat + id)], &m, 8);
      }
      memcpy(&blob[o_offs], t.offs.data(), (t.npat + 1) * 4);
      memcpy(&blob[o_order], t.order_simd.data(), t.npat * 2);
      memcpy(&blob[o_boff], t.bucket_off.data(), (t.nbuckets + 1) * 2);
      for (int id = 0; id < t.npat; id++) {
        uint16_t m;
        const size_t at = o_fp2 + 2 * (size_t)t.bytes[t.offs[id] + 2];
        memcpy(&m, &blob[at], 2);
        m |= (uint16_t)(1u << t.bucket_of[id]);
        memcpy(&blob[at], &m, 2);
      }
      memcpy(&blob[o_bytes], t.bytes.data(), t.bytes.size());
      int r;
      if ((r = d_teddy.ensure(blob.size()))) return r;
      CU(cudaMemcpy(d_teddy.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
      const uint8_t* b = (const uint8_t*)d_teddy.p;
      teddy_dev.fp = (const uint32_t*)(b + o_fp);
      teddy_dev.lit8 = (const uint64_t*)(b + o_lit8);
      teddy_dev.fp2 = (const uint16_t*)(b + o_fp2);
      teddy_dev.use_fp2 = 0;
      for (int id = 0; id < t.npat; id++) {
        auto letter = [](uint8_t ch) { return (ch >= 'a' && ch <= 'z') || (ch >= 'A' && ch <= 'Z') || ch == ' '; };
        if (letter(t.bytes[t.offs[id]]) && letter(t.bytes[t.offs[id] + 1])) teddy_dev.use_fp2 = 1;
      }
      teddy_dev.offs = (const int32_t*)(b + o_offs);
      teddy_dev.order = (const uint16_t*)(b + o_order);
      teddy_dev.bucket_off = (const uint16_t*)(b + o_boff);
      teddy_dev.bytes = b + o_bytes;
      teddy_dev.npat = t.npat;
      teddy_dev.nbuckets = t.nbuckets;
      teddy_dev.min_len = t.min_len;
      teddy_dev.max_len = t.max_len;
      teddy_dev.bytes_len = (int)t.bytes.size();
      teddy_dev.blob_bytes = (int)((total + 3) & ~(size_t)3);
    }
    if (c->kind == ENG_PIKEVM) {
      int r;
      const size_t cb = c->pike_search.code.size() * 4, sb = c->pike_search.sets.size() * 4;
      if ((r = d_pike_search.ensure(cb + sb + 16))) return r;
      CU(cudaMemcpy(d_pike_search.p, c->pike_search.code.data(), cb, cudaMemcpyHostToDevice));
      if (sb) CU(cudaMemcpy((char*)d_pike_search.p + cb, c->pike_search.sets.data(), sb, cudaMemcpyHostToDevice));
    }
    if (c->has_pike) {
      int r;
      size_t cb = c->pike.code.size() * 4, sb = c->pike.sets.size() * 4;
      if ((r = d_pike.ensure(cb + sb + 16))) return r;
      CU(cudaMemcpy(d_pike.p, c->pike.code.data(), cb, cudaMemcpyHostToDevice));
      if (sb) CU(cudaMemcpy((char*)d_pike.p + cb, c->pike.sets.data(), sb, cudaMemcpyHostToDevice));
    }
    device = dev;
    return CGX_OK;
  }
};

extern "C" {

const char* cgx_last_error(void) { return g_last_error.c_str(); }

void cgx_default_config(cgx_config* cfg) {  // reference meta/config.go:101-112
  if (!cfg) return;
  cfg->enable_dfa = 1;
  cfg->enable_prefilter = 1;
  cfg->max_dfa_states = 10000;
  cfg->determinization_limit = 1000;
  cfg->min_literal_len = 1;
  cfg->max_literals = 256;
  cfg->max_recursion_depth = 100;
  cfg->enable_ascii_optimization = 1;
}

// reference meta/config.go:132-170 (Validate) and :179-181 (ConfigError.Error)
int cgx_config_validate(const cgx_config* cfg, char* errbuf, size_t errcap) {
  if (!cfg) return CGX_ERR_ARGS;
  const char* field = nullptr;
  const char* msg = nullptr;
  if (cfg->enable_dfa) {
    if (cfg->max_dfa_states < 1 || cfg->max_dfa_states > 1000000) {
      field = "MaxDFAStates";
      msg = "must be between 1 and 1,000,000";
    } else if (cfg->determinization_limit < 10 || cfg->determinization_limit > 100000) {
      field = "DeterminizationLimit";
      msg = "must be between 10 and 100,000";
    }
  }
  if (!field && cfg->enable_prefilter) {
    if (cfg->min_literal_len < 1 || cfg->min_literal_len > 64) {
      field = "MinLiteralLen";
      msg = "must be between 1 and 64";
    } else if (cfg->max_literals < 1 || cfg->max_literals > 1000) {
      field = "MaxLiterals";
      msg = "must be between 1 and 1,000";
    }
  }
  if (!field && (cfg->max_recursion_depth < 10 || cfg->max_recursion_depth > 1000)) {
    field = "MaxRecursionDepth";
    msg = "must be between 10 and 1,000";
  }
  if (!field) return CGX_OK;
  g_last_error = std::string("regexp: invalid config: ") + field + ": " + msg;
  if (errbuf && errcap) {
    strncpy(errbuf, g_last_error.c_str(), errcap - 1);
    errbuf[errcap - 1] = 0;
  }
  return CGX_ERR_CONFIG;
}

int cgx_compile(const char* pattern, size_t len, cgx_regex** out, char* errbuf, size_t errcap) {
  return cgx_compile_cfg(pattern, len, nullptr, out, errbuf, errcap);
}

int cgx_compile_cfg(const char* pattern, size_t len, const cgx_config* cfg, cgx_regex** out, char* errbuf,
                    size_t errcap) {
  if (out) *out = nullptr;
  if (!pattern || !out) return CGX_ERR_ARGS;
  AnalysisConfig ac;
  if (cfg) {
    const int v = cgx_config_validate(cfg, errbuf, errcap);  // reference meta/compile.go:63-65
    if (v) return v;
    ac.enable_dfa = cfg->enable_dfa != 0;
    ac.enable_prefilter = cfg->enable_prefilter != 0;
    ac.min_literal_len = cfg->min_literal_len;
  }
  std::unique_ptr<Compiled> c;
  std::string err;
  int st = CompilePattern(std::string(pattern, len), c, err, ac);
  if (st != COMPILE_OK) {
    if (errbuf && errcap) {
      strncpy(errbuf, err.c_str(), errcap - 1);
      errbuf[errcap - 1] = 0;
    }
    g_last_error = err;
    return st == COMPILE_SYNTAX ? CGX_ERR_SYNTAX : CGX_ERR_UNSUPPORTED;
  }
  cgx_regex* r = new cgx_regex();
  r->c = std::move(c);
  *out = r;
  return CGX_OK;
}

void cgx_free(cgx_regex* re) { delete re; }

// reference regex.go:464 Longest / meta/engine.go:250 SetLongest.  The reference switches every
// search to its PikeVM in longest mode (meta/find_indices.go:224-226); here the anchored DFA tables
// are rebuilt without the cut at the first match (the table walk then ends at the LAST match state
// before the dead state: leftmost-longest from each start).
int cgx_set_longest(cgx_regex* re, int longest) {
  if (!re) return CGX_ERR_ARGS;
  std::lock_guard<std::mutex> lk(re->mu);
  const bool on = longest != 0;
  if (on == re->longest) return CGX_OK;
  Compiled& c = *re->c;
  if (c.kind == ENG_DFA && c.flat.bs_ok) {
    // flat deterministic pattern: one possible match per start, first == longest
  } else if (c.kind == ENG_DFA) {
    DfaTables t;
    const std::string e = BuildDFA(c.prog, /*anchored=*/true, /*max_states=*/160, t, /*longest=*/on);
    if (!e.empty() || t.matches_empty) {
      g_last_error = "unsupported: leftmost-longest tables: " + (e.empty() ? std::string("pattern can match the empty string") : e);
      return CGX_ERR_UNSUPPORTED;
    }
    c.dfa = t;
    c.kind_lut_needed = false;
    for (int k = 1; k < 5; k++)
      if (c.dfa.start[k] != c.dfa.start[0]) c.kind_lut_needed = true;
    re->device = -1;  // upload the new tables before the next search
  } else if (c.kind == ENG_TEDDY) {
    // a literal alternation: first == longest unless one literal is a proper prefix of another
    const auto& P = c.an.prefixes;
    for (size_t i = 0; i < P.size(); i++)
      for (size_t j = 0; j < P.size(); j++)
        if (i != j && P[i].bytes.size() < P[j].bytes.size() && P[j].bytes.compare(0, P[i].bytes.size(), P[i].bytes) == 0) {
          g_last_error = "unsupported: leftmost-longest for a literal set in which one literal is a prefix of another";
          return CGX_ERR_UNSUPPORTED;
        }
  } else {
    g_last_error = "unsupported: leftmost-longest on the record engine";
    return CGX_ERR_UNSUPPORTED;
  }
  re->longest = on;
  return CGX_OK;
}
const char* cgx_strategy(const cgx_regex* re) { return RefStrategyName(re->c->an.strategy); }
// The engine name says which kernel runs: "...+bitstream" is the NVRTC-specialised kernel; once a
// device scan has found NVRTC unusable (no libnvrtc.so.12, compile error) the name becomes
// "...+bitstream-generic" — the interpreting nvcc-built kernel, about a quarter of the speed
// (cgx_last_error after cgx_debug_jit_state says why).
const char* cgx_engine(const cgx_regex* re) {
  if (re->jit_state < 0 && re->c->kind == ENG_DFA && re->c->flat.bs_ok) {
    thread_local std::string name;
    name = re->c->engine_name + "-generic";
    return name.c_str();
  }
  return re->c->engine_name.c_str();
}
int cgx_num_captures(const cgx_regex* re) { return re->c->prog.num_captures; }
const char* cgx_subexp_name(const cgx_regex* re, int i) {
  if (!re || i < 0 || i >= (int)re->c->prog.cap_names.size()) return "";
  return re->c->prog.cap_names[i].c_str();
}
uint64_t cgx_launch_count(const cgx_regex* re) { return re->launches.load(); }
// 1: the NVRTC-specialised kernel is in use, -1: unavailable (generic kernel; cgx_last_error has
// the reason after this call), 0: no device scan has happened yet / pattern not on that engine
int cgx_debug_jit_state(cgx_regex* re) {
  if (re->jit_state < 0) g_last_error = re->jit_error;
  return re->jit_state;
}
// which kernel FindAllSubmatchIndex runs after the scan: 0 = none (no groups), 1 = item-boundary
// walk of a flat deterministic pattern (flat_caps.cu), 2 = Pike captures kernel, -1 = unsupported.
// No device needed.
int cgx_debug_captures_engine(cgx_regex* re) {
  const Compiled& c = *re->c;
  const int nslots = 2 * c.prog.num_captures;
  if (nslots == 2) return 0;
  if (c.kind == ENG_DFA && c.flat.bs_ok && c.flat_caps.ok && c.flat_caps.nslots == nslots) return 1;
  return c.has_pike ? (c.pike.large ? 3 : 2) : -1;  // 3: the 512-instruction form of the Pike captures kernel
}
// NVRTC only (no device needed): size of the specialised cubin, or -1 with cgx_last_error set
long cgx_debug_jit_compile(cgx_regex* re, char* cubin_out, size_t cap) {
  if (!re->c->flat.bs_ok) {
    g_last_error = "pattern does not run on the bitstream engine";
    return -1;
  }
  std::vector<char> cubin;
  std::string err;
  if (!JitCompileCubin(re->c->flat, CGX_MODE_FINDALL, cubin, err)) {
    g_last_error = err;
    return -1;
  }
  if (cubin_out && cap >= cubin.size()) memcpy(cubin_out, cubin.data(), cubin.size());
  return (long)cubin.size();
}
int cgx_debug_set_bitstream(cgx_regex* re, int on) {
  const int was = re->bitstream ? 1 : 0;
  re->bitstream = on != 0;
  re->teddy2 = on == 3;  // literal sets on the bitstream skeleton (scan_teddy.cu) instead of scan_dfa.cu
  if (on == 2) {  // bitstream engine, but the generic nvcc-built kernel instead of the NVRTC one
    re->jit_state = -1;
    re->jit_error = "specialisation switched off by cgx_debug_set_bitstream(2)";
  } else if (re->jit_state == -1) {
    re->jit_state = 0;
  }
  return was;
}
int cgx_delimiter(const cgx_regex* re) {
  if (!re->c->has_delim) return -1;  // matches may contain every byte value: the haystack is one record
  return re->c->kind == ENG_TEDDY ? '\n' : re->c->delim;
}

static int scan_locked(cgx_regex* re, const uint8_t* d_h, size_t len, int64_t base, int mode,
                       int64_t* d_out, size_t cap, uint64_t* d_result, cudaStream_t st, int64_t after = 0) {
  Compiled& c = *re->c;
  if (((uintptr_t)d_h & 15) || ((uintptr_t)d_out & 15)) {
    g_last_error = "device pointers must be 16-byte aligned";
    return CGX_ERR_ARGS;
  }
  if (c.kind == ENG_PIKEVM) {
    int r;
    if (!c.has_delim && (base != 0 || after != 0)) {
      g_last_error = "a pattern without a record delimiter cannot be scanned in shards";
      return CGX_ERR_ARGS;
    }
    if ((r = re->d_ticket_total.ensure(64))) return r;
    if ((r = re->d_status.ensure(pike_search_scratch_bytes((int64_t)len) + 64))) return r;
    re->epoch = 0;  // the look-back words are used as plain scratch here
    re->scratch_zero = false;
    unsigned long long* tt = (unsigned long long*)re->d_ticket_total.p;
    const uint32_t* code = (const uint32_t*)re->d_pike_search.p;
    int launches = 0;
    CU(launch_pike_search(d_h, (int64_t)len, base, after, code, code + c.pike_search.code.size(), c.pike_search.ninst,
                          c.pike_search.nthreads, c.pike_search.start, c.has_delim ? (int)c.delim : 256, mode, d_out, (int64_t)cap,
                          re->d_status.p, tt, st, &launches));
    re->launches += (uint64_t)launches;
    if (d_result) CU(cudaMemcpyAsync(d_result, tt, 16, cudaMemcpyDeviceToDevice, st));
    return CGX_OK;
  }
  if (c.kind != ENG_DFA && c.kind != ENG_TEDDY && c.kind != ENG_LINE) {
    g_last_error = "engine not available in this build";
    return CGX_ERR_UNSUPPORTED;
  }
  static const bool bs_env = [] {
    const char* e = getenv("CGX_BITSTREAM");
    return !(e && e[0] == '0');
  }();
  const bool use_flat = c.kind == ENG_DFA && c.flat.bs_ok && re->bitstream && bs_env;
  // literal sets can run on the bitstream kernel's skeleton (scan_teddy.cu) when no literal is longer
  // than 32 bytes (its safe-point argument); opt-in, see cgx_regex::teddy2
  const bool use_teddy2 = c.kind == ENG_TEDDY && c.teddy.max_len <= 32 && re->teddy2;
  // the specialised kernel of this pattern and mode (built on first use); it knows its own chunk size
  if (use_flat && re->jit_state >= 0) {
    if (!re->jit[mode]) re->jit[mode] = GetJitKernel(c.flat, mode, re->jit_error);
    re->jit_state = re->jit[mode] ? 1 : -1;
  }
  const bool use_jit = use_flat && re->jit_state == 1;
  const int64_t nchunks = use_jit ? ((int64_t)len + re->jit[mode]->chunk_bytes - 1) / re->jit[mode]->chunk_bytes
                          : use_flat ? scan_flat_chunks((int64_t)len)
                          : use_teddy2 ? scan_teddy_chunks((int64_t)len) : scan_dfa_chunks((int64_t)len);
  if (nchunks >= 0xFFFF0000ll) {
    g_last_error = "haystack too large for one scan call (32-bit chunk tickets); shard it";
    return CGX_ERR_ARGS;
  }
  int r;
  if ((r = re->d_ticket_total.ensure(64))) return r;
  // look-back words: one per chunk, then (bitstream kernel) two words per 32 chunks
  const size_t ngroups = (size_t)(nchunks + 31) / 32 + 1;
  const size_t status_bytes = (size_t)(nchunks > 0 ? nchunks : 1) * 8 + ngroups * 16 + 1024;
  if ((r = re->d_status.ensure(status_bytes))) return r;
  // The buffer is cut by its CAPACITY, not by this call's chunk count: group accumulators | group
  // words | chunk words.  The epoched words and the self-cleaning accumulators survive from launch
  // to launch, so a regex that scans haystacks of different sizes (the pieces of a pipelined host
  // call) must find each of them in the same place every time.
  const size_t status_words = re->d_status.cap / 8;
  const size_t ng_cap = status_words / 34 + 2;  // >= ngroups, and 2 * ng_cap + nchunks <= status_words
  const bool one_launch = (use_flat || use_teddy2) && mode == CGX_MODE_FINDALL && nchunks > 0;
  if (one_launch) {
    if (re->epoch == 0 || re->epoch >= 0xFFFFFu || re->status_cap_seen != re->d_status.cap) {
      CU(cudaMemsetAsync(re->d_status.p, 0, re->d_status.cap, st));  // new buffer, or the epoch counter wrapped
      re->status_cap_seen = re->d_status.cap;
      re->epoch = 0;
    }
    re->epoch++;
    if (!re->scratch_zero || re->diag) CU(cudaMemsetAsync(re->d_ticket_total.p, 0, 64, st));
    re->scratch_zero = true;
  } else {
    CU(cudaMemsetAsync(re->d_ticket_total.p, 0, 64, st));
    if (mode == CGX_MODE_FINDALL && nchunks > 0) {
      CU(cudaMemsetAsync(re->d_status.p, 0, re->d_status.cap, st));
      re->epoch = 0;  // the other kernel wrote un-epoched words
    }
    re->scratch_zero = false;  // (IsMatch leaves the ticket counter wherever the early exit found it)
  }
  ScanArgs a;
  memset(&a, 0, sizeof a);
  a.h = d_h;
  a.n = (int64_t)len;
  a.base = base;
  a.after = after;
  a.dfa.trans = (const uint16_t*)re->d_trans.p;
  a.dfa.eoi = (const uint8_t*)re->d_eoi.p;
  a.dfa.nstates = c.dfa.nstates;
  for (int k = 0; k < 5; k++) a.dfa.start[k] = c.dfa.start[k];
  a.dfa.kind_lut_needed = c.kind_lut_needed ? 1 : 0;
  a.filter.kind = c.filter_kind;
  a.filter.nranges = c.nranges;
  for (int k = 0; k < 4; k++) {
    a.filter.lo[k] = c.rlo[k];
    a.filter.hi[k] = c.rhi[k];
  }
  a.filter.lut = (const uint8_t*)re->d_lut.p;
  a.flat = c.flat;
  a.teddy = re->teddy_dev;
  a.engine = c.kind == ENG_TEDDY ? SEL_TEDDY : SEL_DFA;
  if (c.kind == ENG_TEDDY) {
    a.dfa.nstates = 0;
    a.flat.nops = 0;
    a.filter.kind = F_BYTESET;  // unused by the literal engine; must not point at a LUT
    a.filter.nranges = 0;
    a.filter.lut = nullptr;
    a.skip_safe = 0;
    a.delim = '\n';
  }
  if (c.kind != ENG_TEDDY) {
    a.skip_safe = c.skip_safe ? 1 : 0;
    a.delim = c.delim;
  }
  if (c.kind == ENG_LINE) {
    a.engine = SEL_LINE;
    a.line = re->line_dev;
    a.dfa.nstates = 0;  // the anchored table is not shipped; phase A marks record delimiters
    a.flat.nops = 0;
    a.filter.kind = F_BYTESET;
    a.filter.nranges = 1;
    a.filter.lo[0] = a.filter.hi[0] = c.delim;
    a.filter.lut = nullptr;
    a.skip_safe = 0;
  }
  a.mode = mode;
  a.out = d_out;
  a.cap = (int64_t)cap;
  unsigned long long* tt = (unsigned long long*)re->d_ticket_total.p;
  a.total = tt;                        // [0] count, [1] is-match flag
  a.ticket = (unsigned int*)(tt + 4);  // separate 32-byte sector
  a.gacc = (unsigned long long*)re->d_status.p;
  a.gstatus = a.gacc + ng_cap;
  a.status = a.gstatus + ng_cap;
  a.nchunks = nchunks;
  a.epoch = re->epoch;
  a.result = one_launch ? (unsigned long long*)d_result : nullptr;
  if (use_flat) {
    if (use_jit) CU(launch_scan_flat_jit(re->jit[mode], a, re->sm_count, st));
    else CU(launch_scan_flat(a, re->sm_count, st, nullptr));
  } else if (use_teddy2) {
    CU(launch_scan_teddy(a, re->sm_count, st, nullptr));
  } else {
    CU(launch_scan_dfa(a, re->sm_count, st));
  }
  if (nchunks > 0) re->launches++;
  if (d_result && !one_launch) CU(cudaMemcpyAsync(d_result, tt, 16, cudaMemcpyDeviceToDevice, st));
  return CGX_OK;
}

int cgx_scan_device(cgx_regex* re, const uint8_t* d_h, size_t len, int64_t base, int mode,
                    int64_t* d_out, size_t cap, uint64_t* d_result, void* stream) {
  if (!re) return CGX_ERR_ARGS;
  std::lock_guard<std::mutex> lk(re->mu);
  int r = re->ensure_device();
  if (r) return r;
  return scan_locked(re, d_h, len, base, mode, d_out, cap, d_result, (cudaStream_t)stream);
}

int cgx_scan_shard_device(cgx_regex* re, const uint8_t* d_h, size_t len, int64_t base, int64_t bytes_after,
                          int mode, int64_t* d_out, size_t cap, uint64_t* d_result, void* stream) {
  if (!re || base < 0 || bytes_after < 0) return CGX_ERR_ARGS;
  std::lock_guard<std::mutex> lk(re->mu);
  int r = re->ensure_device();
  if (r) return r;
  return scan_locked(re, d_h, len, base, mode, d_out, cap, d_result, (cudaStream_t)stream, bytes_after);
}

// ---- batch of independent records (the batch entry a GPU-side caller wants) ------
// The reference searches a batch by calling FindAllIndex once per haystack (regex.go:695).  When
// every record ends with the record delimiter no match can span two records, and for a pattern
// without anchors or look-around a match does not depend on where its record begins or ends, so the
// batch equals ONE scan of the concatenation; what remains is to tell the caller which pairs belong
// to which record: prefix[r] = number of matches that start before record r.
__global__ void rec_prefix_kernel(const uint8_t* h, const uint64_t* rec_off, uint64_t nrec, int64_t base,
                                  const int64_t* pairs, const unsigned long long* total, uint64_t cap, uint8_t delim,
                                  uint64_t* prefix, unsigned long long* bad) {
  const uint64_t nm = total[0] < cap ? total[0] : cap;  // pairs actually written
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nrec;
       r += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t off = rec_off[r];
    // the promise that makes the batch one scan: a delimiter before every inner record boundary
    if (r > 0 && r < nrec && (off == 0 || h[off - 1] != delim || off < rec_off[r - 1])) atomicAdd(bad, 1ull);
    const int64_t key = (int64_t)off + base;
    uint64_t lo = 0, hi = nm;  // first pair whose start is >= key
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (pairs[2 * mid] < key) lo = mid + 1;
      else hi = mid;
    }
    prefix[r] = lo;
  }
}

int cgx_scan_records_device(cgx_regex* re, const uint8_t* d_h, size_t len, const uint64_t* d_rec_off, size_t nrec,
                            int64_t base, int64_t* d_out, size_t cap, uint64_t* d_rec_prefix, uint64_t* d_result,
                            void* stream) {
  if (!re || !d_rec_off || !d_rec_prefix || !d_result || base < 0) return CGX_ERR_ARGS;
  std::lock_guard<std::mutex> lk(re->mu);
  int r = re->ensure_device();
  if (r) return r;
  Compiled& c = *re->c;
  if (!c.has_delim) {
    g_last_error = "unsupported: matches of this pattern can contain every byte value, records cannot be batched";
    return CGX_ERR_UNSUPPORTED;
  }
  if (c.an.has_anchors) {
    g_last_error = "unsupported: a pattern with anchors or look-around depends on where each record begins and ends; "
                   "scan such records one call at a time";
    return CGX_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if ((r = scan_locked(re, d_h, len, base, CGX_MODE_FINDALL, d_out, cap, nullptr, st))) return r;
  const unsigned long long* tt = (const unsigned long long*)re->d_ticket_total.p;
  CU(cudaMemcpyAsync(d_result, tt, 16, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemsetAsync(d_result + 2, 0, 8, st));
  const uint8_t delim = c.kind == ENG_TEDDY ? (uint8_t)'\n' : (uint8_t)c.delim;
  const unsigned threads = 256;
  uint64_t blocks = (nrec + 1 + threads - 1) / threads;
  if (blocks > (uint64_t)re->sm_count * 8) blocks = (uint64_t)re->sm_count * 8;
  rec_prefix_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_h, d_rec_off, (uint64_t)nrec, base, d_out, tt, (uint64_t)cap,
                                                          delim, d_rec_prefix, (unsigned long long*)(d_result + 2));
  CU(cudaGetLastError());
  re->launches++;
  return CGX_OK;
}

// scan -> (start,end) pairs -> one Pike lane per match for the group offsets
static int submatch_locked(cgx_regex* re, const uint8_t* d_h, size_t len, int64_t base, int64_t* d_out,
                           size_t cap, uint64_t* d_result, cudaStream_t st, int64_t after = 0) {
  Compiled& c = *re->c;
  const int nslots = 2 * c.prog.num_captures;
  if (nslots == 2)  // no groups: the pairs ARE the result (reference meta/findall.go:109-112)
    return scan_locked(re, d_h, len, base, CGX_MODE_FINDALL, d_out, cap, d_result, st, after);
  // flat deterministic pattern whose groups enclose whole items: the group offsets are item
  // boundaries of the forced greedy walk (flat_caps.cu) — no NFA simulation.  Leftmost-longest
  // changes nothing for a deterministic pattern (one match per start).
  if (c.kind == ENG_DFA && c.flat.bs_ok && c.flat_caps.ok && c.flat_caps.nslots == nslots) {
    int r;
    if ((r = re->d_pairs.ensure((cap ? cap : 1) * 16))) return r;
    if ((r = scan_locked(re, d_h, len, base, CGX_MODE_FINDALL, (int64_t*)re->d_pairs.p, cap, d_result, st, after))) return r;
    CU(launch_flat_captures(d_h, (int64_t)len, base, (const int64_t*)re->d_pairs.p,
                            (const unsigned long long*)re->d_ticket_total.p, cap, c.flat, c.flat_caps.at, nslots, d_out,
                            st));
    re->launches++;
    return CGX_OK;
  }
  if (!c.has_pike) {
    g_last_error = "unsupported: captures kernel limit: " + c.pike_err;
    return CGX_ERR_UNSUPPORTED;
  }
  if (re->longest) {
    g_last_error = "unsupported: submatches in leftmost-longest mode (the captures kernel follows leftmost-first priorities)";
    return CGX_ERR_UNSUPPORTED;
  }
  int r;
  if ((r = re->d_pairs.ensure((cap ? cap : 1) * 16))) return r;
  if ((r = scan_locked(re, d_h, len, base, CGX_MODE_FINDALL, (int64_t*)re->d_pairs.p, cap, d_result, st, after))) return r;
  const uint32_t* code = (const uint32_t*)re->d_pike.p;
  const uint32_t* sets = code + c.pike.code.size();
  CU(launch_pike_captures(d_h, (int64_t)len, base, (const int64_t*)re->d_pairs.p,
                          (const unsigned long long*)re->d_ticket_total.p, cap, code, sets, c.pike.start, nslots,
                          d_out, st, c.pike.large, after == 0));
  re->launches++;
  return CGX_OK;
}

int cgx_scan_submatch_device(cgx_regex* re, const uint8_t* d_h, size_t len, int64_t base, int64_t* d_out,
                             size_t cap, uint64_t* d_result, void* stream) {
  if (!re) return CGX_ERR_ARGS;
  std::lock_guard<std::mutex> lk(re->mu);
  int r = re->ensure_device();
  if (r) return r;
  return submatch_locked(re, d_h, len, base, d_out, cap, d_result, (cudaStream_t)stream);
}

int cgx_scan_submatch_shard_device(cgx_regex* re, const uint8_t* d_h, size_t len, int64_t base, int64_t bytes_after,
                                   int64_t* d_out, size_t cap, uint64_t* d_result, void* stream) {
  if (!re || base < 0 || bytes_after < 0) return CGX_ERR_ARGS;
  std::lock_guard<std::mutex> lk(re->mu);
  int r = re->ensure_device();
  if (r) return r;
  return submatch_locked(re, d_h, len, base, d_out, cap, d_result, (cudaStream_t)stream, bytes_after);
}

// Large host haystacks are cut at record delimiters into pieces that flow through
// H2D(k+1) | scan(k) | D2H(k-1) on three streams, so that the PCIe link is busy in both directions
// while the scan runs.  A piece is independent of its neighbours because no match contains the
// delimiter (the condition the record-parallel kernel already relies on).
static const size_t kPipelineMin = (size_t)64 << 20;
// CGX_PIPELINE_PIECE=<bytes> forces the piece size (tests use it to pipeline small inputs)
static size_t forced_piece() {
  static const size_t v = [] {
    const char* e = getenv("CGX_PIPELINE_PIECE");
    return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)0;
  }();
  return v;
}

// limit > 0: the caller wants the first `limit` matches only — no further piece is queued once
// the pieces finished so far hold that many (the reference's loop stops at n matches,
// meta/findall.go:176-290; pieces are in haystack order, so what they hold is a prefix).
// `nslots` > 2: FindAllSubmatchIndex — every piece also runs its captures pass on the scan stream and
// the rows that travel back are nslots int64 wide instead of pairs.
static int host_scan_pipelined(cgx_regex* re, const uint8_t* h, size_t len, int mode, int64_t* out, size_t cap,
                               uint64_t result[2], int64_t limit, int nslots = 2) {
  int r;
  const size_t row = (size_t)nslots * 8;  // bytes per match in `out`
  auto piece_scan = [&](const uint8_t* d_h, size_t plen, int64_t off, int m, int64_t* d_out, size_t pc,
                        int64_t after) -> int {
    if (nslots > 2 && m == CGX_MODE_FINDALL)
      return submatch_locked(re, d_h, plen, off, d_out, pc, nullptr, re->s_scan, after);
    return scan_locked(re, d_h, plen, off, m, d_out, pc, nullptr, re->s_scan, after);
  };
  if ((r = re->ensure_pipeline())) return r;
  const uint8_t delim = re->c->kind == ENG_TEDDY ? (uint8_t)'\n' : (uint8_t)re->c->delim;
  size_t nominal = len / 16;
  if (nominal < ((size_t)32 << 20)) nominal = (size_t)32 << 20;
  if (nominal > ((size_t)256 << 20)) nominal = (size_t)256 << 20;
  if (forced_piece()) nominal = forced_piece();
  nominal &= ~(size_t)15;
  if (!nominal) nominal = 16;
  // piece boundaries: multiples of 16 bytes are not required for the END of a piece, but every
  // piece is copied to a 16-byte aligned device slot, so any delimiter position will do
  std::vector<size_t> cut{0};
  while (cut.back() < len) {
    size_t b = cut.back() + nominal;
    if (b >= len) {
      cut.push_back(len);
      break;
    }
    const void* q = memrchr(h + cut.back(), delim, b - cut.back());
    if (q) {
      b = (size_t)((const uint8_t*)q - h) + 1;
    } else {
      const void* f = memchr(h + b, delim, len - b);
      b = f ? (size_t)((const uint8_t*)f - h) + 1 : len;
    }
    cut.push_back(b);
  }
  const int np = (int)cut.size() - 1;
  size_t maxlen = 0;
  for (int k = 0; k < np; k++) maxlen = cut[k + 1] - cut[k] > maxlen ? cut[k + 1] - cut[k] : maxlen;
  for (int i = 0; i < 2; i++)
    if ((r = re->d_hay2[i].ensure(maxlen + 16))) return r;

  uint64_t total = 0, flag = 0;
  size_t written = 0;           // pairs delivered to `out`
  size_t pcap[2] = {0, 0};      // output capacity used by the in-flight scan of each slot
  int pmode[2] = {mode, mode};
  bool stop = false;

  auto enqueue = [&](int k) -> int {
    const int sl = k & 1;
    const size_t off = cut[k], plen = cut[k + 1] - cut[k];
    // the slot's previous piece (k-2) was finished — scan consumed, output copied out — by finish(k-2)
    CU(cudaMemcpyAsync(re->d_hay2[sl].p, h + off, plen, cudaMemcpyHostToDevice, re->s_h2d));
    CU(cudaEventRecord(re->ev_h2d[sl], re->s_h2d));
    CU(cudaStreamWaitEvent(re->s_scan, re->ev_h2d[sl], 0));
    int m = mode;
    size_t pc = 0;
    if (mode == CGX_MODE_FINDALL) {
      const size_t remaining = cap - written;  // lower bound: earlier in-flight pieces may still add
      if (remaining == 0) {
        m = CGX_MODE_COUNT;
      } else {
        pc = plen / 16 + 1024;
        if (pc > cap) pc = cap;
        int rr;
        if ((rr = re->d_out2[sl].ensure(pc * row))) return rr;
        CU(cudaStreamWaitEvent(re->s_scan, re->ev_d2h[sl], 0));  // output slot drained (piece k-2)
      }
    }
    pcap[sl] = pc;
    pmode[sl] = m;
    int rr = piece_scan((const uint8_t*)re->d_hay2[sl].p, plen, (int64_t)off, m,
                        pc ? (int64_t*)re->d_out2[sl].p : nullptr, pc, (int64_t)(len - cut[k + 1]));
    if (rr) return rr;
    CU(cudaMemcpyAsync(re->pinned_res + 2 * sl, re->d_ticket_total.p, 16, cudaMemcpyDeviceToHost, re->s_scan));
    CU(cudaEventRecord(re->ev_scan[sl], re->s_scan));
    return CGX_OK;
  };
  auto finish = [&](int k) -> int {
    const int sl = k & 1;
    const size_t off = cut[k], plen = cut[k + 1] - cut[k];
    for (;;) {
      CU(cudaEventSynchronize(re->ev_scan[sl]));
      const uint64_t t = re->pinned_res[2 * sl], f = re->pinned_res[2 * sl + 1];
      if (pmode[sl] == CGX_MODE_FINDALL) {
        const size_t remaining = cap - written;
        const size_t need = t < remaining ? (size_t)t : remaining;
        if (need > pcap[sl]) {
          // denser than estimated: scan this piece again into a buffer that holds what is wanted
          CU(cudaStreamSynchronize(re->s_d2h));
          CU(cudaStreamSynchronize(re->s_scan));
          int rr;
          if ((rr = re->d_out2[sl].ensure(need * row))) return rr;
          pcap[sl] = need;
          if ((rr = piece_scan((const uint8_t*)re->d_hay2[sl].p, plen, (int64_t)off, CGX_MODE_FINDALL,
                               (int64_t*)re->d_out2[sl].p, need, (int64_t)(len - cut[k + 1]))))
            return rr;
          CU(cudaMemcpyAsync(re->pinned_res + 2 * sl, re->d_ticket_total.p, 16, cudaMemcpyDeviceToHost,
                             re->s_scan));
          CU(cudaEventRecord(re->ev_scan[sl], re->s_scan));
          continue;
        }
        if (need) {
          CU(cudaStreamWaitEvent(re->s_d2h, re->ev_scan[sl], 0));
          CU(cudaMemcpyAsync(out + (size_t)nslots * written, re->d_out2[sl].p, need * row, cudaMemcpyDeviceToHost,
                             re->s_d2h));
          written += need;
        }
        CU(cudaEventRecord(re->ev_d2h[sl], re->s_d2h));
      }
      total += t;
      flag |= f;
      if (mode == CGX_MODE_ISMATCH && flag) stop = true;
      if (limit > 0 && total >= (uint64_t)limit) stop = true;
      return CGX_OK;
    }
  };
  int queued = 0;
  for (int k = 0; k < np && !stop; k++) {
    if ((r = enqueue(k))) return r;
    queued = k + 1;
    if (k >= 1 && (r = finish(k - 1))) return r;
  }
  if (queued && !stop && (r = finish(queued - 1))) return r;
  CU(cudaStreamSynchronize(re->s_h2d));
  CU(cudaStreamSynchronize(re->s_scan));
  CU(cudaStreamSynchronize(re->s_d2h));
  result[0] = total;
  result[1] = flag ? 1 : 0;
  return CGX_OK;
}

// host wrapper shared by is_match / count / find_all
static int host_scan(cgx_regex* re, const uint8_t* h, size_t len, int mode, int64_t* out, size_t cap,
                     uint64_t result[2], int64_t limit = 0) {
  if (!re || (!h && len)) return CGX_ERR_ARGS;
  std::lock_guard<std::mutex> lk(re->mu);
  int r = re->ensure_device();
  if (r) return r;
  if (re->c->has_delim && len >= (forced_piece() ? 2 * forced_piece() : kPipelineMin))
    return host_scan_pipelined(re, h, len, mode, out, cap, result, limit);
  if ((r = re->d_hay.ensure(len + 16))) return r;
  if (mode == CGX_MODE_FINDALL && cap && (r = re->d_out.ensure(cap * 16))) return r;
  cudaStream_t st = 0;
  if (len) CU(cudaMemcpyAsync(re->d_hay.p, h, len, cudaMemcpyHostToDevice, st));
  r = scan_locked(re, (const uint8_t*)re->d_hay.p, len, 0, mode,
                  mode == CGX_MODE_FINDALL && cap ? (int64_t*)re->d_out.p : nullptr, cap, nullptr, st);
  if (r) return r;
  CU(cudaMemcpyAsync(result, re->d_ticket_total.p, 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (mode == CGX_MODE_FINDALL && out && cap) {
    size_t w = result[0] < cap ? (size_t)result[0] : cap;
    if (w) CU(cudaMemcpy(out, re->d_out.p, w * 16, cudaMemcpyDeviceToHost));
  }
  return CGX_OK;
}

int cgx_is_match(cgx_regex* re, const uint8_t* h, size_t len, int* matched) {
  uint64_t res[2] = {0, 0};
  int r = host_scan(re, h, len, CGX_MODE_ISMATCH, nullptr, 0, res);
  if (r) return r;
  if (matched) *matched = res[1] ? 1 : 0;
  return CGX_OK;
}

int cgx_count(cgx_regex* re, const uint8_t* h, size_t len, int64_t limit, size_t* count) {
  if (count) *count = 0;
  if (limit == 0) return CGX_OK;
  uint64_t res[2] = {0, 0};
  int r = host_scan(re, h, len, CGX_MODE_COUNT, nullptr, 0, res, limit);
  if (r) return r;
  size_t c = (size_t)res[0];
  if (limit > 0 && c > (size_t)limit) c = (size_t)limit;
  if (count) *count = c;
  return CGX_OK;
}

int cgx_find_all_index(cgx_regex* re, const uint8_t* h, size_t len, int64_t limit, int64_t* out,
                       size_t cap, size_t* count) {
  if (count) *count = 0;
  if (limit == 0) return CGX_OK;  // reference regex.go:696-698
  uint64_t res[2] = {0, 0};
  size_t want = cap;
  if (limit > 0 && (size_t)limit < want) want = (size_t)limit;
  int r = host_scan(re, h, len, (out && want) ? CGX_MODE_FINDALL : CGX_MODE_COUNT, out, want, res, limit);
  if (r) return r;
  size_t c = (size_t)res[0];
  if (limit > 0 && c > (size_t)limit) c = (size_t)limit;  // first `limit` matches are a prefix
  if (count) *count = c;
  return CGX_OK;
}

int cgx_find_all_submatch_index(cgx_regex* re, const uint8_t* h, size_t len, int64_t limit, int64_t* out,
                                size_t cap, size_t* count) {
  if (count) *count = 0;
  if (limit == 0) return CGX_OK;
  if (!re || (!h && len)) return CGX_ERR_ARGS;
  std::lock_guard<std::mutex> lk(re->mu);
  int r = re->ensure_device();
  if (r) return r;
  size_t want = cap;
  if (limit > 0 && (size_t)limit < want) want = (size_t)limit;
  const size_t stride = 2 * (size_t)re->c->prog.num_captures;
  if (re->c->has_delim && out && want &&
      len >= (forced_piece() ? 2 * forced_piece() : kPipelineMin)) {
    // same three-stream pipeline as FindAllIndex: H2D(k+1) | scan + captures(k) | D2H(k-1)
    uint64_t res[2] = {0, 0};
    if ((r = host_scan_pipelined(re, h, len, CGX_MODE_FINDALL, out, want, res, limit, (int)stride))) return r;
    size_t c = (size_t)res[0];
    if (limit > 0 && c > (size_t)limit) c = (size_t)limit;
    if (count) *count = c;
    return CGX_OK;
  }
  if ((r = re->d_hay.ensure(len + 16))) return r;
  if ((r = re->d_out.ensure((want ? want : 1) * stride * 8))) return r;
  cudaStream_t st = 0;
  if (len) CU(cudaMemcpyAsync(re->d_hay.p, h, len, cudaMemcpyHostToDevice, st));
  if ((r = submatch_locked(re, (const uint8_t*)re->d_hay.p, len, 0, (int64_t*)re->d_out.p, want, nullptr, st)))
    return r;
  uint64_t res[2] = {0, 0};
  CU(cudaMemcpyAsync(res, re->d_ticket_total.p, 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  size_t c = (size_t)res[0];
  size_t w = c < want ? c : want;
  if (out && w) CU(cudaMemcpy(out, re->d_out.p, w * stride * 8, cudaMemcpyDeviceToHost));
  if (limit > 0 && c > (size_t)limit) c = (size_t)limit;
  if (count) *count = c;
  return CGX_OK;
}

// ---- debug exports (tests only): the compiled tables exactly as the kernels see them ------------
// record-engine tables for the CPU model in tests/table_model.py (1 = pattern runs on that engine)
int cgx_debug_line_info(const cgx_regex* re, int* un, int* rn, uint16_t ustart[5], uint16_t rstart[5]) {
  const Compiled& c = *re->c;
  if (c.kind != ENG_LINE) return 0;
  *un = c.udfa.nstates;
  *rn = c.rdfa.nstates;
  for (int k = 0; k < 5; k++) {
    ustart[k] = c.udfa.start[k];
    rstart[k] = c.rdfa.start[k];
  }
  return 1;
}
int cgx_debug_line_copy(const cgx_regex* re, uint16_t* ut, uint16_t* rt, uint8_t* ueoi, uint8_t* reoi) {
  const Compiled& c = *re->c;
  if (c.kind != ENG_LINE) return 0;
  memcpy(ut, c.udfa.trans.data(), c.udfa.trans.size() * 2);
  memcpy(rt, c.rdfa.trans.data(), c.rdfa.trans.size() * 2);
  memcpy(ueoi, c.udfa.eoi.data(), c.udfa.eoi.size());
  memcpy(reoi, c.rdfa.eoi.data(), c.rdfa.eoi.size());
  return 1;
}
// the 64-byte per-call scratch {total, flag, t2, t3, ticket, t5, t6, t7}; the t* slots are phase
// cycle counters in the -DCGX_TIMING build (tools/phase_timing.py) and zero otherwise
int cgx_debug_scratch(cgx_regex* re, uint64_t out[8]) {
  if (!re || !re->d_ticket_total.p) return CGX_ERR_ARGS;
  re->diag = true;  // from now on the diagnostics are cleared before every launch
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out, re->d_ticket_total.p, 64, cudaMemcpyDeviceToHost));
  return CGX_OK;
}
int cgx_debug_dfa_info(const cgx_regex* re, int* nstates, uint16_t start[5], int* filter_kind,
                       int* skip_safe, int* kind_lut_needed, uint8_t ranges[8], int* nranges) {
  const Compiled& c = *re->c;
  *nstates = c.dfa.nstates;
  for (int k = 0; k < 5; k++) start[k] = c.dfa.start[k];
  *filter_kind = c.filter_kind;
  *skip_safe = c.skip_safe ? 1 : 0;
  *kind_lut_needed = c.kind_lut_needed ? 1 : 0;
  *nranges = c.nranges;
  for (int k = 0; k < 4; k++) {
    ranges[2 * k] = c.rlo[k];
    ranges[2 * k + 1] = c.rhi[k];
  }
  return c.kind == ENG_DFA ? 1 : 0;
}
int cgx_debug_dfa_copy(const cgx_regex* re, uint16_t* trans, uint8_t* eoi, uint8_t* lut) {
  const Compiled& c = *re->c;
  memcpy(trans, c.dfa.trans.data(), c.dfa.trans.size() * 2);
  memcpy(eoi, c.dfa.eoi.data(), c.dfa.eoi.size());
  memcpy(lut, c.lut, 256);
  return CGX_OK;
}

// flat start-filter program: ops[2*i]=kind, ops[2*i+1]=class; ranges[c][r] = lo,hi
int cgx_debug_flat(const cgx_regex* re, uint8_t* ops, int* nclasses, uint8_t* nranges4, uint8_t* ranges32) {
  const FlatDev& f = re->c->flat;
  for (int i = 0; i < f.nops; i++) {
    ops[2 * i] = f.op_kind[i];
    ops[2 * i + 1] = f.op_class[i];
  }
  *nclasses = f.nclasses;
  for (int c = 0; c < 4; c++) {
    nranges4[c] = f.cls_nranges[c];
    for (int r = 0; r < 4; r++) {
      ranges32[(c * 4 + r) * 2] = f.cls_lo[c][r];
      ranges32[(c * 4 + r) * 2 + 1] = f.cls_hi[c][r];
    }
  }
  return f.nops;
}

// ---- synthetic corpora ---------------------------------------------------------------------------
__global__ void synth_kernel(int kind, uint64_t seed, uint64_t first_block, uint8_t* out, uint64_t nblocks,
                             const uint8_t* lits, const int32_t* offs, int nlit) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nblocks) return;
  if (kind == 0) synth::gen_block_log(out + i * synth::kBlock01, seed, first_block + i);
  else if (kind == 1) synth::gen_block_text(out + i * synth::kBlock01, seed, first_block + i, lits, offs, nlit);
  else synth::gen_block_email(out + i * synth::kBlock2, seed, first_block + i);
}

static int synth_check(int kind, size_t len, size_t& bs) {
  if (kind < 0 || kind > 2) return CGX_ERR_ARGS;
  bs = kind == 2 ? synth::kBlock2 : synth::kBlock01;
  if (len % bs) {
    g_last_error = "synthetic corpus length must be a multiple of the block size";
    return CGX_ERR_ARGS;
  }
  return CGX_OK;
}

int cgx_synth_device(int kind, uint64_t seed, uint64_t first_block, uint8_t* d_out, size_t len,
                     const uint8_t* d_lits, const int32_t* d_offs, int nlit, void* stream) {
  size_t bs;
  int r = synth_check(kind, len, bs);
  if (r) return r;
  uint64_t nb = len / bs;
  if (!nb) return CGX_OK;
  unsigned threads = 128;
  uint64_t grid = (nb + threads - 1) / threads;
  synth_kernel<<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(kind, seed, first_block, d_out, nb,
                                                                      d_lits, d_offs, nlit);
  CU(cudaGetLastError());
  return CGX_OK;
}

int cgx_synth_host(int kind, uint64_t seed, uint64_t first_block, uint8_t* out, size_t len,
                   const uint8_t* lits, const int32_t* offs, int nlit) {
  size_t bs;
  int r = synth_check(kind, len, bs);
  if (r) return r;
  uint64_t nb = len / bs;
  for (uint64_t i = 0; i < nb; i++) {
    if (kind == 0) synth::gen_block_log(out + i * bs, seed, first_block + i);
    else if (kind == 1) synth::gen_block_text(out + i * bs, seed, first_block + i, lits, offs, nlit);
    else synth::gen_block_email(out + i * bs, seed, first_block + i);
  }
  return CGX_OK;
}

}  // extern "C"
