// rkb_jit.cu — run-time specialisation of the serial-chain kernels.
//
// The ahead-of-time table (rkb_serial_n.cu) holds the general code and the shapes of the reference's own
// models; any other chain whose joints are axis-aligned, whose links run along one axis or whose tensors
// are diagonal would fall back to the general code at half the speed.  rkb_chain_specialize compiles
// kte_serial.cuh for exactly the chain's (N, feature mask, SHAPE) with NVRTC — the same source, the same
// templates, ~1 s per kernel — loads the cubin through the runtime's library API and routes the handle's
// launches to it.  NVRTC is dlopen-ed on first use: the library has no load-time dependency on it and the
// ahead-of-time kernels keep working where it is absent.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "kte_serial.cuh"
#include "rkb_internal.h"
#include "jit_src.inc"

namespace {

typedef void* nvrtcProgram;
struct Nvrtc {
  void* so = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
  int (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  bool load() {
    if (so) return true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names) { so = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (so) break; }
    if (!so) return false;
#define RKB_SYM(f) f = reinterpret_cast<decltype(f)>(dlsym(so, "nvrtc" #f)); if (!f) { so = nullptr; return false; }
    RKB_SYM(CreateProgram) RKB_SYM(DestroyProgram) RKB_SYM(CompileProgram) RKB_SYM(GetCUBINSize) RKB_SYM(GetCUBIN)
    RKB_SYM(GetProgramLogSize) RKB_SYM(GetProgramLog) RKB_SYM(AddNameExpression) RKB_SYM(GetLoweredName)
#undef RKB_SYM
    return true;
  }
};

Nvrtc g_nvrtc;
std::mutex g_mu;
std::map<std::tuple<int, int, unsigned long long>, JitKernels*> g_cache;
thread_local std::string g_log;

const char* kKernelNames[RKB_JIT_COUNT] = {"serial_eval_kernel", "serial_forces_kernel", "serial_mass_kernel", "serial_mass_kernel",
                                           "serial_rollout_kernel", "serial_rollout_rk_kernel", "serial_rollout_seq_kernel", "serial_steer_kernel"};

}  // namespace

const char* rkb_jit_log() { return g_log.c_str(); }

// the opt-in to more than 48 KB of dynamic shared memory is per device: call with the device current
cudaError_t rkb_jit_prepare(const JitKernels& J) {
  for (int k = 0; k < RKB_JIT_COUNT; ++k) {
    const cudaError_t e = cudaFuncSetAttribute(J.kernel[k], cudaFuncAttributeMaxDynamicSharedMemorySize, J.smem[k]);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

int rkb_jit_get(int n, int fl, unsigned long long shape, const JitKernels** out) {
  std::lock_guard<std::mutex> lock(g_mu);
  const auto key = std::make_tuple(n, fl, shape);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) { *out = it->second; return RKB_OK; }
  g_log.clear();
  if (!g_nvrtc.load()) { g_log = "libnvrtc.so.12 not found"; return RKB_ERR_UNSUPPORTED; }
  // program: the kernel header plus one name expression per kernel
  std::string src = "#include \"kte_serial.cuh\"\n";
  std::vector<std::string> exprs(RKB_JIT_COUNT);
  char buf[256];
  for (int k = 0; k < RKB_JIT_COUNT; ++k) {
    if (k == RKB_JIT_MASS || k == RKB_JIT_MASSDOT)
      std::snprintf(buf, sizeof buf, "rkb::%s<%d, %d, %lluull, %s>", kKernelNames[k], n, fl, shape, k == RKB_JIT_MASSDOT ? "true" : "false");
    else
      std::snprintf(buf, sizeof buf, "rkb::%s<%d, %d, %lluull>", kKernelNames[k], n, fl, shape);
    exprs[k] = buf;
  }
  std::vector<const char*> hn, ht;
  for (const auto& h : kJitHeaders) { hn.push_back(h.name); ht.push_back(h.text); }
  static const char kStdint[] = "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
                                "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n";
  const char* stubs[][2] = {{"cuda_runtime.h", ""}, {"math.h", ""}, {"stddef.h", ""}, {"stdint.h", kStdint}};
  for (auto& s : stubs) { hn.push_back(s[0]); ht.push_back(s[1]); }
  nvrtcProgram prog = nullptr;
  if (g_nvrtc.CreateProgram(&prog, src.c_str(), "rkb_jit.cu", (int)hn.size(), ht.data(), hn.data()) != 0) { g_log = "nvrtcCreateProgram failed"; return RKB_ERR_CUDA; }
  for (auto& e : exprs) g_nvrtc.AddNameExpression(prog, e.c_str());
  // -default-device: the C-ABI prototypes of reak_b200.h carry no execution-space annotation
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
  const int rc = g_nvrtc.CompileProgram(prog, 4, opts);
  size_t ls = 0;
  g_nvrtc.GetProgramLogSize(prog, &ls);
  if (ls > 1) { g_log.resize(ls); g_nvrtc.GetProgramLog(prog, &g_log[0]); }
  if (rc != 0) { g_nvrtc.DestroyProgram(&prog); return RKB_ERR_CUDA; }
  size_t cs = 0;
  g_nvrtc.GetCUBINSize(prog, &cs);
  std::vector<char> cubin(cs);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  JitKernels* J = new JitKernels();
  J->n = n; J->fl = fl; J->shape = shape;
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) { g_log = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); cudaGetLastError(); g_nvrtc.DestroyProgram(&prog); delete J; return RKB_ERR_CUDA; }
  J->library = lib;
  for (int k = 0; k < RKB_JIT_COUNT; ++k) {
    const char* low = nullptr;
    if (g_nvrtc.GetLoweredName(prog, exprs[k].c_str(), &low) != 0 || !low) { g_log = "nvrtcGetLoweredName failed for " + exprs[k]; g_nvrtc.DestroyProgram(&prog); delete J; return RKB_ERR_CUDA; }
    cudaKernel_t kern = nullptr;
    e = cudaLibraryGetKernel(&kern, lib, low);
    if (e != cudaSuccess) { g_log = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e); cudaGetLastError(); g_nvrtc.DestroyProgram(&prog); delete J; return RKB_ERR_CUDA; }
    J->kernel[k] = (const void*)kern;
  }
  g_nvrtc.DestroyProgram(&prog);
  // dynamic shared memory per CTA, as rkb_serial_n.cu sizes it
  const int B = RKB_BLOCK * (int)sizeof(double);
  J->smem[RKB_JIT_EVAL] = (1 > 2 * n + 1 ? 1 : 2 * n + 1) * B;
  J->smem[RKB_JIT_FORCES] = (n + 1) * B;
  J->smem[RKB_JIT_MASS] = J->smem[RKB_JIT_MASSDOT] = (n * n + 1) * B;
  J->smem[RKB_JIT_ROLLOUT] = J->smem[RKB_JIT_ROLLOUT_SEQ] = J->smem[RKB_JIT_STEER] = RKB_SMEM_ROLLOUT(n) * B;
  J->smem[RKB_JIT_ROLLOUT_RK] = (2 * n + 2 * n * RKB_RK_MAX_STAGES) * B;
  e = rkb_jit_prepare(*J);
  if (e != cudaSuccess) { g_log = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); cudaGetLastError(); delete J; return RKB_ERR_CUDA; }
  g_cache[key] = J;
  *out = J;
  return RKB_OK;
}

cudaError_t rkb_jit_launch(const JitKernels& J, int which, const SerialParams& P, const void* args, const void* extra, long long n_samples,
                           int smem_override, cudaStream_t s) {
  if (n_samples <= 0) return cudaSuccess;
  void* argv[3] = {const_cast<SerialParams*>(&P), const_cast<void*>(args), const_cast<void*>(extra)};
  const unsigned grid = (unsigned)((n_samples + RKB_BLOCK - 1) / RKB_BLOCK);
  return cudaLaunchKernel(J.kernel[which], dim3(grid), dim3(RKB_BLOCK), argv, smem_override > 0 ? smem_override : J.smem[which], s);
}
