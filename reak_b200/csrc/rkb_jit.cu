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

#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "kte_serial.cuh"
#include "rkb_internal.h"
#include "jit_src.inc"
#include "build_id.h"

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
std::mutex g_mu, g_nvrtc_mu;
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

namespace {

typedef std::tuple<int, int, unsigned long long> Key;

// What NVRTC produced for one (n, fl, shape): the cubin and the lowered (mangled) name of every kernel in it.
struct Cubin {
  std::vector<char> image;
  std::vector<std::string> names;  // RKB_JIT_COUNT entries
};

// ---- disk cache -----------------------------------------------------------------------------------------
// $RKB_CACHE_DIR, else $XDG_CACHE_HOME/reak_b200, else $HOME/.cache/reak_b200.  One file per (build, n, fl, shape):
// a header line per kernel name, then the cubin.  A cubin is only valid for the sources it was compiled from, so the
// file name carries rkb_build_id().  Failures to read or write are not errors: the kernels are compiled instead.
std::string cache_path(const std::string& stem) {
  std::string dir;
  if (const char* d = std::getenv("RKB_CACHE_DIR")) dir = d;
  else if (const char* x = std::getenv("XDG_CACHE_HOME")) dir = std::string(x) + "/reak_b200";
  else if (const char* h = std::getenv("HOME")) dir = std::string(h) + "/.cache/reak_b200";
  else return std::string();
  if (dir.empty() || dir == "off") return std::string();
  std::string acc;
  for (size_t i = 0; i <= dir.size(); ++i)  // mkdir -p
    if (i == dir.size() || (dir[i] == '/' && i > 0)) { acc = dir.substr(0, i); ::mkdir(acc.c_str(), 0755); }
  return dir + "/" + RKB_BUILD_ID + "-" + stem + ".cubin";
}
std::string stem_of(const Key& key) {
  char buf[96];
  std::snprintf(buf, sizeof buf, "n%d-fl%d-%016llx", std::get<0>(key), std::get<1>(key), std::get<2>(key));
  return buf;
}
bool cache_read(const std::string& stem, int n_names, Cubin& out) {
  const std::string path = cache_path(stem);
  if (path.empty()) return false;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  bool ok = true;
  out.names.clear();
  char line[1024];
  unsigned long long size = 0;
  if (!std::fgets(line, sizeof line, f) || std::sscanf(line, "RKBCUBIN %llu", &size) != 1 || size == 0 || size > (1ull << 28)) ok = false;
  for (int k = 0; ok && k < n_names; ++k) {
    if (!std::fgets(line, sizeof line, f)) { ok = false; break; }
    std::string n(line);
    while (!n.empty() && (n.back() == '\n' || n.back() == '\r')) n.pop_back();
    if (n.empty()) ok = false;
    out.names.push_back(n);
  }
  if (ok) {
    out.image.resize(size);
    ok = std::fread(out.image.data(), 1, size, f) == size;
  }
  std::fclose(f);
  return ok;
}
void cache_write(const std::string& stem, const Cubin& c) {
  const std::string path = cache_path(stem);
  if (path.empty()) return;
  char tmp[32];
  std::snprintf(tmp, sizeof tmp, ".%d.tmp", (int)::getpid());
  const std::string t = path + tmp;
  FILE* f = std::fopen(t.c_str(), "wb");
  if (!f) return;
  std::fprintf(f, "RKBCUBIN %llu\n", (unsigned long long)c.image.size());
  for (const auto& n : c.names) std::fprintf(f, "%s\n", n.c_str());
  const bool ok = std::fwrite(c.image.data(), 1, c.image.size(), f) == c.image.size();
  std::fclose(f);
  if (ok) std::rename(t.c_str(), path.c_str());  // atomic: concurrent processes never see half a file
  else std::remove(t.c_str());
}

// ---- NVRTC: host-only, needs no CUDA context (may run on a helper thread) ----------------------------------
// One compilation: the program text, the kernels wanted from it (C++ name expressions when `lowered`, else the
// extern "C" names themselves) and the file stem of its disk-cache entry.
struct Job {
  std::string stem, src;
  std::vector<std::string> exprs;
  bool lowered = true;
};
Job chain_job(const Key& key) {
  const int n = std::get<0>(key), fl = std::get<1>(key);
  const unsigned long long shape = std::get<2>(key);
  Job j;
  j.stem = stem_of(key);
  j.src = "#include \"kte_serial.cuh\"\n";
  j.exprs.resize(RKB_JIT_COUNT);
  char buf[256];
  for (int k = 0; k < RKB_JIT_COUNT; ++k) {
    if (k == RKB_JIT_MASS || k == RKB_JIT_MASSDOT)
      std::snprintf(buf, sizeof buf, "rkb::%s<%d, %d, %lluull, %s>", kKernelNames[k], n, fl, shape, k == RKB_JIT_MASSDOT ? "true" : "false");
    else
      std::snprintf(buf, sizeof buf, "rkb::%s<%d, %d, %lluull>", kKernelNames[k], n, fl, shape);
    j.exprs[k] = buf;
  }
  return j;
}
int compile_cubin(const Job& job, Cubin& out, std::string& log) {
  {
    std::lock_guard<std::mutex> lock(g_nvrtc_mu);
    if (!g_nvrtc.load()) { log = "libnvrtc.so.12 not found"; return RKB_ERR_UNSUPPORTED; }
  }
  std::vector<const char*> hn, ht;
  for (const auto& h : kJitHeaders) { hn.push_back(h.name); ht.push_back(h.text); }
  static const char kStdint[] = "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
                                "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n";
  const char* stubs[][2] = {{"cuda_runtime.h", ""}, {"math.h", ""}, {"stddef.h", ""}, {"stdint.h", kStdint}};
  for (auto& s : stubs) { hn.push_back(s[0]); ht.push_back(s[1]); }
  nvrtcProgram prog = nullptr;
  if (g_nvrtc.CreateProgram(&prog, job.src.c_str(), "rkb_jit.cu", (int)hn.size(), ht.data(), hn.data()) != 0) { log = "nvrtcCreateProgram failed"; return RKB_ERR_CUDA; }
  if (job.lowered)
    for (auto& e : job.exprs) g_nvrtc.AddNameExpression(prog, e.c_str());
  // -default-device: the C-ABI prototypes of reak_b200.h carry no execution-space annotation
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
  const int rc = g_nvrtc.CompileProgram(prog, 4, opts);
  size_t ls = 0;
  g_nvrtc.GetProgramLogSize(prog, &ls);
  if (ls > 1) { log.resize(ls); g_nvrtc.GetProgramLog(prog, &log[0]); }
  if (rc != 0) { g_nvrtc.DestroyProgram(&prog); return RKB_ERR_CUDA; }
  size_t cs = 0;
  g_nvrtc.GetCUBINSize(prog, &cs);
  out.image.resize(cs);
  g_nvrtc.GetCUBIN(prog, out.image.data());
  out.names.clear();
  for (const auto& e : job.exprs) {
    if (!job.lowered) { out.names.push_back(e); continue; }
    const char* low = nullptr;
    if (g_nvrtc.GetLoweredName(prog, e.c_str(), &low) != 0 || !low) { log = "nvrtcGetLoweredName failed for " + e; g_nvrtc.DestroyProgram(&prog); return RKB_ERR_CUDA; }
    out.names.push_back(low);
  }
  g_nvrtc.DestroyProgram(&prog);
  return RKB_OK;
}

// ---- loading: needs the CUDA context of the calling thread --------------------------------------------------
int load_cubin(const Key& key, const Cubin& c, JitKernels** out, std::string& log) {
  const int n = std::get<0>(key);
  JitKernels* J = new JitKernels();
  J->n = n; J->fl = std::get<1>(key); J->shape = std::get<2>(key);
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, c.image.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) { log = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); cudaGetLastError(); delete J; return RKB_ERR_CUDA; }
  J->library = lib;
  for (int k = 0; k < RKB_JIT_COUNT; ++k) {
    cudaKernel_t kern = nullptr;
    e = cudaLibraryGetKernel(&kern, lib, c.names[k].c_str());
    if (e != cudaSuccess) { log = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e); cudaGetLastError(); cudaLibraryUnload(lib); delete J; return RKB_ERR_CUDA; }
    J->kernel[k] = (const void*)kern;
  }
  // dynamic shared memory per CTA, as rkb_serial_n.cu sizes it
  const int B = RKB_BLOCK * (int)sizeof(double);
  J->smem[RKB_JIT_EVAL] = (1 > 2 * n + 1 ? 1 : 2 * n + 1) * B;
  J->smem[RKB_JIT_FORCES] = (n + 1) * B;
  J->smem[RKB_JIT_MASS] = J->smem[RKB_JIT_MASSDOT] = (n * n + 1) * B;
  J->smem[RKB_JIT_ROLLOUT] = J->smem[RKB_JIT_ROLLOUT_SEQ] = J->smem[RKB_JIT_STEER] = RKB_SMEM_ROLLOUT(n) * B;
  J->smem[RKB_JIT_ROLLOUT_RK] = (2 * n + 2 * n * RKB_RK_MAX_STAGES) * B;
  e = rkb_jit_prepare(*J);
  if (e != cudaSuccess) { log = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); cudaGetLastError(); delete J; return RKB_ERR_CUDA; }
  *out = J;
  return RKB_OK;
}

// background compilations: key -> state
struct Pending {
  int state = 0;  // 0 compiling, 1 ready, -1 failed
  Cubin cubin;
  std::string log;
};
std::map<Key, std::shared_ptr<Pending> > g_pending;
std::vector<std::thread> g_threads;
bool g_atexit = false;
void join_all() {
  std::vector<std::thread> t;
  { std::lock_guard<std::mutex> lock(g_mu); t.swap(g_threads); }
  for (auto& x : t) if (x.joinable()) x.join();
}

}  // namespace

// Synchronous: disk cache, else NVRTC now.  `from_cache` (nullable) tells which.
int rkb_jit_get(int n, int fl, unsigned long long shape, const JitKernels** out) {
  const Key key = std::make_tuple(n, fl, shape);
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { *out = it->second; return RKB_OK; }
  }
  g_log.clear();
  Cubin c;
  const Job job = chain_job(key);
  bool cached = cache_read(job.stem, RKB_JIT_COUNT, c);
  if (!cached) {
    const int rc = compile_cubin(job, c, g_log);
    if (rc) return rc;
    cache_write(job.stem, c);
  }
  JitKernels* J = nullptr;
  int rc = load_cubin(key, c, &J, g_log);
  if (rc && cached) {  // a stale or damaged cache file: compile afresh
    if ((rc = compile_cubin(job, c, g_log))) return rc;
    cache_write(job.stem, c);
    rc = load_cubin(key, c, &J, g_log);
  }
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) { *out = it->second; return RKB_OK; }  // (another thread was faster; the spare library stays loaded)
  g_cache[key] = J;
  *out = J;
  return RKB_OK;
}

// Asynchronous: returns RKB_OK with *out set when the kernels are available (loaded before, in the disk cache, or a
// background compilation has finished: they are loaded into the CALLER's context), RKB_OK with *out == NULL while a
// background compilation is running (it is started on the first call), an error when it failed.
int rkb_jit_poll(int n, int fl, unsigned long long shape, const JitKernels** out) {
  *out = nullptr;
  const Key key = std::make_tuple(n, fl, shape);
  std::shared_ptr<Pending> p;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { *out = it->second; return RKB_OK; }
    auto pit = g_pending.find(key);
    if (pit != g_pending.end()) p = pit->second;
  }
  const Cubin* ready = nullptr;
  if (!p) {
    Cubin from_disk;
    if (cache_read(stem_of(key), RKB_JIT_COUNT, from_disk)) {
      JitKernels* Jd = nullptr;
      if (load_cubin(key, from_disk, &Jd, g_log) == RKB_OK) {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) { *out = it->second; return RKB_OK; }
        g_cache[key] = Jd;
        *out = Jd;
        return RKB_OK;
      }
      std::remove(cache_path(stem_of(key)).c_str());  // stale or damaged: not trusted, compiled afresh below
    }
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_pending.find(key) == g_pending.end()) {
      p = std::make_shared<Pending>();
      g_pending[key] = p;
      if (!g_atexit) { std::atexit(join_all); g_atexit = true; }
      g_threads.emplace_back([key, p]() {
        Cubin c;
        std::string log;
        const Job job = chain_job(key);
        const int rc = compile_cubin(job, c, log);
        if (rc == RKB_OK) cache_write(job.stem, c);
        std::lock_guard<std::mutex> lock2(g_mu);
        p->cubin.image.swap(c.image);
        p->cubin.names.swap(c.names);
        p->log = log;
        p->state = rc == RKB_OK ? 1 : -1;
      });
    }
    return RKB_OK;
  } else {
    std::lock_guard<std::mutex> lock(g_mu);
    if (p->state == 0) return RKB_OK;
    if (p->state < 0) { g_log = p->log; return RKB_ERR_CUDA; }
    ready = &p->cubin;
  }
  JitKernels* J = nullptr;
  const int rc = load_cubin(key, *ready, &J, g_log);
  if (rc) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto np = std::make_shared<Pending>();
    np->state = -1; np->log = g_log;
    g_pending[key] = np;  // do not try again
    return rc;
  }
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) { *out = it->second; return RKB_OK; }
  g_cache[key] = J;
  *out = J;
  return RKB_OK;
}

// ---- kernels from a generated source (rkb_prox_jit.cu): extern "C" kernels, cached by a hash of the text ---------
namespace {
struct SrcPending {
  int state = 0;  // 0 compiling, 1 ready, -1 failed
  Cubin cubin;
  std::string log;
};
std::map<std::string, SourceKernels*> g_src_cache;
std::map<std::string, std::shared_ptr<SrcPending> > g_src_pending;

std::string source_stem(const char* prefix, const std::string& src) {
  unsigned long long h = 1469598103934665603ull;  // FNV-1a
  for (unsigned char ch : src) { h ^= ch; h *= 1099511628211ull; }
  char buf[96];
  std::snprintf(buf, sizeof buf, "%s-%016llx", prefix, h);
  return buf;
}
Job source_job(const std::string& stem, const std::string& src, const char* const* names, int n_names, bool lowered) {
  Job j;
  j.stem = stem;
  j.src = src;
  j.lowered = lowered;
  for (int k = 0; k < n_names; ++k) j.exprs.push_back(names[k]);
  return j;
}
int load_source(const Cubin& c, int n_names, SourceKernels** out, std::string& log) {
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, c.image.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) { log = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); cudaGetLastError(); return RKB_ERR_CUDA; }
  SourceKernels* S = new SourceKernels();
  S->library = lib;
  for (int k = 0; k < n_names && k < RKB_SRC_MAX_KERNELS; ++k) {
    cudaKernel_t kern = nullptr;
    e = cudaLibraryGetKernel(&kern, lib, c.names[k].c_str());
    if (e != cudaSuccess) { log = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e); cudaGetLastError(); cudaLibraryUnload(lib); delete S; return RKB_ERR_CUDA; }
    S->kernel[k] = (const void*)kern;
  }
  *out = S;
  return RKB_OK;
}
const SourceKernels* src_publish(const std::string& stem, SourceKernels* S) {  // g_mu held
  auto it = g_src_cache.find(stem);
  if (it != g_src_cache.end()) return it->second;  // (another thread was faster; the spare library stays loaded)
  g_src_cache[stem] = S;
  return S;
}
}  // namespace

// Synchronous: this process's table, else the disk cache, else NVRTC now.
int rkb_jit_source_get(const char* prefix, const std::string& src, const char* const* names, int n_names, bool lowered, const SourceKernels** out) {
  if (n_names < 1 || n_names > RKB_SRC_MAX_KERNELS) return RKB_ERR_INVALID;
  const std::string stem = source_stem(prefix, src);
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_src_cache.find(stem);
    if (it != g_src_cache.end()) { *out = it->second; return RKB_OK; }
  }
  g_log.clear();
  const Job job = source_job(stem, src, names, n_names, lowered);
  Cubin c;
  bool cached = cache_read(stem, n_names, c);
  if (!cached) {
    const int rc = compile_cubin(job, c, g_log);
    if (rc) return rc;
    cache_write(stem, c);
  }
  SourceKernels* S = nullptr;
  int rc = load_source(c, n_names, &S, g_log);
  if (rc && cached) {  // a stale or damaged cache file: compile afresh
    if ((rc = compile_cubin(job, c, g_log))) return rc;
    cache_write(stem, c);
    rc = load_source(c, n_names, &S, g_log);
  }
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_mu);
  *out = src_publish(stem, S);
  return RKB_OK;
}

// Asynchronous, like rkb_jit_poll: *out == NULL while a background compilation (started by the first call) runs.
int rkb_jit_source_poll(const char* prefix, const std::string& src, const char* const* names, int n_names, bool lowered, const SourceKernels** out) {
  *out = nullptr;
  if (n_names < 1 || n_names > RKB_SRC_MAX_KERNELS) return RKB_ERR_INVALID;
  const std::string stem = source_stem(prefix, src);
  std::shared_ptr<SrcPending> p;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_src_cache.find(stem);
    if (it != g_src_cache.end()) { *out = it->second; return RKB_OK; }
    auto pit = g_src_pending.find(stem);
    if (pit != g_src_pending.end()) p = pit->second;
  }
  if (!p) {
    Cubin from_disk;
    if (cache_read(stem, n_names, from_disk)) {
      SourceKernels* S = nullptr;
      if (load_source(from_disk, n_names, &S, g_log) == RKB_OK) {
        std::lock_guard<std::mutex> lock(g_mu);
        *out = src_publish(stem, S);
        return RKB_OK;
      }
      std::remove(cache_path(stem).c_str());
    }
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_src_pending.find(stem) == g_src_pending.end()) {
      p = std::make_shared<SrcPending>();
      g_src_pending[stem] = p;
      if (!g_atexit) { std::atexit(join_all); g_atexit = true; }
      const Job job = source_job(stem, src, names, n_names, lowered);
      g_threads.emplace_back([job, p]() {
        Cubin c;
        std::string log;
        const int rc = compile_cubin(job, c, log);
        if (rc == RKB_OK) cache_write(job.stem, c);
        std::lock_guard<std::mutex> lock2(g_mu);
        p->cubin.image.swap(c.image);
        p->cubin.names.swap(c.names);
        p->log = log;
        p->state = rc == RKB_OK ? 1 : -1;
      });
    }
    return RKB_OK;
  }
  {
    std::lock_guard<std::mutex> lock(g_mu);
    if (p->state == 0) return RKB_OK;
    if (p->state < 0) { g_log = p->log; return RKB_ERR_CUDA; }
  }
  SourceKernels* S = nullptr;
  const int rc = load_source(p->cubin, n_names, &S, g_log);
  std::lock_guard<std::mutex> lock(g_mu);
  if (rc) {
    auto np = std::make_shared<SrcPending>();
    np->state = -1; np->log = g_log;
    g_src_pending[stem] = np;  // do not try again
    return rc;
  }
  *out = src_publish(stem, S);
  return RKB_OK;
}

// a generated serial-chain kernel taking (SerialParams, args) with the rollout kernels' shared-memory layout
cudaError_t rkb_jit_launch_source_rollout(const SourceKernels& K, int which, const SerialParams& P, const void* args, long long n_samples, int n,
                                          cudaStream_t s) {
  if (n_samples <= 0) return cudaSuccess;
  const int smem = RKB_SMEM_ROLLOUT(n) * RKB_BLOCK * (int)sizeof(double);
  // (per device and cheap: the opt-in to more than 48 KB of dynamic shared memory)
  const cudaError_t e = cudaFuncSetAttribute(K.kernel[which], cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  void* argv[2] = {const_cast<SerialParams*>(&P), const_cast<void*>(args)};
  const unsigned grid = (unsigned)((n_samples + RKB_BLOCK - 1) / RKB_BLOCK);
  return cudaLaunchKernel(K.kernel[which], dim3(grid), dim3(RKB_BLOCK), argv, smem, s);
}

cudaError_t rkb_jit_launch(const JitKernels& J, int which, const SerialParams& P, const void* args, const void* extra, long long n_samples,
                           int smem_override, cudaStream_t s) {
  if (n_samples <= 0) return cudaSuccess;
  void* argv[3] = {const_cast<SerialParams*>(&P), const_cast<void*>(args), const_cast<void*>(extra)};
  const unsigned grid = (unsigned)((n_samples + RKB_BLOCK - 1) / RKB_BLOCK);
  return cudaLaunchKernel(J.kernel[which], dim3(grid), dim3(RKB_BLOCK), argv, smem_override > 0 ? smem_override : J.smem[which], s);
}
