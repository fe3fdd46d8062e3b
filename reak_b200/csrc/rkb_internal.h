// rkb_internal.h — host-side glue between rkb_api.cu and the kernel translation units.
#ifndef RKB_INTERNAL_H
#define RKB_INTERNAL_H

#include <cuda_runtime.h>
#include "rkb_types.h"

// One entry per compiled (N, feature-mask) instantiation of the serial kernels.
struct SerialKernels {
  int n, fl;
  unsigned long long shape;     // structural promises compiled in (0 = none)
  int smem_eval, smem_rollout;  // dynamic shared memory per CTA, bytes
  int block;
  cudaError_t (*prepare)(void);  // opt in to > 48 KB dynamic shared memory
  cudaError_t (*eval)(const SerialParams&, const EvalArgs&, cudaStream_t);
  cudaError_t (*forces)(const SerialParams&, const EvalArgs&, cudaStream_t);
  cudaError_t (*mass)(const SerialParams&, const EvalArgs&, cudaStream_t);
  cudaError_t (*rollout)(const SerialParams&, const RolloutArgs&, cudaStream_t);                     // RK4
  cudaError_t (*rollout_rk)(const SerialParams&, const RolloutArgs&, const RkTable&, cudaStream_t);  // any explicit scheme
  cudaError_t (*rollout_seq)(const SerialParams&, const RolloutSeqArgs&, cudaStream_t);              // RK4, control sequence
  cudaError_t (*steer)(const SerialParams&, const SteerArgs&, cudaStream_t);                         // the whole steering loop
  // small batches: one sample on a pair of warps (kte_serial.cuh: DuoCtx); same results bit for bit
  cudaError_t (*rollout_duo)(const SerialParams&, const RolloutArgs&, cudaStream_t);
  cudaError_t (*rollout_seq_duo)(const SerialParams&, const RolloutSeqArgs&, cudaStream_t);
  cudaError_t (*steer_duo)(const SerialParams&, const SteerArgs&, cudaStream_t);
  cudaError_t (*rollout_scatter)(const SerialParams&, const RolloutScatterArgs&, cudaStream_t);  // rollout + all-gather by peer stores
  int (*rollout_ctas_per_sm)(void);  // occupancy of the RK4 rollout kernel on the current device
};

// defined in rkb_serial_n.cu, compiled once per N with -DRKB_N=<n>
#define RKB_DECL_TABLE(n) extern "C" const SerialKernels* rkb_serial_table_##n(int* count)
RKB_DECL_TABLE(1); RKB_DECL_TABLE(2); RKB_DECL_TABLE(3); RKB_DECL_TABLE(4);
RKB_DECL_TABLE(5); RKB_DECL_TABLE(6); RKB_DECL_TABLE(7); RKB_DECL_TABLE(8);

// generic interpreter (kte_generic.cu); `prog` is a device pointer to a GenericProgram
cudaError_t rkb_generic_eval(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s);
cudaError_t rkb_generic_forces(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s);
cudaError_t rkb_generic_frames(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s);
cudaError_t rkb_generic_tmt(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s);
cudaError_t rkb_generic_proximity(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, const ProxProgram& pp, cudaStream_t s);
cudaError_t rkb_generic_collisions(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, const ProxProgram& pp, int max_records,
                                   int32_t* finder, cudaStream_t s);
cudaError_t rkb_generic_frame_jac(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, int frame, unsigned upstream, cudaStream_t s);
cudaError_t rkb_generic_mass(const GenericProgram* prog, const GenericProgram& host, const EvalArgs& a, cudaStream_t s);
cudaError_t rkb_generic_rollout(const GenericProgram* prog, const GenericProgram& host, const RolloutArgs& a, const RkTable* table,
                                cudaStream_t s);  // table == NULL: the reference's RK4 arithmetic

// steer helpers (rkb_steer.cu): cost of every rollout end state against its pair's goal, then
// per-pair arg-min (first index wins ties) and gather of the winning end state.
cudaError_t rkb_steer_reduce(int nx, long long n_pairs, long long n_rollouts, const double* xend, const double* goal,
                             int32_t* best_idx, double* best_x, double* best_cost, cudaStream_t s);

// run-time specialised kernels (rkb_jit.cu): the same templates compiled with NVRTC for one (n, fl, shape)
enum { RKB_JIT_EVAL, RKB_JIT_FORCES, RKB_JIT_MASS, RKB_JIT_MASSDOT, RKB_JIT_ROLLOUT, RKB_JIT_ROLLOUT_RK, RKB_JIT_ROLLOUT_SEQ, RKB_JIT_STEER,
       RKB_JIT_COUNT };
struct JitKernels {
  int n, fl;
  unsigned long long shape;
  void* library;                      // cudaLibrary_t
  const void* kernel[RKB_JIT_COUNT];  // cudaKernel_t, launchable through cudaLaunchKernel
  int smem[RKB_JIT_COUNT];
};
int rkb_jit_get(int n, int fl, unsigned long long shape, const JitKernels** out);  // compiles on first use, cached per process
// non-blocking: *out == NULL while a background NVRTC compilation runs (started on the first call); a disk cache of cubins
// ($RKB_CACHE_DIR, $XDG_CACHE_HOME/reak_b200, ~/.cache/reak_b200; keyed by build id) serves later processes in milliseconds
int rkb_jit_poll(int n, int fl, unsigned long long shape, const JitKernels** out);
cudaError_t rkb_jit_prepare(const JitKernels& J);                                  // per-device kernel attributes (current device)
const char* rkb_jit_log();                                                         // NVRTC log / error text of the calling thread
// args: the kernel's second parameter (EvalArgs, RolloutArgs, ...); extra: its third (RkTable) or NULL
cudaError_t rkb_jit_launch(const JitKernels& J, int which, const SerialParams& P, const void* args, const void* extra, long long n_samples,
                           int smem_override, cudaStream_t s);

// kernels compiled at run time from a generated source (extern "C" names), keyed by a hash of the text; same disk cache
#define RKB_SRC_MAX_KERNELS 4
struct SourceKernels {
  void* library;                            // cudaLibrary_t
  const void* kernel[RKB_SRC_MAX_KERNELS];  // cudaKernel_t in the order of `names`
};
#ifdef __cplusplus
#include <string>
// lowered: `names` are C++ name expressions of templates to instantiate (else the extern "C" names themselves)
int rkb_jit_source_get(const char* prefix, const std::string& src, const char* const* names, int n_names, bool lowered, const SourceKernels** out);
int rkb_jit_source_poll(const char* prefix, const std::string& src, const char* const* names, int n_names, bool lowered, const SourceKernels** out);
cudaError_t rkb_jit_launch_source_rollout(const SourceKernels& K, int which, const SerialParams& P, const void* args, long long n_samples, int n,
                                          cudaStream_t s);
// rkb_prox_jit.cu: the source of the proximity kernels (rkb_prox_spec_d: distance and finder, rkb_prox_spec_p: with the two
// points) of one chain and one proxy pair; empty when the chain cannot be written as straight-line code
std::string rkb_prox_source(const GenericProgram& G, const ProxProgram& P, int min_blocks);
// the source of serial_steer_kernel<n, fl, shape, RkbSteerCheck>: the steering loop with the collision test of the pairs
// compiled in; *expr: the kernel's name expression
std::string rkb_steer_checked_source(int n, int fl, unsigned long long shape, const int* coord_of_stage, const GenericProgram& G,
                                     const ProxProgram* const* pairs, int n_pairs, std::string* expr);
#endif

// steering law between two control intervals (rkb_steer.cu)
cudaError_t rkb_steer_law(const SteerLawArgs& a, cudaStream_t s);
cudaError_t rkb_steer_commit(const SteerCommitArgs& a, cudaStream_t s);
cudaError_t rkb_free_combine(const double* dist, int n_pairs, long long n, int32_t* out, cudaStream_t s);


// rkb_linearize.cu: perturbed batch (rows (i, d, +-)) and the central differences of its evaluation
cudaError_t rkb_lin_perturb(long long n, int nx, int nu, double eps, const double* x, const double* u, double* xp, double* up, cudaStream_t s);
cudaError_t rkb_lin_combine(long long n, int nx, int nu, const double* xp, const double* up, const double* fd, const int32_t* st_in, double* A,
                            double* B, int32_t* status, cudaStream_t s);

#endif
