// rkb_serial_n.cu — instantiates the serial-chain kernels of kte_serial.cuh for one chain
// length (compile with -DRKB_N=<n>; the Makefile builds n = 1..8 in parallel).
#ifndef RKB_N
#error "compile with -DRKB_N=<number of coordinates>"
#endif
#include "kte_serial.cuh"
#include "rkb_internal.h"

namespace {

using namespace rkb;

template <int N, int FL>
struct Launch {
  static constexpr int kSmemEval = 6 * N * RKB_BLOCK * (int)sizeof(double);
  static constexpr int kSmemRollout = 12 * N * RKB_BLOCK * (int)sizeof(double);
  static unsigned grid(long long n) { return (unsigned)((n + RKB_BLOCK - 1) / RKB_BLOCK); }
  static cudaError_t prepare() {
    cudaError_t e = cudaFuncSetAttribute(serial_rollout_kernel<N, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemRollout);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_eval_kernel<N, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEval);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_forces_kernel<N, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEval);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(serial_mass_kernel<N, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEval);
  }
  static cudaError_t eval(const SerialParams& P, const EvalArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_eval_kernel<N, FL><<<grid(A.n_samples), RKB_BLOCK, kSmemEval, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t forces(const SerialParams& P, const EvalArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_forces_kernel<N, FL><<<grid(A.n_samples), RKB_BLOCK, kSmemEval, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t mass(const SerialParams& P, const EvalArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_mass_kernel<N, FL><<<grid(A.n_samples), RKB_BLOCK, kSmemEval, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t rollout(const SerialParams& P, const RolloutArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_rollout_kernel<N, FL><<<grid(A.n_samples), RKB_BLOCK, kSmemRollout, s>>>(P, A);
    return cudaGetLastError();
  }
  static SerialKernels entry() {
    SerialKernels k;
    k.n = N; k.fl = FL; k.smem_eval = kSmemEval; k.smem_rollout = kSmemRollout; k.block = RKB_BLOCK;
    k.prepare = &prepare; k.eval = &eval; k.forces = &forces; k.mass = &mass; k.rollout = &rollout;
    return k;
  }
};

}  // namespace

#define RKB_CAT2(a, b) a##b
#define RKB_CAT(a, b) RKB_CAT2(a, b)

extern "C" const SerialKernels* RKB_CAT(rkb_serial_table_, RKB_N)(int* count) {
  static const SerialKernels table[] = {
      Launch<RKB_N, 0>::entry(),
      Launch<RKB_N, RKB_FL_SPRINGS>::entry(),
      Launch<RKB_N, RKB_FL_ALL>::entry(),
  };
  *count = (int)(sizeof(table) / sizeof(table[0]));
  return table;
}
