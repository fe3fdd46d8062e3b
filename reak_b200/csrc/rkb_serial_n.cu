// rkb_serial_n.cu — instantiates the serial-chain kernels of kte_serial.cuh for one chain
// length (compile with -DRKB_N=<n>; the Makefile builds n = 1..8 in parallel).
#ifndef RKB_N
#error "compile with -DRKB_N=<number of coordinates>"
#endif
#include "kte_serial.cuh"
#include "rkb_internal.h"

namespace {

using namespace rkb;

template <int N, int FL, shape_t SHAPE>
struct Launch {
  static constexpr int kSmemEval = RKB_SMEM_EVAL_K(N) * RKB_BLOCK * (int)sizeof(double);
  static constexpr int kSmemForces = RKB_SMEM_FORCES_K(N) * RKB_BLOCK * (int)sizeof(double);
  static constexpr int kSmemMass = RKB_SMEM_MASS_K(N) * RKB_BLOCK * (int)sizeof(double);
  static constexpr int kSmemRollout = RKB_SMEM_ROLLOUT(N) * RKB_BLOCK * (int)sizeof(double);
  static constexpr int kSmemScatter = RKB_SMEM_ROLLOUT_K(N) * RKB_BLOCK * (int)sizeof(double);
  static constexpr int kSmemDuo = RKB_SMEM_DUO(N) * (int)sizeof(double);
  static unsigned grid_duo(long long n) { return (unsigned)((n + 63) / 64); }
  static constexpr int smem_rk(int stages) { return (2 * N + 2 * N * stages) * RKB_BLOCK * (int)sizeof(double); }
  static unsigned grid(long long n) { return (unsigned)((n + RKB_BLOCK - 1) / RKB_BLOCK); }
  static cudaError_t prepare() {
    cudaError_t e = cudaFuncSetAttribute(serial_rollout_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemRollout);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_rollout_rk_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_rk(RKB_RK_MAX_STAGES));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_rollout_seq_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemRollout);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_steer_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemRollout);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_rollout_scatter_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemScatter);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_rollout_duo_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDuo);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_rollout_seq_duo_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDuo);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_steer_duo_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDuo);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_eval_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEval);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_forces_kernel<N, FL, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemForces);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(serial_mass_kernel<N, FL, SHAPE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMass);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(serial_mass_kernel<N, FL, SHAPE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMass);
  }
  static cudaError_t eval(const SerialParams& P, const EvalArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_eval_kernel<N, FL, SHAPE><<<grid(A.n_samples), RKB_BLOCK, kSmemEval, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t forces(const SerialParams& P, const EvalArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_forces_kernel<N, FL, SHAPE><<<grid(A.n_samples), RKB_BLOCK, kSmemForces, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t mass(const SerialParams& P, const EvalArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    if (A.out2.p) serial_mass_kernel<N, FL, SHAPE, true><<<grid(A.n_samples), RKB_BLOCK, kSmemMass, s>>>(P, A);
    else serial_mass_kernel<N, FL, SHAPE, false><<<grid(A.n_samples), RKB_BLOCK, kSmemMass, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t rollout(const SerialParams& P, const RolloutArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_rollout_kernel<N, FL, SHAPE><<<grid(A.n_samples), RKB_BLOCK, kSmemRollout, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t rollout_rk(const SerialParams& P, const RolloutArgs& A, const RkTable& T, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_rollout_rk_kernel<N, FL, SHAPE><<<grid(A.n_samples), RKB_BLOCK, smem_rk(T.stages), s>>>(P, A, T);
    return cudaGetLastError();
  }
  static cudaError_t rollout_seq(const SerialParams& P, const RolloutSeqArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_rollout_seq_kernel<N, FL, SHAPE><<<grid(A.n_samples), RKB_BLOCK, kSmemRollout, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t steer(const SerialParams& P, const SteerArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_steer_kernel<N, FL, SHAPE><<<grid(A.n_samples), RKB_BLOCK, kSmemRollout, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t rollout_scatter(const SerialParams& P, const RolloutScatterArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_rollout_scatter_kernel<N, FL, SHAPE><<<grid(A.n_samples), RKB_BLOCK, kSmemScatter, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t rollout_duo(const SerialParams& P, const RolloutArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_rollout_duo_kernel<N, FL, SHAPE><<<grid_duo(A.n_samples), RKB_DUO_BLOCK, kSmemDuo, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t rollout_seq_duo(const SerialParams& P, const RolloutSeqArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_rollout_seq_duo_kernel<N, FL, SHAPE><<<grid_duo(A.n_samples), RKB_DUO_BLOCK, kSmemDuo, s>>>(P, A);
    return cudaGetLastError();
  }
  static cudaError_t steer_duo(const SerialParams& P, const SteerArgs& A, cudaStream_t s) {
    if (A.n_samples <= 0) return cudaSuccess;
    serial_steer_duo_kernel<N, FL, SHAPE><<<grid_duo(A.n_samples), RKB_DUO_BLOCK, kSmemDuo, s>>>(P, A);
    return cudaGetLastError();
  }
  // resident CTAs per SM of the RK4 rollout kernel on the current device (for sizing chunks in whole waves)
  static int rollout_ctas_per_sm() {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, serial_rollout_kernel<N, FL, SHAPE>, RKB_BLOCK, kSmemRollout) != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
  }
  static SerialKernels entry() {
    SerialKernels k;
    k.n = N; k.fl = FL; k.shape = SHAPE; k.smem_eval = kSmemEval; k.smem_rollout = kSmemRollout; k.block = RKB_BLOCK;
    k.prepare = &prepare; k.eval = &eval; k.forces = &forces; k.mass = &mass; k.rollout = &rollout; k.rollout_rk = &rollout_rk; k.rollout_seq = &rollout_seq; k.steer = &steer;
    k.rollout_ctas_per_sm = &rollout_ctas_per_sm; k.rollout_scatter = &rollout_scatter;
    k.rollout_duo = &rollout_duo; k.rollout_seq_duo = &rollout_seq_duo; k.steer_duo = &steer_duo;
    return k;
  }
};

}  // namespace

#define RKB_CAT2(a, b) a##b
#define RKB_CAT(a, b) RKB_CAT2(a, b)

// Structural shapes compiled ahead of time besides the general code (shape 0).  The lowering
// (rkb_api.cu) picks the most specialised entry whose promises the chain keeps.
//   Z/Y/X: revolute about that axis, link offset along z (or none), diagonal inertia tensor —
//   the anthropomorphic-arm-with-spherical-wrist pattern z,y,y,z,y,z of the CRS A465 preset
//   (examples/robot_airship/old/CRS_A465_models.cpp:304-640) and its prefixes.
//   `signed_axes` also promises the directions z, -y, -y, z, -y, z of that preset (and of manip_3R3R_arm.cpp:104-212):
//   every multiplication by the axis sign folds away, 7 % of the FP64 instructions of an evaluation.
constexpr int kArmAxis[6] = {3, 2, 2, 3, 2, 3};
constexpr int kArmSign[6] = {1, 3, 3, 1, 3, 1};
constexpr shape_t arm_shape(int n, int first, int inertia = 1, bool signed_axes = false) {
  shape_t s = 0;
  for (int k = 0; k < n; ++k)
    s |= RKB_SHAPE_AT(RKB_SHAPE_STAGE_SIGNED(kArmAxis[k % 6], 3, inertia, signed_axes ? kArmSign[k % 6] : 0), k + first);
  return s;
}
// general joint (prismatic track) with a z-aligned link and a diagonal tensor, then the arm
// (signed: the track runs along +e_x, CRS_A465_models.cpp:304-347)
constexpr shape_t track_arm_shape(int n, bool signed_axes = false) {
  return RKB_SHAPE_AT(signed_axes ? RKB_SHAPE_STAGE_SIGNED(5, 3, 1, 1) : RKB_SHAPE_STAGE(0, 3, 1), 0) | arm_shape(n - 1, 1, 1, signed_axes);
}

// planar chains embedded in the x-y plane (rkb_api.cu: embed_planar): every joint about z, links along x,
// diagonal tensors — cfg 1 and the 2D analog of the CRS arm (examples/robot_airship/old/CRS_A465_2D_analog.cpp);
// with the prismatic track of that model in front, whose link has no offset
constexpr shape_t planar_shape(int n, int first = 0) {
  shape_t s = 0;
  for (int k = 0; k < n; ++k) s |= RKB_SHAPE_AT(RKB_SHAPE_STAGE_SIGNED(3, 1, 1, 1), k + first);  // revolute_joint_2D turns about +e_z
  return s;
}
constexpr shape_t track_planar_shape(int n) { return RKB_SHAPE_AT(RKB_SHAPE_STAGE_SIGNED(5, 3, 1, 1), 0) | planar_shape(n - 1, 1); }

extern "C" const SerialKernels* RKB_CAT(rkb_serial_table_, RKB_N)(int* count) {
  static const SerialKernels table[] = {
      Launch<RKB_N, 0, 0>::entry(),
      Launch<RKB_N, RKB_FL_SPRINGS, 0>::entry(),
      Launch<RKB_N, RKB_FL_ALL, 0>::entry(),
      Launch<RKB_N, 0, arm_shape(RKB_N, 0, 1, true)>::entry(),               // the CRS arm as the reference builds it
      Launch<RKB_N, RKB_FL_SPRINGS, arm_shape(RKB_N, 0, 1, true)>::entry(),
      Launch<RKB_N, RKB_FL_SPRINGS, arm_shape(RKB_N, 0)>::entry(),           // same axes, any directions
      Launch<RKB_N, RKB_FL_SPRINGS, arm_shape(RKB_N, 0, 0)>::entry(),        // ... and full inertia tensors
#if RKB_N >= 2
      Launch<RKB_N, RKB_FL_PRISMATIC, track_arm_shape(RKB_N, true)>::entry(),
      Launch<RKB_N, RKB_FL_PRISMATIC | RKB_FL_SPRINGS, track_arm_shape(RKB_N)>::entry(),
#endif
#if RKB_N <= 4
      Launch<RKB_N, RKB_FL_SPRINGS, planar_shape(RKB_N)>::entry(),
#if RKB_N >= 2
      Launch<RKB_N, RKB_FL_PRISMATIC | RKB_FL_SPRINGS, track_planar_shape(RKB_N)>::entry(),
#endif
#endif
  };
  *count = (int)(sizeof(table) / sizeof(table[0]));
  return table;
}
