// Test infrastructure (never shipped, never loaded by the product): compiles the DEVICE source of the interpreter
// kernels, reak_b200/csrc/kte_generic.cu, for the host — each kernel body becomes an ordinary function that processes
// "thread 0" — so that the CPU suite can run the interpreter's chain walk, mass matrix, twist-shaping matrices and
// integrators against the oracle without a GPU.  Built by tests/test_generic_host.py into a temporary directory (g++,
// CUDA headers for the types only).  The program is the GenericProgram rkb_chain_create lowered (rkb_chain_program, a
// test hook of the product library).
#define RKB_HOST_TEST 1
#include <cuda_runtime.h>
#define __launch_bounds__(...)
#ifndef __grid_constant__
#define __grid_constant__
#endif
const uint3 threadIdx = {0, 0, 0};
const uint3 blockIdx = {0, 0, 0};
const dim3 blockDim = dim3(1, 1, 1);
#include "../../reak_b200/csrc/kte_generic.cu"

namespace {
EvalArgs one_sample(const double* x, int nx, const double* u, int nu, double* out, int out_dim, double* out2, int32_t* status) {
  EvalArgs A;
  A.x = ConstBatchView{x, nx, 1, 0};
  A.u = ConstBatchView{u, nu > 0 ? nu : 1, 1, 0};
  A.out = BatchView{out, out_dim, 1, 0};
  A.out2 = BatchView{out2, out_dim, 1, 0};
  A.status = status;
  A.n_samples = 1;
  return A;
}
}  // namespace

// op: 0 state derivative (out: nx), 1 generalised forces (out: na), 2 M and Mdot (out, out2: na x na),
//     3 Tcm and Tcm_dot (out, out2: rows x na), 4 frames (out: 25 per frame)
extern "C" int gen_host_eval(const GenericProgram* G, int op, long long n_samples, const double* x, int nx, const double* u, int nu,
                             double* out, double* out2, int out_dim, int32_t* status) {
  for (long long i = 0; i < n_samples; ++i) {
    int32_t st = 0;
    const EvalArgs A = one_sample(x + i * nx, nx, u ? u + i * nu : x, nu, out + i * out_dim, out_dim, out2 ? out2 + i * out_dim : nullptr, &st);
    if (G->dim == 3) {
      if (op == 0) generic_eval_kernel<3, RKB_GEN_MAX_FRAMES>(G, A);
      else if (op == 1) generic_forces_kernel<3, RKB_GEN_MAX_FRAMES>(G, A);
      else if (op == 2) generic_mass_kernel<3, RKB_GEN_MAX_FRAMES>(G, A);
      else if (op == 3) generic_tmt_kernel<3, RKB_GEN_MAX_FRAMES>(G, A);
      else generic_frames_kernel<3, RKB_GEN_MAX_FRAMES>(G, A);
    } else {
      if (op == 0) generic_eval_kernel<2, RKB_GEN_MAX_FRAMES>(G, A);
      else if (op == 1) generic_forces_kernel<2, RKB_GEN_MAX_FRAMES>(G, A);
      else if (op == 2) generic_mass_kernel<2, RKB_GEN_MAX_FRAMES>(G, A);
      else if (op == 3) generic_tmt_kernel<2, RKB_GEN_MAX_FRAMES>(G, A);
      else generic_frames_kernel<2, RKB_GEN_MAX_FRAMES>(G, A);
    }
    if (status) status[i] = st;
  }
  return 0;
}

// n_steps of RK4 (table == NULL) or of a table-driven scheme, input held constant
extern "C" int gen_host_rollout(const GenericProgram* G, long long n_samples, const double* x0, int nx, const double* u, int nu, double dt,
                                int n_steps, const RkTable* table, double* xout, int32_t* status) {
  for (long long i = 0; i < n_samples; ++i) {
    int32_t st = 0;
    RolloutArgs A;
    A.x0 = ConstBatchView{x0 + i * nx, nx, 1, 0};
    A.u = ConstBatchView{u ? u + i * nu : x0, nu > 0 ? nu : 1, 1, 0};
    A.xout = BatchView{xout + i * nx, nx, 1, 0};
    A.traj = BatchView{nullptr, 0, 0, 0};
    A.status = &st;
    A.n_samples = 1; A.x0_div = 1; A.dt = dt; A.n_steps = n_steps; A.status_or = 0; A.active = nullptr; A.u_node_stride = 0;
    RkTable none = RkTable();
    if (G->dim == 3) {
      if (table) generic_rollout_kernel<3, RKB_GEN_MAX_FRAMES, true>(G, A, *table);
      else generic_rollout_kernel<3, RKB_GEN_MAX_FRAMES, false>(G, A, none);
    } else {
      if (table) generic_rollout_kernel<2, RKB_GEN_MAX_FRAMES, true>(G, A, *table);
      else generic_rollout_kernel<2, RKB_GEN_MAX_FRAMES, false>(G, A, none);
    }
    if (status) status[i] = st;
  }
  return 0;
}
extern "C" int gen_host_program_size(void) { return (int)sizeof(GenericProgram); }
