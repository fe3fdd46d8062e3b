// rkb_steer.cu — the control law between two control intervals of rkb_steer_feedback: decides per
// sample whether the steering loop goes on and, if so, computes the input of the next interval.
//
// Reference semantics (paths relative to ReaK's source tree):
//   loop condition and feedback       examples/misc/MEAQR_topology.hpp:503-561 (steer_with_constant_control),
//                                     examples/misc/IHAQR_topology.hpp:349-378 (move_position_toward_impl)
//   get_bounded_input                 examples/misc/IHAQR_topology.hpp:304-327
//   bring_point_in_bounds / is_in_bounds of the hyperbox input spaces
//                                     ctrl/topologies/hyperbox_topology.hpp:114-128, 178-189
//
// One thread per sample.  The gain matrix (n_inputs x 2n doubles per sample) is the only sizeable
// operand; it is read once per interval with 128-bit loads.
#include <cuda_runtime.h>
#include <math.h>

#include "rkb_internal.h"

namespace {

#define STEER_BLOCK 128

template <bool BOUNDED>
__device__ __forceinline__ void clamp_box(int nu, const double* lo, const double* hi, double* a) {
  if (!BOUNDED) return;
  for (int i = 0; i < nu; ++i) {
    if (a[i] < lo[i]) a[i] = lo[i];
    else if (a[i] > hi[i]) a[i] = hi[i];
  }
}
template <bool BOUNDED>
__device__ __forceinline__ bool in_box(int nu, const double* lo, const double* hi, const double* a) {
  if (!BOUNDED) return true;
  for (int i = 0; i < nu; ++i)
    if ((a[i] < lo[i]) || (a[i] > hi[i])) return false;
  return true;
}

__global__ void __launch_bounds__(STEER_BLOCK) steer_law_kernel(const SteerLawArgs A) {
  const long long i = (long long)blockIdx.x * STEER_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  const int nx = A.nx, nu = A.nu, k = A.interval;
  if (k == 0) A.n_done[i] = 0;
  else if (A.n_done[i] != k) { A.active[i] = 0; return; }  // the loop of this sample has ended earlier
  const double* x = (k == 0 ? A.x0 : A.x) + i * nx;
  const double* goal = A.goal + i * nx;
  double dx[2 * RKB_MAX_COORDS];
  double d2 = 0.0;
  for (int c = 0; c < nx; ++c) {
    dx[c] = x[c] - goal[c];
    d2 += dx[c] * dx[c];
  }
  if (k == 0 && A.x != A.x0)
    for (int c = 0; c < nx; ++c) A.x[i * nx + c] = x[c];  // the state buffer the intervals work on in place
  if (!(sqrt(d2) > A.proximity)) { A.active[i] = 0; return; }  // MEAQR_topology.hpp:513-514
  double corr[RKB_MAX_COORDS], u_bias[RKB_MAX_COORDS], u_prev[RKB_MAX_COORDS], u_cur[RKB_MAX_COORDS], du[RKB_MAX_COORDS];
  const double* G = A.gain + i * (long long)nu * nx;
  for (int r = 0; r < nu; ++r) {
    double s = 0.0;
    if ((nx & 1) == 0 && ((reinterpret_cast<unsigned long long>(G + r * nx) & 15ull) == 0)) {
      const double2* g2 = reinterpret_cast<const double2*>(G + r * nx);
      for (int c = 0; c < nx; c += 2) {
        const double2 g = g2[c >> 1];
        s += g.x * dx[c];
        s += g.y * dx[c + 1];
      }
    } else {
      for (int c = 0; c < nx; ++c) s += G[r * nx + c] * dx[c];
    }
    corr[r] = -s;
    u_bias[r] = A.u_bias[i * nu + r];
    u_prev[r] = A.u_prev[i * nu + r];
  }
  const double T = A.time_step;
  // u_prev = u_current once the interval is accepted (MEAQR_topology.hpp:553); with a collision test the
  // acceptance comes later (steer_commit_kernel) and the input waits in u_next
  double* u_out = (A.u_next ? A.u_next : A.u_prev) + i * nu;
  if (k == 0 && !A.saturate_first) {  // MEAQR_topology.hpp:521-522
    for (int r = 0; r < nu; ++r) u_out[r] = u_bias[r] + corr[r];
  } else {
    // get_bounded_input(u_prev, u_bias, u_correction), IHAQR_topology.hpp:304-327
    const bool bu = A.have_u_box != 0, bd = A.have_du_box != 0;
    if (bu) clamp_box<true>(nu, A.u_lo, A.u_hi, u_bias);
    for (int r = 0; r < nu; ++r) u_cur[r] = u_bias[r] + corr[r];
    const bool inside = bu ? in_box<true>(nu, A.u_lo, A.u_hi, u_cur) : true;
    if (inside) {
      for (int r = 0; r < nu; ++r) du[r] = (u_cur[r] - u_prev[r]) * (1.0 / T);
    } else {
      for (int j = 0; j < 10; ++j) {
        for (int r = 0; r < nu; ++r) { corr[r] *= 0.5; u_cur[r] -= corr[r]; }
        if (in_box<true>(nu, A.u_lo, A.u_hi, u_cur))
          for (int r = 0; r < nu; ++r) { u_bias[r] = u_cur[r]; u_cur[r] += corr[r]; }
      }
      for (int r = 0; r < nu; ++r) du[r] = (u_bias[r] - u_prev[r]) * (1.0 / T);
    }
    if (bd) clamp_box<true>(nu, A.du_lo, A.du_hi, du);
    for (int r = 0; r < nu; ++r) u_out[r] = u_prev[r] + T * du[r];
  }
  A.active[i] = 1;
  A.n_done[i] = k + 1;
}

// `if((!with_collision_check) || is_free_impl(x_next)) { accept } else { was_collision_free = false; break; }`
// (MEAQR_topology.hpp:550-559, IHAQR_topology.hpp:369-375): is_free = no proxy pair reports a negative minimum
// distance (MEAQR_topology.hpp:921-940).
__global__ void __launch_bounds__(STEER_BLOCK) steer_commit_kernel(const SteerCommitArgs A) {
  const long long i = (long long)blockIdx.x * STEER_BLOCK + threadIdx.x;
  if (i >= A.n_samples) return;
  if (!A.active[i]) return;
  bool is_free = true;
  for (int p = 0; p < A.n_pairs; ++p)
    if (A.dist[(long long)p * A.n_samples + i] < 0.0) is_free = false;
  if (is_free) {
    for (int c = 0; c < A.nx; ++c) {
      const double v = A.x_next[i * A.nx + c];
      A.x[i * A.nx + c] = v;
      if (A.traj) A.traj[(i * A.max_intervals + A.interval) * A.nx + c] = v;
    }
    for (int r = 0; r < A.nu; ++r) A.u_prev[i * A.nu + r] = A.u_next[i * A.nu + r];
  } else {
    A.n_done[i] = A.interval;  // the law kernel counted this interval in advance
    A.active[i] = 0;
    A.collided[i] = 1;
  }
}

// is_free: no pair with a negative minimum distance (manip_free_workspace.hpp:84-97)
__global__ void __launch_bounds__(STEER_BLOCK) free_combine_kernel(const double* __restrict__ dist, int n_pairs, long long n, int32_t* out) {
  const long long i = (long long)blockIdx.x * STEER_BLOCK + threadIdx.x;
  if (i >= n) return;
  int ok = 1;
  for (int p = 0; p < n_pairs; ++p)
    if (dist[(long long)p * n + i] < 0.0) ok = 0;
  out[i] = ok;
}

}  // namespace

cudaError_t rkb_free_combine(const double* dist, int n_pairs, long long n, int32_t* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  free_combine_kernel<<<(unsigned)((n + STEER_BLOCK - 1) / STEER_BLOCK), STEER_BLOCK, 0, s>>>(dist, n_pairs, n, out);
  return cudaGetLastError();
}

cudaError_t rkb_steer_commit(const SteerCommitArgs& a, cudaStream_t s) {
  if (a.n_samples <= 0) return cudaSuccess;
  steer_commit_kernel<<<(unsigned)((a.n_samples + STEER_BLOCK - 1) / STEER_BLOCK), STEER_BLOCK, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t rkb_steer_law(const SteerLawArgs& a, cudaStream_t s) {
  if (a.n_samples <= 0) return cudaSuccess;
  steer_law_kernel<<<(unsigned)((a.n_samples + STEER_BLOCK - 1) / STEER_BLOCK), STEER_BLOCK, 0, s>>>(a);
  return cudaGetLastError();
}
