// rkb_linearize.cu — helper kernels of rkb_linearize: the perturbed batch that the evaluation kernels are run on, and
// the central differences taken from their results.  (The evaluation itself is the ordinary rkb_eval path.)
#include <cuda_runtime.h>
#include "rkb_internal.h"

namespace {

// row r = (i * D + d) * 2 + s of the perturbed batch: state / input of sample i with component d moved by +h (s = 0)
// or -h (s = 1); d < nx addresses the state, d >= nx the input d - nx.  The step is relative to the component's size:
// h_d = eps * max(1, |value|).
__global__ void __launch_bounds__(256) lin_perturb_kernel(long long n, int nx, int nu, double eps, const double* __restrict__ x,
                                                           const double* __restrict__ u, double* __restrict__ xp, double* __restrict__ up) {
  const int D = nx + nu;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n * D * 2) return;
  const long long i = r / (2 * D);
  const int d = (int)((r / 2) % D), s = (int)(r & 1);
  for (int k = 0; k < nx; ++k) {
    double v = x[i * nx + k];
    if (k == d) { const double h = eps * fmax(1.0, fabs(v)); v = s ? v - h : v + h; }
    xp[r * nx + k] = v;
  }
  for (int k = 0; k < nu; ++k) {
    double v = u[i * nu + k];
    if (nx + k == d) { const double h = eps * fmax(1.0, fabs(v)); v = s ? v - h : v + h; }
    up[r * nu + k] = v;
  }
}

// A[i][row][d] = (f_row(x + h e_d) - f_row(x - h e_d)) / ((x_d + h) - (x_d - h)), likewise B for the inputs: the
// divisor is the difference of the perturbed values as they were actually formed, not 2 h.
__global__ void __launch_bounds__(256) lin_combine_kernel(long long n, int nx, int nu, const double* __restrict__ xp, const double* __restrict__ up,
                                                           const double* __restrict__ fd, const int32_t* __restrict__ st_in, double* __restrict__ A,
                                                           double* __restrict__ B, int32_t* __restrict__ status) {
  const int D = nx + nu;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * D) return;
  const long long i = t / D;
  const int d = (int)(t % D);
  const long long r0 = t * 2, r1 = t * 2 + 1;
  const double hi = d < nx ? xp[r0 * nx + d] : up[r0 * nu + (d - nx)];
  const double lo = d < nx ? xp[r1 * nx + d] : up[r1 * nu + (d - nx)];
  const double inv = 1.0 / (hi - lo);
  for (int row = 0; row < nx; ++row) {
    const double v = (fd[r0 * nx + row] - fd[r1 * nx + row]) * inv;
    if (d < nx) { if (A) A[(i * nx + row) * nx + d] = v; }
    else if (B) B[(i * nx + row) * nu + (d - nx)] = v;
  }
  if (status && st_in) {
    const int s = st_in[r0] | st_in[r1];
    if (s) atomicOr(&status[i], s);
  }
}

}  // namespace

cudaError_t rkb_lin_perturb(long long n, int nx, int nu, double eps, const double* x, const double* u, double* xp, double* up, cudaStream_t s) {
  const long long rows = n * (nx + nu) * 2;
  if (rows <= 0) return cudaSuccess;
  lin_perturb_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(n, nx, nu, eps, x, u, xp, up);
  return cudaGetLastError();
}
cudaError_t rkb_lin_combine(long long n, int nx, int nu, const double* xp, const double* up, const double* fd, const int32_t* st_in, double* A,
                            double* B, int32_t* status, cudaStream_t s) {
  const long long t = n * (nx + nu);
  if (t <= 0) return cudaSuccess;
  lin_combine_kernel<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(n, nx, nu, xp, up, fd, st_in, A, B, status);
  return cudaGetLastError();
}
