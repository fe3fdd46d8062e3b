// rkb_probe.cu — instrumentation: measures the FP64 (DFMA pipe) peak of a device, the roofline
// denominator for the serial-chain kernels (MEASURED_PEAKS.json carries HBM and BF16 only).
#include <cuda_runtime.h>
#include "../../include/reak_b200.h"

namespace {

// 16 independent DFMA chains per thread, nothing else in the loop body.  The multiplier comes from
// the constant bank and the addend is the chain's neighbour, so that each DFMA reads two registers
// (no register-bank pressure from a shared third operand).
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, const double a, const double b) {
  double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
  double s0 = r0 * 0.5, s1 = r1 * 0.5, s2 = r2 * 0.5, s3 = r3 * 0.5, s4 = r4 * 0.5, s5 = r5 * 0.5, s6 = r6 * 0.5, s7 = r7 * 0.5;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      r0 = fma(r0, a, r0); r1 = fma(r1, a, r1); r2 = fma(r2, a, r2); r3 = fma(r3, a, r3);
      r4 = fma(r4, a, r4); r5 = fma(r5, a, r5); r6 = fma(r6, a, r6); r7 = fma(r7, a, r7);
      s0 = fma(s0, a, s0); s1 = fma(s1, a, s1); s2 = fma(s2, a, s2); s3 = fma(s3, a, s3);
      s4 = fma(s4, a, s4); s5 = fma(s5, a, s5); s6 = fma(s6, a, s6); s7 = fma(s7, a, s7);
    }
  }
  const double r = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)) + ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
  if (r == 123.456) out[0] = r;  // never true; keeps the chains alive
}

}  // namespace

extern "C" RKB_API int rkb_measure_fp64_peak(int device, double seconds, double* tflops_out, double* sm_clock_mhz_out) {
  if (!tflops_out) return RKB_ERR_INVALID;
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return RKB_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return RKB_ERR_CUDA; }
  double* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) { cudaGetLastError(); return RKB_ERR_NOMEM; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  int iters = 2000;
  double best = 0.0;
  dfma_peak_kernel<<<blocks, threads>>>(d, 200, 1e-9, 0.0);  // warm-up
  cudaDeviceSynchronize();
  double spent = 0.0;
  for (int rep = 0; rep < 50 && spent < seconds; ++rep) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1e-9, 0.0);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    const double tf = flops / ((double)ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    spent += (double)ms * 1e-3;
    if (ms < 20.f) iters *= 2;
  }
  cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (prev >= 0) cudaSetDevice(prev);
  if (err != cudaSuccess) return RKB_ERR_CUDA;
  *tflops_out = best;
  if (sm_clock_mhz_out) *sm_clock_mhz_out = (double)prop.clockRate / 1000.0;
  return RKB_OK;
}
