// small_batch_microbench.cu — developer tool: the numbers that decide how a SMALL batch (a few hundred to a few
// thousand samples, the planner's regime) should be mapped to the machine.  One warp per SM sub-partition is the
// situation of the thread-per-sample kernels at 1024 samples; the questions are
//   (1) how long a dependent DFMA takes (the floor for one sample's critical path),
//   (2) how fast ONE warp issues independent DFMAs, and whether a half-empty warp issues faster,
//   (3) what a 64-bit value costs to move between lanes (SHFL) or between warps (shared memory + named barrier),
// because a sub-warp or cross-warp split of one sample's evaluation pays (3) to buy issue slots (2).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/small_batch_microbench tools/small_batch_microbench.cu
//   gpurun -- tools/bin/small_batch_microbench
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ long long clk() { return clock64(); }

// (1) dependent chain: r = fma(r, a, b), ITER times, one warp
__global__ void dep_dfma(double* out, long long* cyc, double a, double b, int iters) {
  double r = threadIdx.x * 1e-3;
  const long long t0 = clk();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 32; ++k) r = fma(r, a, b);
  }
  const long long t1 = clk();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (r == 123.456) out[0] = r;
}

// (2) ILP independent chains in one warp; only lanes < active take part (the others exit)
template <int ILP>
__global__ void ilp_dfma(double* out, long long* cyc, double a, double b, int iters, int active) {
  if ((int)(threadIdx.x & 31) >= active) return;
  double r[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) r[i] = threadIdx.x * 1e-3 + i;
  const long long t0 = clk();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int i = 0; i < ILP; ++i) r[i] = fma(r[i], a, b);
  }
  const long long t1 = clk();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += r[i];
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
  if (s == 123.456) out[0] = s;
}

// (3a) dependent 64-bit shuffle chain (two SHFL.32 per double)
__global__ void dep_shfl(double* out, long long* cyc, int iters) {
  double r = threadIdx.x * 1.5;
  const long long t0 = clk();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 32; ++k) r = __shfl_sync(0xffffffffu, r, (threadIdx.x + 1) & 31);
  }
  const long long t1 = clk();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (r == 123.456) out[0] = r;
}
// (3b) throughput: ILP independent DFMA chains with S 64-bit shuffles per 8 DFMAs
template <int S>
__global__ void mix_shfl(double* out, long long* cyc, double a, double b, int iters) {
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 1e-3 + i;
  const long long t0 = clk();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      r[i] = fma(r[i], a, b);
      if ((i % 8) < S) r[(i + 8) % 16] = __shfl_xor_sync(0xffffffffu, r[(i + 8) % 16], 1);
    }
  }
  const long long t1 = clk();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (s == 123.456) out[0] = s;
}

// (3c) two warps of one CTA hand a double back and forth through shared memory with named barriers:
// warp 0 writes, bar, warp 1 reads + writes, bar, ... one round trip = 2 hand-offs
__global__ void ping_pong(double* out, long long* cyc, int iters) {
  __shared__ double box[2][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double r = l;
  box[0][l] = 0; box[1][l] = 0;
  __syncthreads();
  const long long t0 = clk();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (w == 0) {
      box[0][l] = r + 1.0;
      asm volatile("bar.arrive 1, 64;");   // tell warp 1 the value is there
      asm volatile("bar.sync 2, 64;");     // wait for its answer
      r = box[1][l];
    } else {
      asm volatile("bar.sync 1, 64;");
      r = box[0][l] + 1.0;
      box[1][l] = r;
      asm volatile("bar.arrive 2, 64;");
    }
  }
  const long long t1 = clk();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (r == -1.0) out[0] = r;
}

int main() {
  double* d; long long* c; long long h[1024];
  cudaMalloc(&d, 64); cudaMalloc(&c, sizeof h);
  const int iters = 2000;
  dep_dfma<<<1, 32>>>(d, c, 1.0000001, 1e-9, iters); cudaDeviceSynchronize();
  dep_dfma<<<1, 32>>>(d, c, 1.0000001, 1e-9, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency            : %.2f cycles\n", (double)h[0] / (32.0 * iters));
  for (int active : {32, 16, 8, 1}) {
    ilp_dfma<16><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters, active); cudaDeviceSynchronize();
    ilp_dfma<16><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters, active); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
    printf("one warp, 16 chains, %2d lanes     : %.2f cycles per DFMA\n", active, (double)h[0] / (8.0 * 16 * iters));
  }
  for (int warps : {2, 4, 8}) {  // several warps of one CTA: 4 = one per sub-partition, 8 = two per sub-partition
    ilp_dfma<16><<<1, 32 * warps>>>(d, c, 1.0000001, 1e-9, iters, 32); cudaDeviceSynchronize();
    ilp_dfma<16><<<1, 32 * warps>>>(d, c, 1.0000001, 1e-9, iters, 32); cudaMemcpy(h, c, 8 * warps, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%d warps in one CTA, 16 chains     : %.2f cycles per DFMA per warp\n", warps, (double)mx / (8.0 * 16 * iters));
  }
  for (int warps : {8}) {  // two half-full warps per sub-partition against one full one
    ilp_dfma<16><<<1, 32 * warps>>>(d, c, 1.0000001, 1e-9, iters, 16); cudaDeviceSynchronize();
    ilp_dfma<16><<<1, 32 * warps>>>(d, c, 1.0000001, 1e-9, iters, 16); cudaMemcpy(h, c, 8 * warps, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%d half-full warps in one CTA      : %.2f cycles per DFMA per warp\n", warps, (double)mx / (8.0 * 16 * iters));
  }
  ilp_dfma<2><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters, 32); cudaDeviceSynchronize();
  ilp_dfma<2><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters, 32); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
  printf("one warp, 2 chains                : %.2f cycles per DFMA\n", (double)h[0] / (8.0 * 2 * iters));
  ilp_dfma<4><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters, 32); cudaDeviceSynchronize();
  ilp_dfma<4><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters, 32); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
  printf("one warp, 4 chains                : %.2f cycles per DFMA\n", (double)h[0] / (8.0 * 4 * iters));
  dep_shfl<<<1, 32>>>(d, c, iters); cudaDeviceSynchronize();
  dep_shfl<<<1, 32>>>(d, c, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
  printf("dependent 64-bit shuffle          : %.2f cycles\n", (double)h[0] / (32.0 * iters));
  mix_shfl<0><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters); cudaDeviceSynchronize();
  mix_shfl<0><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
  const double base = (double)h[0] / (16.0 * iters);
  printf("16 DFMA, no shuffle               : %.2f cycles per DFMA\n", base);
  mix_shfl<2><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters); cudaDeviceSynchronize();
  mix_shfl<2><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
  printf("16 DFMA + 4 64-bit shuffles       : %.2f cycles per DFMA (%.2f extra cycles per shuffle)\n", (double)h[0] / (16.0 * iters),
         ((double)h[0] / iters - 16.0 * base) / 4.0);
  mix_shfl<8><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters); cudaDeviceSynchronize();
  mix_shfl<8><<<1, 32>>>(d, c, 1.0000001, 1e-9, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
  printf("16 DFMA + 16 64-bit shuffles      : %.2f cycles per DFMA (%.2f extra cycles per shuffle)\n", (double)h[0] / (16.0 * iters),
         ((double)h[0] / iters - 16.0 * base) / 16.0);
  ping_pong<<<1, 64>>>(d, c, iters); cudaDeviceSynchronize();
  ping_pong<<<1, 64>>>(d, c, iters); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
  printf("smem + named-barrier hand-off     : %.1f cycles per one-way hand-off between two warps\n", (double)h[0] / (2.0 * iters));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
