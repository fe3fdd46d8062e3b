// fp64_microbench.cu — developer tool: what the B200 FP64 pipe sustains for operand patterns,
// occupancies and instruction mixes that look like the serial-chain kernels, as opposed to the
// two-register DFMA loop that defines the nominal peak (rkb_measure_fp64_peak).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_microbench tools/fp64_microbench.cu
//   gpurun -- tools/bin/fp64_microbench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHAINS 16

// PATTERN 0: r = fma(r, const, r)            two register reads per DFMA
// PATTERN 1: r_i = fma(r_i, r_j, r_k)        three distinct registers
// PATTERN 2: r_i = fma(r_i, c[bank], r_k)    constant-bank multiplier, two distinct registers
// PATTERN 3: like 1 but mul / add / fma mixed 1:1:2 (DMUL, DADD, DFMA)
// PATTERN 4-6: DMUL r*const, DADD r+const, DMUL r*r;  7-9: three registers at other chain distances
// MIXI: integer instructions inserted per 8 FP64 instructions (0, 3, 5)
template <int PATTERN, int ILP, int MIXI>
__global__ void __launch_bounds__(128) fp64_kernel(double* out, int iters, const double a, const double b, int salt) {
  double r[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) r[i] = threadIdx.x * 1e-3 + i;
  unsigned x = threadIdx.x + salt, y = blockIdx.x * 7 + salt;
  const double sv = a * (double)(threadIdx.x + 1), tv = b * (double)(threadIdx.x + 3);  // per-thread values: vector registers
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        const int j = (i + 5) % CHAINS, k = (i + 11) % CHAINS;
        if (PATTERN == 0) r[i] = fma(r[i], a, r[i]);
        else if (PATTERN == 1) r[i] = fma(r[i], r[j], r[k]);
        else if (PATTERN == 2) r[i] = fma(r[i], a, r[k]);
        else if (PATTERN == 4) r[i] = r[i] * a;                      // DMUL, one register + constant
        else if (PATTERN == 5) r[i] = r[i] + b;                      // DADD, one register + constant
        else if (PATTERN == 6) r[i] = r[i] * r[j];                   // DMUL, two registers
        else if (PATTERN == 7) r[i] = fma(r[i], r[(i + 1) % CHAINS], r[(i + 2) % CHAINS]);  // neighbours: different banks
        else if (PATTERN == 8) r[i] = fma(r[i], r[(i + 2) % CHAINS], r[(i + 4) % CHAINS]);  // same parity
        else if (PATTERN == 9) r[i] = fma(r[i], r[(i + 1) % CHAINS], r[(i + 3) % CHAINS]);
        else if (PATTERN == 10) r[i] = fma(r[i], sv, tv);            // three registers, two of them shared by every chain (.reuse)
        else if (PATTERN == 11) r[i] = fma(r[i], sv, r[k]);          // three registers, one shared
        else if (PATTERN == 12) r[i] = r[i] * sv;                    // DMUL two registers, one shared
        else {
          if ((i & 3) == 0) r[i] = r[i] * r[j];
          else if ((i & 3) == 1) r[i] = r[i] + r[k];
          else r[i] = fma(r[i], r[j], r[k]);
        }
        if (MIXI > 0 && (i % 8) < MIXI) x = x * 1664525u + y;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += r[i];
  if (s == 123.456 || x == 0x12345u) out[0] = s + y;
}

template <int PATTERN, int ILP, int MIXI>
void run(const char* name, int ctas_per_sm, int sms, double* d) {
  auto kern = fp64_kernel<PATTERN, ILP, MIXI>;
  // dynamic shared memory sized so that exactly ctas_per_sm CTAs are resident
  const int smem = (227 * 1024 / ctas_per_sm - 1024) & ~1023;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int blocks = sms * ctas_per_sm;
  int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, 128, smem>>>(d, 100, 1e-9, 0.5, 1);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    kern<<<blocks, 128, smem>>>(d, iters, 1e-9, 0.5, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double instr = 8.0 * ILP * (double)iters * blocks * 128;
  const double per_clk_sm = instr / (best * 1e-3) / 1.965e9 / sms;
  printf("%-34s warps/SMSP %2d  ILP %2d  int/8 %d : %6.2f FP64 instr/clk/SM (%.1f %% of 64)  err=%s\n", name, ctas_per_sm, ILP, MIXI, per_clk_sm,
         100.0 * per_clk_sm / 64.0, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 64);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  for (int c : {4}) {
    run<4, 16, 0>("mul(r,const)", c, sms, d);
    run<5, 16, 0>("add(r,const)", c, sms, d);
    run<6, 16, 0>("mul(r_i,r_j)", c, sms, d);
    run<7, 16, 0>("fma(r_i,r_i+1,r_i+2)", c, sms, d);
    run<8, 16, 0>("fma(r_i,r_i+2,r_i+4)", c, sms, d);
    run<9, 16, 0>("fma(r_i,r_i+1,r_i+3)", c, sms, d);
    run<10, 16, 0>("fma(r_i,s,t) s,t shared regs", c, sms, d);
    run<11, 16, 0>("fma(r_i,s,r_k) s shared reg", c, sms, d);
    run<12, 16, 0>("mul(r_i,s) s shared reg", c, sms, d);
    run<10, 16, 5>("fma(r_i,s,t) + 5 int per 8", c, sms, d);
    run<12, 16, 5>("mul(r_i,s) + 5 int per 8", c, sms, d);
    run<6, 16, 5>("mul(r_i,r_j) + 5 int per 8", c, sms, d);
    run<4, 16, 3>("mul(r,const) + 3 int per 8", c, sms, d);
    run<0, 16, 3>("fma(r,const,r) + 3 int per 8", c, sms, d);
    run<0, 16, 5>("fma(r,const,r) + 5 int per 8", c, sms, d);
  }
  for (int c : {1, 2, 4, 8, 16}) {
    run<0, 16, 0>("fma(r,const,r)", c, sms, d);
    run<1, 16, 0>("fma(r_i,r_j,r_k)", c, sms, d);
    run<2, 16, 0>("fma(r_i,const,r_k)", c, sms, d);
    run<3, 16, 0>("mul/add/fma mix, 3 regs", c, sms, d);
    run<1, 16, 3>("fma 3 regs + 3 int per 8", c, sms, d);
    run<1, 16, 5>("fma 3 regs + 5 int per 8", c, sms, d);
    run<1, 4, 0>("fma 3 regs", c, sms, d);
    run<1, 2, 0>("fma 3 regs", c, sms, d);
    run<1, 1, 0>("fma 3 regs", c, sms, d);
  }
  return 0;
}
