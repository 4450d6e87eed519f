// FP64 pipe: dependent-issue latency and how much ILP 4 warps per scheduler need to saturate it (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64lat ubench_fp64lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int NCH, int MIX>  // MIX 0: DFMA only; 1: DFMA,DADD,DMUL rotation
__global__ void k(double* out, double a, double b)
{
  double d[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) d[i] = 1.0 + i * 1e-9 + threadIdx.x * 1e-12;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      if (MIX == 0 || (it % 3) == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(a), "d"(b));
      else if ((it % 3) == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(b));
      else asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(a));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCH, int MIX>
void run(int threads)
{
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NCH, MIX><<<148, threads>>>(out, 1.0000001, 1e-9); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NCH, MIX><<<148, threads>>>(out, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9;
  const double warp_instr_per_smsp = (double)ITERS * NCH * (threads / 32) / 4.0;
  printf("threads %4d  chains %2d  mix %d: %8.3f ms  %6.2f cycles per dependent step  %5.3f FP64 warp-instr/clk/SMSP\n", threads, NCH,
         MIX, ms, cyc / ITERS, warp_instr_per_smsp / cyc);
  cudaFree(out);
}
int main()
{
  run<1, 0>(32); run<1, 1>(32); run<2, 0>(32); run<4, 0>(32); run<8, 0>(32); run<16, 0>(32);
  run<1, 0>(128); run<2, 0>(128); run<4, 0>(128); run<8, 0>(128);
  run<1, 0>(512); run<2, 0>(512); run<4, 0>(512); run<8, 0>(512); run<16, 0>(512); run<8, 1>(512);
  run<4, 0>(1024); run<8, 0>(1024);
  return 0;
}
