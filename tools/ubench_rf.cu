// Does a DFMA with three distinct 64-bit register operands issue slower than one with two (register-file
// bank limits)?  And what do constant-bank / uniform operands cost?  B200, sm_100a.  Not on the product path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_rf ubench_rf.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 1024
#define NCH 8
__constant__ double kc[4];

template <int OP>
__global__ void __launch_bounds__(512) k(double* out, double s0, double s1) {
  double d[NCH], a[NCH], b[NCH];
  uint32_t x[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) {
    d[i] = 1.0 + i * 1e-9 + threadIdx.x * 1e-12;
    a[i] = 1.0 + i * 3e-9 + threadIdx.x * 1e-13;
    b[i] = 1e-9 * (i + 1) + threadIdx.x * 1e-14;
    x[i] = threadIdx.x * 7 + i;
  }
  const double c0 = kc[0], c1 = kc[1];
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(a[i]));            // 2 distinct regs
      if (OP == 1) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(a[i]), "d"(b[i])); // 3 distinct regs
      if (OP == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(a[i]), "d"(c0));   // 2 regs + constant
      if (OP == 3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c1), "d"(c0));     // 1 reg + 2 constants
      if (OP == 4) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(a[i]));                // DADD 2 regs
      if (OP == 5) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(a[i]));                // DMUL 2 regs
      if (OP == 6) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(a[i]), "d"(s0));   // 2 regs + kernel param
      if (OP == 7) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d[i]) : "d"(a[i]), "d"(b[i]), "d"(d[(i + 1) % NCH])); // 3 distinct + separate dest
      if (OP == 8) {  // the product path's modmul as written today (6 FP64 instr)
        const double y = d[i], w = a[i], winv = b[i];
        const double cc = __dadd_rn(__fma_rn(y, winv, c0), -c0);
        const double h = __dmul_rn(y, w), l = __fma_rn(y, w, -h);
        d[i] = __dadd_rn(__fma_rn(-cc, c1, h), l);
      }
      if (OP == 9) {  // same with the quotient taken from h (8-byte twiddles)
        const double y = d[i], w = a[i];
        const double h = __dmul_rn(y, w), l = __fma_rn(y, w, -h);
        const double cc = __dadd_rn(__fma_rn(h, s1, c0), -c0);
        d[i] = __dadd_rn(__fma_rn(-cc, c1, h), l);
      }
      if (OP == 10) {  // 1 DFMA(2 regs) + 1 independent LOP3
        asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(a[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) % NCH]), "r"(it));
      }
      if (OP == 11) {  // 1 DFMA(3 regs) + 1 independent LOP3
        asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(a[i]), "d"(b[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) % NCH]), "r"(it));
      }
      if (OP == 12) {  // 2 DADD + 1 LOP3
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(a[i]));
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) % NCH]), "r"(it));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += d[i] + a[i] + b[i] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int fp64_per_iter, int threads, int blocks_per_sm) {
  int nsm = 148, blocks = nsm * blocks_per_sm;
  double* out;
  cudaMalloc(&out, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<blocks, threads>>>(out, 1.0000001, 1e-15);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 1.0000001, 1e-15);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double groups = (double)blocks * threads * ITERS * NCH;
  // cycles per warp-level group per scheduler (4 schedulers per SM), at 1.965 GHz
  double warp_groups_per_sched = groups / 32.0 / (nsm * 4.0);
  double cyc = best * 1e-3 * 1.965e9 / warp_groups_per_sched;
  printf("%-44s warps/sched %2d  %8.3f ms  %6.2f cyc/group  (%.2f cyc per FP64 instr)\n", name,
         threads * blocks_per_sm / 128, best, cyc, fp64_per_iter ? cyc / fp64_per_iter : 0.0);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
  cudaFree(out);
}

int main() {
  double h[4] = {6755399441055744.0, 562949894668289.0, 0, 0};
  cudaMemcpyToSymbol(kc, h, sizeof(h));
  for (int cfg = 0; cfg < 2; cfg++) {
    int th = 512, bps = cfg == 0 ? 1 : 4;  // 4 or 16 warps per scheduler
    run<0>("DFMA d=d*a+a (2 distinct regs)", 1, th, bps);
    run<1>("DFMA d=a*b+d (3 distinct regs)", 1, th, bps);
    run<7>("DFMA d=a*b+e (3 distinct + other dest)", 1, th, bps);
    run<2>("DFMA d=d*a+const", 1, th, bps);
    run<3>("DFMA d=d*const+const", 1, th, bps);
    run<6>("DFMA d=a*param+d", 1, th, bps);
    run<4>("DADD d=d+a", 1, th, bps);
    run<5>("DMUL d=d*a", 1, th, bps);
    run<8>("modmul, quotient from y*winv (6 FP64)", 6, th, bps);
    run<9>("modmul, quotient from h*qinv (6 FP64)", 6, th, bps);
    run<10>("DFMA(2 regs) + LOP3", 1, th, bps);
    run<11>("DFMA(3 regs) + LOP3", 1, th, bps);
    run<12>("2 DADD + LOP3", 2, th, bps);
  }
  return 0;
}
