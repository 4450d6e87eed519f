// Register-only pipe microbenchmark for B200 (sm_100a): measures per-SM throughput of the
// instructions the NTT butterfly is built from.  Not part of the product path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define NCH 8

struct P { uint64_t w, wc, q; uint32_t w0, w1, v0, v1, n0, n1; };
template <int OP>
__global__ void __launch_bounds__(1024) k(uint64_t* out, uint32_t seed, P p) {
  __shared__ double sh[4096];
  if (OP >= 24) { for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = i; __syncthreads(); }
  uint32_t a[NCH], b[NCH];
  uint64_t c[NCH];
  double d[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) {
    a[i] = seed + threadIdx.x * 7 + i;
    b[i] = seed * 3 + i * 5 + 1;
    c[i] = (uint64_t)a[i] * b[i] + i;
    d[i] = 1.0 + i * 1e-9 + threadIdx.x * 1e-12;
  }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i]) : "r"(a[i]), "r"(b[i]));
      if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      if (OP == 3) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(seed + 1));
      if (OP == 4) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001));
      if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      if (OP == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
      if (OP == 7) {  // exact 64-bit Shoup modmul: mul.hi.u64 + 2 mul.lo.u64 + sub
        uint64_t y = c[i], w = p.w + i, wc = p.wc + i, q = p.q;
        uint64_t Q = __umul64hi(wc, y);
        c[i] = w * y - Q * q;
      }
      if (OP == 8) {  // 9-IMAD approximate-quotient modmul (design candidate)
        uint32_t y0 = (uint32_t)c[i], y1 = (uint32_t)(c[i] >> 32);
        uint32_t w0 = p.w0 + i, w1 = p.w1, v0 = p.v0 + i, v1 = p.v1;
        uint32_t n0 = p.n0, n1 = p.n1;
        uint32_t h = __umulhi(v1, y0);
        h = __umulhi(v0, y1) + h;
        uint64_t S = (uint64_t)v1 * y1 + h;
        uint32_t s0 = (uint32_t)S, s1 = (uint32_t)(S >> 32);
        uint64_t r = (uint64_t)w0 * y0;
        uint32_t rh = (uint32_t)(r >> 32) + w1 * y0 + w0 * y1;
        r = ((uint64_t)rh << 32) | (uint32_t)r;
        r = (uint64_t)s0 * n0 + r;
        rh = (uint32_t)(r >> 32) + s1 * n0 + s0 * n1;
        c[i] = ((uint64_t)rh << 32) | (uint32_t)r;
      }
      if (OP == 9) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(c[i]) : "r"((uint32_t)c[i]), "r"(b[i]));
      if (OP == 10) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(seed), "r"(seed)); }
      if (OP == 11) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i]) : "r"(a[i]), "r"(b[i]));
                      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(seed), "r"(seed)); }
      if (OP == 12) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i]) : "r"(a[i]), "r"(b[i]));
                      asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(seed + 1)); }
      if (OP == 13) { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b[i])); c[i] ^= t; }
      if (OP == 14) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001)); }
      if (OP == 15) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001)); }
      if (OP == 16) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(seed), "r"(seed));
                      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001)); }
      if (OP == 17) { float f = __uint_as_float(a[i]); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(1.0001f)); a[i] = __float_as_uint(f); }
      if (OP == 18) { float f = __uint_as_float(a[i]); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(1.0001f)); a[i] = __float_as_uint(f);
                      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(seed), "r"(seed)); }
      if (OP == 19 || OP == 20 || OP == 21) {  // n DFMA per LOP3: n = 2, 3, 4
        constexpr int n = OP - 17;
#pragma unroll
        for (int r = 0; r < n; r++) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      }
      if (OP == 22) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(1.0000001));
      if (OP == 23) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(1.0000001));
      if (OP == 26) asm volatile("cvt.rni.f64.f64 %0, %0;" : "+d"(d[i]));
      if (OP == 27 || OP == 28) {  // n DFMA + one f64 round-to-integer: n = 3, 7
        constexpr int n = OP == 27 ? 3 : 7;
#pragma unroll
        for (int r = 0; r < n; r++) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001));
        double v; asm volatile("cvt.rni.f64.f64 %0, %1;" : "=d"(v) : "d"(d[i]));
        c[i] ^= (uint64_t)__double_as_longlong(v);
      }
      if (OP == 29) { double v; asm volatile("cvt.rn.f64.u64 %0, %1;" : "=d"(v) : "l"(c[i])); c[i] = (uint64_t)__double_as_longlong(v) + i; }
      if (OP == 30) { long long v; asm volatile("cvt.rni.s64.f64 %0, %1;" : "=l"(v) : "d"(d[i])); d[i] = __longlong_as_double(v | 0x3ff0000000000000ll); }
      if (OP == 24 || OP == 25) {  // n DFMA per shared-memory load (8 bytes): n = 4, 8
        constexpr int n = OP == 24 ? 4 : 8;
#pragma unroll
        for (int r = 0; r < n; r++) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001));
        double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((threadIdx.x * 8 + i * 8192) & 0x7ff8));
        c[i] ^= (uint64_t)__double_as_longlong(v);
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += a[i] + b[i] + c[i] + (uint64_t)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int instr_per_iter) {
  int nsm = 148, blocks = nsm * 2, threads = 1024;
  uint64_t* out;
  cudaMalloc(&out, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  P p{0x1234567890abcULL,0x9e3779b97f4a7c15ULL,0x1fffffc800001ULL,0x7890abcu,0x12345u,0x7f4a7c15u,0x4e3779b9u,0x6fffffeu,0xfffc0000u};
  k<OP><<<blocks, threads>>>(out, 12345, p);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 12345 + r, p);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double ops = (double)blocks * threads * ITERS * NCH;
  double rate = ops / (best * 1e-3);
  printf("%-28s %8.3f ms  %9.2f Gop/s  %7.2f op/ns/SM  (x%d instr => %.1f Ginstr/s)\n", name, best, rate * 1e-9,
         rate * 1e-9 / nsm, instr_per_iter, rate * 1e-9 * instr_per_iter);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s, SMs %d, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("mad.lo.u32 (IMAD)", 1);
  run<1>("mad.wide.u32 (IMAD.WIDE)", 1);
  run<2>("mad.hi.u32 (IMAD.HI)", 1);
  run<9>("mul.wide.u32", 1);
  run<3>("add.cc+addc (IADD3 x2)", 2);
  run<4>("fma.rn.f64 (DFMA)", 1);
  run<5>("lop3", 1);
  run<6>("shf", 1);
  run<7>("shoup64 exact modmul", 1);
  run<8>("approx9 modmul", 1);
  run<10>("IMAD + LOP3 (2 instr)", 2);
  run<11>("IMAD.WIDE + LOP3 (2 instr)", 2);
  run<12>("IMAD.WIDE + IADD3 x2 (3 instr)", 3);
  run<13>("mul.wide + 2 lop3 (3 instr)", 3);
  run<14>("IMAD + DFMA (2 instr)", 2);
  run<15>("LOP3 + DFMA (2 instr)", 2);
  run<16>("IMAD + LOP3 + DFMA (3 instr)", 3);
  run<17>("FFMA", 1);
  run<18>("FFMA + IMAD (2 instr)", 2);
  run<19>("2 DFMA + LOP3 (3 instr)", 3);
  run<20>("3 DFMA + LOP3 (4 instr)", 4);
  run<21>("4 DFMA + LOP3 (5 instr)", 5);
  run<22>("add.rn.f64 (DADD)", 1);
  run<23>("mul.rn.f64 (DMUL)", 1);
  run<26>("cvt.rni.f64.f64 (round to integer)", 1);
  run<27>("3 DFMA + cvt.rni.f64.f64 (4 instr)", 4);
  run<28>("7 DFMA + cvt.rni.f64.f64 (8 instr)", 8);
  run<29>("cvt.rn.f64.u64 (+IADD)", 1);
  run<30>("cvt.rni.s64.f64 (+LOP)", 1);
  return 0;
}
