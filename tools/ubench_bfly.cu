// Register-resident butterfly-network microbenchmark (no DRAM traffic): how many lazy Harvey butterflies per
// clock an SM sustains for a given radix / thread count.  This is the "IMAD roofline" of the NTT kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<pkg>/csrc -o ubench_bfly ubench_bfly.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ntt_device.cuh"
using namespace nttb200;

template <int R, int THREADS, int MINB, bool SMEMTW>
__global__ void __launch_bounds__(THREADS, MINB) k(uint64_t* a, const uint4* wu, const uint2* qq, ntt_cuda_params_t p, int iters)
{
  __shared__ uint4 swu[64];
  __shared__ uint2 sqq[64];
  if (threadIdx.x < 64) { swu[threadIdx.x] = wu[threadIdx.x]; sqq[threadIdx.x] = qq[threadIdx.x]; }
  __syncthreads();
  constexpr int n = 1 << R;
  uint64_t x[n];
#pragma unroll
  for (int i = 0; i < n; i++) x[i] = a[(size_t)blockIdx.x * THREADS * n + threadIdx.x + i * THREADS];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < R; u++) {
      const int d = n >> (u + 1);
#pragma unroll
      for (int sub = 0; sub < (1 << u); sub++) {
        Mulc m;
        if (SMEMTW) { uint4 A = swu[(1 << u) + sub]; uint2 B = sqq[(1 << u) + sub]; m = Mulc{A.x, A.y, A.z, A.w, B.x, B.y}; }
        else        { uint4 A = __ldg(wu + (1 << u) + sub); uint2 B = __ldg(qq + (1 << u) + sub); m = Mulc{A.x, A.y, A.z, A.w, B.x, B.y}; }
#pragma unroll
        for (int kk = 0; kk < d; kk++) bfly_fwd<false>(x[sub * 2 * d + kk], x[sub * 2 * d + kk + d], m, p, p.c10q);
      }
    }
    if ((it & 7) == 7) {  // keep values bounded like the real kernel does once per transform
      const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
#pragma unroll
      for (int i = 0; i < n; i++) x[i] = reduce_2q(x[i], rc);
    }
  }
#pragma unroll
  for (int i = 0; i < n; i++) a[(size_t)blockIdx.x * THREADS * n + threadIdx.x + i * THREADS] = x[i];
}

template <int R, int THREADS, int MINB, bool SMEMTW>
void run(const char* name, uint64_t* a, uint4* wu, uint2* qq, ntt_cuda_params_t p)
{
  const int iters = 256, blocks = 148 * MINB;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<R, THREADS, MINB, SMEMTW><<<blocks, THREADS>>>(a, wu, qq, p, 8);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    k<R, THREADS, MINB, SMEMTW><<<blocks, THREADS>>>(a, wu, qq, p, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<R, THREADS, MINB, SMEMTW>);
  const double bfly = (double)blocks * THREADS * iters * R * (1 << (R - 1));
  const double per_s = bfly / (best * 1e-3);
  printf("%-34s regs %3d  %7.3f ms  %8.1f Gbfly/s  => %6.2f M NTT/s at N=2^14 (114688 bfly)  cyc/warp-bfly/SMSP %.1f\n", name, fa.numRegs, best,
         per_s * 1e-9, per_s / 114688 * 1e-6, 1.965e9 * 148 * 4 * 32 / per_s);
  cudaError_t err = cudaGetLastError(); if (err != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(err));
}

int main()
{
  const uint64_t q = 0x1fffffc800001ULL;
  ntt_cuda_params_t p{};
  p.q = q; p.neg2q = 0 - 2 * q; p.negq = 0 - q; p.c10q = 10 * q; p.red_shift = 49 - 9;
  p.red_mu = (uint32_t)((((unsigned __int128)1) << (32 + p.red_shift)) / q);
  uint64_t* a; uint4* wu; uint2* qq;
  cudaMalloc(&a, (size_t)148 * 4 * 1024 * 32 * 8); cudaMemset(a, 1, (size_t)148 * 4 * 1024 * 32 * 8);
  cudaMalloc(&wu, 64 * 16); cudaMalloc(&qq, 64 * 8);
  uint4 hwu[64]; uint2 hqq[64];
  for (int i = 0; i < 64; i++) {
    uint64_t w = (0x123456789abcdefULL * (i + 3)) % q; uint64_t u = (uint64_t)((((unsigned __int128)w) << 32) % q);
    hwu[i] = make_uint4((uint32_t)w, (uint32_t)(w >> 32), (uint32_t)u, (uint32_t)(u >> 32));
    hqq[i] = make_uint2((uint32_t)((((unsigned __int128)w) << 30) / q), (uint32_t)((((unsigned __int128)u) << 30) / q));
  }
  cudaMemcpy(wu, hwu, sizeof(hwu), cudaMemcpyHostToDevice); cudaMemcpy(qq, hqq, sizeof(hqq), cudaMemcpyHostToDevice);
  run<5, 512, 1, true>("radix32  512thr x1 smem-tw", a, wu, qq, p);
  run<5, 512, 1, false>("radix32  512thr x1 ldg-tw", a, wu, qq, p);
  run<5, 256, 1, true>("radix32  256thr x1 smem-tw", a, wu, qq, p);
  run<5, 256, 2, true>("radix32  256thr x2 smem-tw", a, wu, qq, p);
  run<4, 1024, 1, true>("radix16 1024thr x1 smem-tw", a, wu, qq, p);
  run<4, 512, 2, true>("radix16  512thr x2 smem-tw", a, wu, qq, p);
  run<4, 512, 1, true>("radix16  512thr x1 smem-tw", a, wu, qq, p);
  run<4, 768, 1, true>("radix16  768thr x1 smem-tw", a, wu, qq, p);
  run<3, 1024, 1, true>("radix8  1024thr x1 smem-tw", a, wu, qq, p);
  run<3, 1024, 2, true>("radix8  1024thr x2 smem-tw", a, wu, qq, p);
  return 0;
}
