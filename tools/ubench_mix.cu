// Feasibility microbenchmark: do integer butterflies (FMA/ALU pipes) and FP64 butterflies (FP64 pipe) overlap
// when different warps of the same SM run them?  Prints butterflies/s for int-only, fp64-only and mixed CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ntt_device.cuh"
using namespace nttb200;

// FP64 lazy butterfly on integers stored as doubles (q < 2^50): t = w*y - rint(y*winv)*q, x' = x + t, y' = x - t.
// 6 FP64 ops for the product, 2 for the butterfly, 3 for a range fold of x' (mirrors what a real kernel needs).
struct TwD { double w, winv; };
__device__ __forceinline__ void bfly_fp64(double& x, double& y, const TwD& m, double q, double qinv, double magic)
{
  const double c = __fma_rn(y, m.winv, magic) - magic;
  const double h = y * m.w;
  const double l = __fma_rn(y, m.w, -h);
  const double d = __fma_rn(-c, q, h);
  const double t = d + l;
  double xs = x + t;
  y = x - t;
  const double k = __fma_rn(xs, qinv, magic) - magic;   // fold x' back towards (-q/2, q/2)
  x = __fma_rn(-k, q, xs);
}

template <int R, int MODE>  // MODE 0: all warps int, 1: all warps fp64, 2: odd warps fp64, 3: warps with (warp&3)==3 fp64
__global__ void __launch_bounds__(512, 1) k(uint64_t* a, const uint4* wu, const uint2* qq, ntt_cuda_params_t p, int iters_int, int iters_fp)
{
  __shared__ uint4 swu[64];
  __shared__ uint2 sqq[64];
  __shared__ TwD sd[64];
  if (threadIdx.x < 64) {
    swu[threadIdx.x] = wu[threadIdx.x]; sqq[threadIdx.x] = qq[threadIdx.x];
    const double w = (double)((uint64_t)wu[threadIdx.x].x | ((uint64_t)(wu[threadIdx.x].y & 0x1ffff) << 32));
    sd[threadIdx.x] = TwD{w, w / (double)p.q};
  }
  __syncthreads();
  constexpr int n = 1 << R;
  const int warp = threadIdx.x >> 5;
  const bool fp = MODE == 1 || (MODE == 2 && (warp & 1)) || (MODE == 3 && (warp & 3) == 3);
  if (!fp) {
    uint64_t x[n];
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = a[(size_t)blockIdx.x * 512 * n + threadIdx.x + i * 512];
    for (int it = 0; it < iters_int; it++) {
#pragma unroll
      for (int u = 0; u < R; u++) {
        const int d = n >> (u + 1);
#pragma unroll
        for (int sub = 0; sub < (1 << u); sub++) {
          uint4 A = swu[(1 << u) + sub]; uint2 B = sqq[(1 << u) + sub];
          const Mulc m{A.x, A.y, A.z, A.w, B.x, B.y};
#pragma unroll
          for (int kk = 0; kk < d; kk++) bfly_fwd<false>(x[sub * 2 * d + kk], x[sub * 2 * d + kk + d], m, p, p.c10q);
        }
      }
      if ((it & 7) == 7) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
#pragma unroll
        for (int i = 0; i < n; i++) x[i] = reduce_2q(x[i], rc);
      }
    }
#pragma unroll
    for (int i = 0; i < n; i++) a[(size_t)blockIdx.x * 512 * n + threadIdx.x + i * 512] = x[i];
  } else {
    double x[n];
    const double q = (double)p.q, qinv = 1.0 / q, magic = 6755399441055744.0;
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = (double)(a[(size_t)blockIdx.x * 512 * n + threadIdx.x + i * 512] >> 16);
    for (int it = 0; it < iters_fp; it++) {
#pragma unroll
      for (int u = 0; u < R; u++) {
        const int d = n >> (u + 1);
#pragma unroll
        for (int sub = 0; sub < (1 << u); sub++) {
          const TwD m = sd[(1 << u) + sub];
#pragma unroll
          for (int kk = 0; kk < d; kk++) bfly_fp64(x[sub * 2 * d + kk], x[sub * 2 * d + kk + d], m, q, qinv, magic);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < n; i++) a[(size_t)blockIdx.x * 512 * n + threadIdx.x + i * 512] = (uint64_t)(long long)x[i];
  }
}

template <int R, int MODE>
void run(const char* name, uint64_t* a, uint4* wu, uint2* qq, ntt_cuda_params_t p, int it_int, int it_fp, double frac_fp)
{
  const int blocks = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<R, MODE><<<blocks, 512>>>(a, wu, qq, p, 8, 8);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    k<R, MODE><<<blocks, 512>>>(a, wu, qq, p, it_int, it_fp);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double per_thread = (double)R * (1 << (R - 1));
  const double bfly = (double)blocks * 512 * per_thread * ((1.0 - frac_fp) * it_int + frac_fp * it_fp);
  printf("%-40s %7.3f ms  %8.1f Gbfly/s (int iters %d, fp iters %d)\n", name, best, bfly / (best * 1e-3) * 1e-9, it_int, it_fp);
  cudaError_t err = cudaGetLastError(); if (err != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(err));
}

int main()
{
  const uint64_t q = 0x1fffffc800001ULL;
  ntt_cuda_params_t p{};
  p.q = q; p.neg2q = 0 - 2 * q; p.negq = 0 - q; p.c10q = 10 * q; p.red_shift = 49 - 9;
  p.red_mu = (uint32_t)((((unsigned __int128)1) << (32 + p.red_shift)) / q);
  uint64_t* a; uint4* wu; uint2* qq;
  cudaMalloc(&a, (size_t)148 * 512 * 32 * 8); cudaMemset(a, 1, (size_t)148 * 512 * 32 * 8);
  cudaMalloc(&wu, 64 * 16); cudaMalloc(&qq, 64 * 8);
  uint4 hwu[64]; uint2 hqq[64];
  for (int i = 0; i < 64; i++) {
    uint64_t w = (0x123456789abcdefULL * (i + 3)) % q; uint64_t u = (uint64_t)((((unsigned __int128)w) << 32) % q);
    hwu[i] = make_uint4((uint32_t)w, (uint32_t)(w >> 32), (uint32_t)u, (uint32_t)(u >> 32));
    hqq[i] = make_uint2((uint32_t)((((unsigned __int128)w) << 30) / q), (uint32_t)((((unsigned __int128)u) << 30) / q));
  }
  cudaMemcpy(wu, hwu, sizeof(hwu), cudaMemcpyHostToDevice); cudaMemcpy(qq, hqq, sizeof(hqq), cudaMemcpyHostToDevice);
  run<4, 0>("radix16 int only", a, wu, qq, p, 512, 0, 0.0);
  run<4, 1>("radix16 fp64 only", a, wu, qq, p, 0, 512, 1.0);
  run<4, 2>("radix16 half int / half fp64 (512/512)", a, wu, qq, p, 512, 512, 0.5);
  run<4, 2>("radix16 half int / half fp64 (512/700)", a, wu, qq, p, 512, 700, 0.5);
  run<4, 3>("radix16 3/4 int, 1/4 fp64 (512/512)", a, wu, qq, p, 512, 512, 0.25);
  run<4, 3>("radix16 3/4 int, 1/4 fp64 (512/1024)", a, wu, qq, p, 512, 1024, 0.25);
  run<4, 3>("radix16 3/4 int, 1/4 fp64 (512/1536)", a, wu, qq, p, 512, 1536, 0.25);
  return 0;
}
