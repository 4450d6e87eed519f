// How fast does one register-tiled pass of the FP64 ring kernel run in isolation?  Shared memory -> registers ->
// R butterfly stages -> shared memory, no TMA, no block barrier: isolates the LDS/FP64/STS interleaving at a given
// occupancy.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../optimized-number-theoretic-transform-implementations_b200/csrc -o ubench_pass ubench_pass.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "ntt_ring_fp.cuh"
using namespace nttb200;

// MODE 0: pass-B shape (32 values at stride 128 B, swizzled), THREADS threads, UNITS units per thread per iteration
template <int R, int THREADS, int MINB, bool SYNC>
__global__ void __launch_bounds__(THREADS, MINB) k_pass(ntt_cuda_params_t p, int iters, double* out)
{
  extern __shared__ uint8_t smem[];
  constexpr int n = 1 << R;
  double2* tw_s = reinterpret_cast<double2*>(smem);                 // 1024 twiddles (16 KiB)
  uint8_t* data = smem + 16384;                                     // THREADS * n doubles
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 1024; i += THREADS) tw_s[i] = make_double2((double)(p.q - 12345 - i), (double)(p.q - 12345 - i) / (double)p.q);
  for (int i = tid; i < THREADS * n; i += THREADS) reinterpret_cast<double*>(data)[i] = (double)(i * 977 % 100003);
  __syncthreads();
  const FpC c{p.q_fd, p.qinv_fd, NTT_FP_MAGIC};
  // a warp owns 32*n contiguous doubles; lane owns the values at stride 32 (conflict-free)
  double* base = reinterpret_cast<double*>(data) + (size_t)warp * 32 * n + lane;
  const double2* tw = tw_s + (warp & 15) * (n - 1);
  for (int it = 0; it < iters; it++) {
    double x[n];
#pragma unroll
    for (int k = 0; k < n; k++) x[k] = base[k * 32];
    fp_network<R, true, false, false, 0>(x, c, p, [&](int t) { return tw[t]; });
#pragma unroll
    for (int k = 0; k < n; k++) base[k * 32] = x[k];
    if (SYNC) __syncthreads(); else __syncwarp();
  }
  if (out) out[blockIdx.x * THREADS + tid] = base[0];
}

template <int R, int THREADS, int MINB, bool SYNC>
void run(const char* name, ntt_cuda_params_t p)
{
  const int smem = 16384 + THREADS * (1 << R) * 8;
  cudaFuncSetAttribute(k_pass<R, THREADS, MINB, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000, blocks = 148 * MINB;
  k_pass<R, THREADS, MINB, SYNC><<<blocks, THREADS, smem>>>(p, 10, nullptr);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k_pass<R, THREADS, MINB, SYNC><<<blocks, THREADS, smem>>>(p, iters, nullptr); cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  // FP64 instructions per unit: fold 3n + R * n/2 * 8
  const double fp = 3.0 * (1 << R) + R * (1 << (R - 1)) * 8.0;
  const double warp_instr = fp * iters * (THREADS / 32) * MINB / 4.0;  // per SMSP
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("%-44s %7.3f ms  %5.3f FP64 warp-instr/clk/SMSP (%4.1f%% of 0.5)  %s\n", name, ms, warp_instr / cyc, 200.0 * warp_instr / cyc,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main()
{
  ntt_cuda_params_t p{};
  p.q = 0x1fffffc800001ULL; p.q_fd = (double)p.q; p.qinv_fd = 1.0 / (double)p.q; p.logn = 14;
  run<5, 512, 1, false>("32 values, 512 thr (4 warps/SMSP), syncwarp", p);
  run<5, 512, 1, true>("32 values, 512 thr, syncthreads per unit", p);
  run<5, 256, 1, false>("32 values, 256 thr (2 warps/SMSP)", p);
  run<4, 512, 1, false>("16 values, 512 thr (4 warps/SMSP)", p);
  run<4, 1024, 1, false>("16 values, 1024 thr (8 warps/SMSP)", p);
  run<4, 1024, 1, true>("16 values, 1024 thr, syncthreads per unit", p);
  run<4, 512, 2, false>("16 values, 2 x 512 thr (8 warps/SMSP)", p);
  run<3, 1024, 2, false>("8 values, 2 x 1024 thr (16 warps/SMSP)", p);
  return 0;
}
