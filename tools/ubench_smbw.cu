// How much HBM bandwidth can K SMs move?  K CTAs (one per SM: each asks for 200 KB of shared memory), 1024 threads,
// every thread reads U x 16 bytes at a large stride (like a strided NTT pass), adds one, writes back in place.
// Decides whether a few SMs could run the strided pass of an N >= 2^15 transform beside the ring kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_smbw tools/ubench_smbw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int U>
__global__ void __launch_bounds__(1024, 1) k(ulonglong2 *a, size_t n16, size_t stride16)
{
  extern __shared__ char dummy[];
  // groups of U elements at distance stride16; consecutive threads take consecutive 16-byte words
  const size_t groups = n16 / U;
  for(size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const size_t blk = g / stride16, j = g % stride16;
    ulonglong2 *base = a + blk * stride16 * U + j;
    ulonglong2  x[U];
#pragma unroll
    for(int i = 0; i < U; i++) x[i] = base[(size_t)i * stride16];
#pragma unroll
    for(int i = 0; i < U; i++) { x[i].x += 1; x[i].y += x[(i + 1) % U].x; }
#pragma unroll
    for(int i = 0; i < U; i++) base[(size_t)i * stride16] = x[i];
  }
}
int main()
{
  const size_t bytes = (size_t)1 << 30, n16 = bytes / 16;
  ulonglong2 *d; cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int Ks[] = {8, 16, 24, 32, 48, 74, 148};
  for(int u = 4; u <= 8; u += 4)
    for(int K : Ks) {
      for(int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if(u == 4) k<4><<<K, 1024, 200 * 1024>>>(d, n16, 8192); else k<8><<<K, 1024, 200 * 1024>>>(d, n16, 8192);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("U=%d K=%3d CTAs(SMs): %.3f ms  %.0f GB/s total (read+write)  %.1f GB/s per SM\n", u, K, ms, 2.0 * bytes / ms / 1e6, 2.0 * bytes / ms / 1e6 / K);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
