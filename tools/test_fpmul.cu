// brute-force check of fp_mul / fp_mul_wide / fp_fold against exact integer arithmetic
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ntt_ring_fp.cuh"
using namespace nttb200;
typedef unsigned __int128 u128;
__device__ uint64_t splitmix(uint64_t& s){ uint64_t z=(s+=0x9e3779b97f4a7c15ULL); z=(z^(z>>30))*0xbf58476d1ce4e5b9ULL; z=(z^(z>>27))*0x94d049bb133111ebULL; return z^(z>>31); }
__global__ void k(uint64_t q, unsigned long long* bad, double* ex, int range_mult)
{
  const FpC c{(double)q, 1.0/(double)q, NTT_FP_MAGIC};
  uint64_t s = 0x1234 + blockIdx.x * 7919ull + threadIdx.x * 104729ull;
  for (int it = 0; it < 2000; it++) {
    const uint64_t w = splitmix(s) % q;
    const uint64_t ymag = splitmix(s) % ((uint64_t)range_mult * q);
    const bool neg = splitmix(s) & 1;
    const double y = neg ? -(double)ymag : (double)ymag;
    const double wd = (double)w, winv = __ddiv_rn(wd, c.q);
    const double t = (it & 1) ? fp_mul_wide(y, wd, winv, c) : ((range_mult <= 3) ? fp_mul(y, wd, winv, c) : fp_mul_wide(y, wd, winv, c));
    // exact: (w * ymag) mod q with sign
    uint64_t r = (uint64_t)(((u128)w * ymag) % q);
    if (neg && r) r = q - r;
    // t mod q
    const double tf = fp_fold(t, c);
    long long ti = (long long)tf; if (ti < 0) ti += (long long)q;
    const bool integral = (t == floor(t)) && (tf == floor(tf));
    if (!integral || (uint64_t)ti != r || fabs(t) > 2.0 * c.q) {
      unsigned long long idx = atomicAdd(bad, 1ull);
      if (idx < 4) { ex[idx*4+0] = y; ex[idx*4+1] = wd; ex[idx*4+2] = t; ex[idx*4+3] = (double)r; }
    }
  }
}
int main(){
  unsigned long long* bad; double* ex; cudaMallocManaged(&bad, 8); cudaMallocManaged(&ex, 16*8);
  const uint64_t q = 0x1fffffc800001ULL;
  for (int rm : {1, 2, 3, 4, 8}) {
    *bad = 0; k<<<148*4, 256>>>(q, bad, ex, rm); cudaDeviceSynchronize();
    printf("range +-%dq: bad %llu", rm, *bad);
    if (*bad) printf("  e.g. y=%.1f w=%.1f t=%.1f expect %.1f", ex[0], ex[1], ex[2], ex[3]);
    printf("\n");
  }
  return 0;
}
