/* The one-kernel negacyclic multiply (ntt_polymul_fp.cuh): instantiation and launcher. */
#include "ntt_launch.h"
#include "ntt_polymul_fp.cuh"

namespace nttb200 {

template <bool Q50>
static int polymul_launch_one(int device, const ntt_cuda_params_t &p, const CUtensorMap &ta, const CUtensorMap &tb,
                              unsigned grid, size_t n_pairs, uint64_t *d_c, cudaStream_t st)
{
  auto        kern      = k_polymul_fp<Q50>;
  static bool ready[64] = {false};
  if(!ready[device & 63]) {
    NL_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PolymulCfg::SMEM));
    ready[device & 63] = true;
  }
  kern<<<grid, PolymulCfg::T, PolymulCfg::SMEM, st>>>(p, ta, tb, n_pairs, d_c);
  NL_CU(cudaGetLastError());
  return 0;
}

/* c[k] = a[k] * b[k] in Z_q[X]/(X^N+1) for k < n_pairs; N = 2^13, FP64-eligible modulus, a and b 128-byte aligned */
int polymul_fp_launch(int device, const ntt_cuda_params_t &p, uint64_t *d_a, uint64_t *d_b, uint64_t *d_c,
                      size_t n_pairs, cudaStream_t st)
{
  CUtensorMap ta, tb;
  if(nl_make_block_tmap(&ta, d_a, n_pairs << PolymulCfg::L, 32 * PolymulCfg::BOXB)) return -1;
  if(nl_make_block_tmap(&tb, d_b, n_pairs << PolymulCfg::L, 32 * PolymulCfg::BOXB)) return -1;
  size_t grid = (size_t)nl_sm_count(device);
  if(grid > n_pairs) grid = n_pairs;
  return p.fp64 == 2 ? polymul_launch_one<true>(device, p, ta, tb, (unsigned)grid, n_pairs, d_c, st)
                     : polymul_launch_one<false>(device, p, ta, tb, (unsigned)grid, n_pairs, d_c, st);
}

}  // namespace nttb200
