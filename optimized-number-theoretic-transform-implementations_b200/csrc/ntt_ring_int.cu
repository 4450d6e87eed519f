/* Integer (lazy split multiplier) ring kernels, chunks of 2^12 .. 2^14: what 2^50 <= q < 2^56 gets (ntt_ring.cuh). */
#include "ntt_launch.h"
#include "ntt_ring.cuh"

namespace nttb200 {

template <int L, bool FWD>
static int ring_int_launch_one(int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks, cudaStream_t st)
{
  using C = RingCfg<L>;
  CUtensorMap tm;
  if(nl_make_block_tmap(&tm, d_a, n_chunks << L, 32)) return -1;
  size_t grid = (size_t)nl_sm_count(device) * C::CTAS;
  if(grid > n_chunks) grid = n_chunks;
  const size_t mc = nl_min_chunks_per_cta();
  if(mc && grid * mc > n_chunks) grid = (n_chunks + mc - 1) / mc;
  auto        kern      = k_ring<L, FWD>;
  static bool ready[64] = {false};
  if(!ready[device & 63]) {
    NL_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    ready[device & 63] = true;
  }
  kern<<<(unsigned)grid, C::THREADS, C::SMEM, st>>>(p, tm, n_chunks, d_a);
  NL_CU(cudaGetLastError());
  return 0;
}

int ring_int_launch(int L, bool fwd, int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks,
                    cudaStream_t st)
{
  switch(L) {
    case 12: return fwd ? ring_int_launch_one<12, true>(device, p, d_a, n_chunks, st)
                        : ring_int_launch_one<12, false>(device, p, d_a, n_chunks, st);
    case 13: return fwd ? ring_int_launch_one<13, true>(device, p, d_a, n_chunks, st)
                        : ring_int_launch_one<13, false>(device, p, d_a, n_chunks, st);
    case 14: return fwd ? ring_int_launch_one<14, true>(device, p, d_a, n_chunks, st)
                        : ring_int_launch_one<14, false>(device, p, d_a, n_chunks, st);
  }
  return nl_fail_msg("unsupported ring chunk size");
}

}  // namespace nttb200
