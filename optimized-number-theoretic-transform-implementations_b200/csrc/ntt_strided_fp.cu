/* csrc/ntt_strided_fp.cu -- instantiation and launch of the FP64 strided passes (ntt_strided_fp.cuh). */
#include "ntt_strided_fp.cuh"

namespace nttb200 {

template <int R, bool FWD, bool LAST, bool Q50, bool V2, bool MULTI>
static int strided_fp_launch_one(int device, const ntt_cuda_params_t &p0, const RingLimbs<MULTI> &lb, uint64_t *d_a,
                                 uint32_t s0, size_t n_groups, cudaStream_t st)
{
  size_t       grid = (n_groups / (V2 ? 2 : 1) + 255) / 256;
  const size_t cap  = (size_t)nl_sm_count(device) * 32;
  if(grid > cap) grid = cap;
  k_strided_fp<R, FWD, LAST, Q50, V2, MULTI><<<(unsigned)grid, 256, 0, st>>>(p0, lb, d_a, s0, n_groups);
  NL_CU(cudaGetLastError());
  return 0;
}

template <int R, bool MULTI>
static int strided_fp_launch_r(bool fwd, bool last, int device, const ntt_cuda_params_t &p0, const RingLimbs<MULTI> &lb,
                               uint64_t *d_a, uint32_t s0, size_t n_groups, cudaStream_t st)
{
  /* two adjacent groups per thread where the group stride is even, the data 16-byte aligned and 2 * 2^R doubles
   * still leave room for two CTAs per SM */
  const uint32_t es_log = p0.logn - s0 - R;
  const bool     v2     = R <= 4 && es_log >= 1 && (n_groups & 1) == 0 && ((uintptr_t)d_a & 15) == 0;
  const bool     q50    = p0.fp64 == 2;
#define NTT_SF(F, LA, Q, V) return strided_fp_launch_one<R, F, LA, Q, V, MULTI>(device, p0, lb, d_a, s0, n_groups, st)
  if constexpr(R <= 4) {
    if(v2) {
      if(fwd) { if(q50) NTT_SF(true, false, true, true); else NTT_SF(true, false, false, true); }
      if(last) { if(q50) NTT_SF(false, true, true, true); else NTT_SF(false, true, false, true); }
      if(q50) NTT_SF(false, false, true, true); else NTT_SF(false, false, false, true);
    }
  }
  if(fwd) { if(q50) NTT_SF(true, false, true, false); else NTT_SF(true, false, false, false); }
  if(last) { if(q50) NTT_SF(false, true, true, false); else NTT_SF(false, true, false, false); }
  if(q50) NTT_SF(false, false, true, false); else NTT_SF(false, false, false, false);
#undef NTT_SF
}

template <bool MULTI>
static int strided_fp_dispatch(bool fwd, bool last, int R, int device, const ntt_cuda_params_t &p0,
                               const RingLimbs<MULTI> &lb, uint64_t *d_a, uint32_t s0, size_t n_groups, cudaStream_t st)
{
  switch(R) {
    case 1: return strided_fp_launch_r<1, MULTI>(fwd, last, device, p0, lb, d_a, s0, n_groups, st);
    case 2: return strided_fp_launch_r<2, MULTI>(fwd, last, device, p0, lb, d_a, s0, n_groups, st);
    case 3: return strided_fp_launch_r<3, MULTI>(fwd, last, device, p0, lb, d_a, s0, n_groups, st);
    case 4: return strided_fp_launch_r<4, MULTI>(fwd, last, device, p0, lb, d_a, s0, n_groups, st);
    case 5: return strided_fp_launch_r<5, MULTI>(fwd, last, device, p0, lb, d_a, s0, n_groups, st);
    default: return nl_fail_msg("unsupported strided radix");
  }
}

int strided_fp_launch(bool fwd, bool last, int R, int device, const ntt_cuda_params_t &p, uint64_t *d_a, uint32_t s0,
                      size_t batch, cudaStream_t st)
{
  return strided_fp_dispatch<false>(fwd, last, R, device, p, RingLimbs<false>{0, 1}, d_a, s0, batch << (p.logn - R), st);
}

int strided_fp_launch_multi(bool fwd, bool last, int R, int device, const RingLimbs<true> &lb, uint64_t *d_a, uint32_t s0,
                            size_t total_polys, cudaStream_t st)
{
  return strided_fp_dispatch<true>(fwd, last, R, device, lb.e[0], lb, d_a, s0, total_polys << (lb.e[0].logn - R), st);
}

}  // namespace nttb200
