/* csrc/ntt_ring_fp_launch.inl -- body shared by ntt_ring_fp_{12,13,14}.cu; NTT_RING_L selects the chunk size. */
#include "ntt_launch.h"
#include "ntt_ring_fp.cuh"

namespace nttb200 {

template <int L, bool FWD, int MODE, bool Q50>
static int ring_fp_launch_one(int device, const ntt_cuda_params_t &p, const CUtensorMap &tm, const CUtensorMap &tm2,
                              unsigned grid, uint64_t *d_a, size_t n_chunks, cudaStream_t st, const RingOpts &o)
{
  using C               = RingCfg<L>;
  auto        kern      = k_ring_fp<L, FWD, MODE, Q50>;
  static bool ready[64] = {false};
  if(!ready[device & 63]) {
    NL_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_FP));
    ready[device & 63] = true;
  }
  kern<<<grid, C::THREADS, C::SMEM_FP, st>>>(p, tm, tm2, n_chunks, d_a, o.d_other, o.other_mask, RingLimbs<false>{0, 1});
  NL_CU(cudaGetLastError());
  return 0;
}

#define NTT_RING_CAT2(a, b) a##b
#define NTT_RING_CAT(a, b) NTT_RING_CAT2(a, b)

#if defined(NTT_EXPERIMENT) && NTT_RING_L != 14
/* -DNTT_EXPERIMENT: scratch builds for kernel A/B timing (tools/exp_build.sh) carry the L = 14 kernels only */
int NTT_RING_CAT(ring_fp_launch_, NTT_RING_L)(bool, int, const ntt_cuda_params_t &, uint64_t *, size_t, cudaStream_t,
                                              const RingOpts &)
{
  return nl_fail_msg("experiment build: only the L = 14 FP64 ring kernels are compiled");
}
#else
int NTT_RING_CAT(ring_fp_launch_, NTT_RING_L)(bool fwd, int device, const ntt_cuda_params_t &p, uint64_t *d_a,
                                              size_t n_chunks, cudaStream_t st, const RingOpts &o)
{
  constexpr int L = NTT_RING_L;
  using C         = RingCfg<L>;
  CUtensorMap tm, tm2;
  if(nl_make_block_tmap(&tm, d_a, n_chunks << L, 32)) return -1;
  if(nl_make_block_tmap(&tm2, d_a, n_chunks << L, 32 * C::BOXB)) return -1; /* BOXB blocks per box */
  size_t grid = (size_t)nl_sm_count(device) * C::CTAS;
  if(grid > n_chunks) grid = n_chunks;
  const size_t mc = nl_min_chunks_per_cta();
  if(mc && grid * mc > n_chunks) grid = (n_chunks + mc - 1) / mc;
  const bool     q50 = p.fp64 == 2; /* 50-bit range schedule */
  const unsigned g   = (unsigned)grid;
#ifdef NTT_EXPERIMENT
  (void)q50;
  return fwd ? ring_fp_launch_one<L, true, RING_PLAIN, false>(device, p, tm, tm2, g, d_a, n_chunks, st, o)
             : ring_fp_launch_one<L, false, RING_PLAIN, false>(device, p, tm, tm2, g, d_a, n_chunks, st, o);
#else
  if(!fwd) {
    return q50 ? ring_fp_launch_one<L, false, RING_PLAIN, true>(device, p, tm, tm2, g, d_a, n_chunks, st, o)
               : ring_fp_launch_one<L, false, RING_PLAIN, false>(device, p, tm, tm2, g, d_a, n_chunks, st, o);
  }
  if(o.d_other) {
    return q50 ? ring_fp_launch_one<L, true, RING_MUL, true>(device, p, tm, tm2, g, d_a, n_chunks, st, o)
               : ring_fp_launch_one<L, true, RING_MUL, false>(device, p, tm, tm2, g, d_a, n_chunks, st, o);
  }
  if(o.lazy_out) {
    return q50 ? ring_fp_launch_one<L, true, RING_LAZY, true>(device, p, tm, tm2, g, d_a, n_chunks, st, o)
               : ring_fp_launch_one<L, true, RING_LAZY, false>(device, p, tm, tm2, g, d_a, n_chunks, st, o);
  }
  return q50 ? ring_fp_launch_one<L, true, RING_PLAIN, true>(device, p, tm, tm2, g, d_a, n_chunks, st, o)
             : ring_fp_launch_one<L, true, RING_PLAIN, false>(device, p, tm, tm2, g, d_a, n_chunks, st, o);
#endif
}
#endif

}  // namespace nttb200

#if NTT_RING_L == 14 && !defined(NTT_EXPERIMENT)
namespace nttb200 {
/* One launch over the chunks of several plans (RNS limbs), see k_ring_fp<..., MULTI>.  All plans share N and the
 * range schedule; limb l owns polys_per_limb consecutive polynomials of d_a. */
template <bool FWD, bool Q50>
static int ring_fp_launch_multi_one(int device, const ntt_cuda_params_t *const *plist, size_t n_limbs,
                                    size_t polys_per_limb, const CUtensorMap &tm, const CUtensorMap &tm2, unsigned grid,
                                    uint64_t *d_a, size_t n_chunks, cudaStream_t st)
{
  using C               = RingCfg<14>;
  auto        kern      = k_ring_fp<14, FWD, RING_PLAIN, Q50, true>;
  static bool ready[64] = {false};
  if(!ready[device & 63]) {
    NL_CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_FP));
    ready[device & 63] = true;
  }
  static thread_local RingLimbs<true> lb; /* 21 KiB: kept off the stack */
  for(size_t l = 0; l < n_limbs; l++) lb.e[l] = *plist[l];
  lb.polys_per_limb = (uint32_t)polys_per_limb;
  lb.ctas_per_limb  = grid / (unsigned)n_limbs;
  kern<<<dim3(lb.ctas_per_limb, (unsigned)n_limbs), C::THREADS, C::SMEM_FP, st>>>(*plist[0], tm, tm2, n_chunks, d_a, nullptr, ~(size_t)0, lb);
  NL_CU(cudaGetLastError());
  return 0;
}

int ring_fp_launch_multi_14(bool fwd, int device, const ntt_cuda_params_t *const *plist, size_t n_limbs,
                            size_t polys_per_limb, uint64_t *d_a, cudaStream_t st)
{
  using C = RingCfg<14>;
  if(n_limbs < 1 || n_limbs > (size_t)RING_MAX_LIMBS) return nl_fail_msg("multi-plan launch: 1..48 plans");
  const ntt_cuda_params_t &p  = *plist[0];
  const size_t             n_chunks = (n_limbs * polys_per_limb) << (p.logn - 14);
  if(n_chunks >= ((size_t)1 << 31) || polys_per_limb >= ((size_t)1 << 31)) return nl_fail_msg("multi-plan launch: batch too large");
  CUtensorMap tm, tm2;
  if(nl_make_block_tmap(&tm, d_a, n_chunks << 14, 32)) return -1;
  if(nl_make_block_tmap(&tm2, d_a, n_chunks << 14, 32 * C::BOXB)) return -1;
  /* a CTA serves one limb: as many CTAs per limb as the SMs allow (at least one, at most one per chunk) */
  size_t cpl = (size_t)nl_sm_count(device) / n_limbs;
  if(cpl < 1) cpl = 1;
  if(cpl > n_chunks / n_limbs) cpl = n_chunks / n_limbs;
  const size_t grid = cpl * n_limbs;
  const bool   q50  = p.fp64 == 2;
  if(fwd) {
    return q50 ? ring_fp_launch_multi_one<true, true>(device, plist, n_limbs, polys_per_limb, tm, tm2, (unsigned)grid, d_a, n_chunks, st)
               : ring_fp_launch_multi_one<true, false>(device, plist, n_limbs, polys_per_limb, tm, tm2, (unsigned)grid, d_a, n_chunks, st);
  }
  return q50 ? ring_fp_launch_multi_one<false, true>(device, plist, n_limbs, polys_per_limb, tm, tm2, (unsigned)grid, d_a, n_chunks, st)
             : ring_fp_launch_multi_one<false, false>(device, plist, n_limbs, polys_per_limb, tm, tm2, (unsigned)grid, d_a, n_chunks, st);
}
}  // namespace nttb200
#elif NTT_RING_L == 14
namespace nttb200 {
int ring_fp_launch_multi_14(bool, int, const ntt_cuda_params_t *const *, size_t, size_t, uint64_t *, cudaStream_t)
{
  return nl_fail_msg("experiment build: the multi-plan launch is not compiled");
}
}  // namespace nttb200
#endif

#if defined(NTT_RING_TRACE) && NTT_RING_L == 14
/* -DNTT_RING_TRACE builds only: phase timestamps of CTA 0 (tools/trace_phases.py) */
extern "C" int ntt_cuda_trace_read(long long *out, size_t n)
{
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, nttb200::g_trace, n * sizeof(long long));
  return 0;
}
#endif
