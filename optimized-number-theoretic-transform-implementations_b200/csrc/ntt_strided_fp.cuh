/*
 * csrc/ntt_strided_fp.cuh -- the strided passes of a transform larger than a chunk (N >= 2^15), FP64 butterflies.
 *
 * Global stages s0 .. s0+R-1 (forward order; the inverse runs them backwards) as ONE register network of 2^R values
 * per thread straight from global memory, like k_strided (ntt_kernels.cu), with the arithmetic of the FP64 ring kernel
 * (ntt_ring_fp.cuh): 16 issue slots per butterfly instead of about 30, and 2^R doubles + one twiddle pair in
 * registers instead of 2^R u64 + a 24-byte multiplier, so that radix 8 and 16 still run two adjacent groups per thread
 * with 16-byte accesses (V2) and radix 32 keeps two CTAs per SM.  The integer pass is HBM-bound only up to radix 4
 * (N = 2^16: 5.9 TB/s); at radix 8 / 16 / 32 it is bound by its own instruction count and registers (3.7 / 2.6 TB/s).
 *
 * Contract between kernels: every pass reads u64 and writes the CANONICAL residue in [0,q) --
 *   forward: input in [0,4q) (the caller's contract, src/ntt_reference.c:11-31, or a previous pass), centred to
 *            |v| <= 2q by the conversion; outputs folded and converted; the chunk kernel accepts [0,4q);
 *   inverse: input = canonical output of the FP64 chunk kernel (or of the pass before), centred to |v| <= q;
 *            LAST: the pass ends with global stage 0, whose products carry N^-1 (harvey_bkw_butterfly_final,
 *            include/internal/fast_mul_operators.h:94-106) and are below q in magnitude: converted without a fold.
 * Range schedules: FP_SCHED_STRIDED_* of ntt_fp_schedule.h (tools/gen_fp_schedule.py, re-checked on the CPU by
 * tests/test_fp64_arith_model.py).
 */
#pragma once
#include "ntt_launch.h"
#include "ntt_ring_fp.cuh"

namespace nttb200 {

/* KIND 0 forward, 1 inverse ending with the N^-1 stage, 2 inverse with more passes to follow */
template <int KIND, bool Q50, int R>
struct FpSelStrided {
  static __host__ __device__ constexpr FpPass get()
  {
    return KIND == 0 ? FP_SCHED_STRIDED_FWD[Q50][R - 1]
                     : (KIND == 1 ? FP_SCHED_STRIDED_INV[Q50][R - 1] : FP_SCHED_STRIDED_INV_NOFINAL[Q50][R - 1]);
  }
};

__host__ __device__ constexpr int strided_ilog2(int v) { return v <= 1 ? 0 : 1 + strided_ilog2(v >> 1); }

/* Thread <-> group g of a polynomial as in k_strided: es = N >> (s0+R), block i = g / es, offset j = g % es,
 * coefficients at i*(es << R) + j + k*es; V2: groups g and g+1 (same block, same twiddles), 16-byte accesses. */
/* resident CTAs the register allocation is held to: 32 values per thread -> 2 (128 registers, like pass A of the ring
 * kernel), 16 -> 3, fewer -> 4 */
#ifndef NTT_SFP_CTAS32
#define NTT_SFP_CTAS32 2
#endif
#ifndef NTT_SFP_CTAS16
#define NTT_SFP_CTAS16 3
#endif
__host__ __device__ constexpr int strided_fp_min_ctas(int values) { return values >= 32 ? NTT_SFP_CTAS32 : (values >= 16 ? NTT_SFP_CTAS16 : 4); }

template <int R, bool FWD, bool LAST, bool Q50, bool V2, bool MULTI>
__global__ void __launch_bounds__(256, strided_fp_min_ctas((V2 ? 2 : 1) << R))
  k_strided_fp(const __grid_constant__ ntt_cuda_params_t p0, const __grid_constant__ RingLimbs<MULTI> limbs,
               uint64_t *__restrict__ a, uint32_t s0, size_t n_groups)
{
  constexpr int  n = 1 << R, W = V2 ? 2 : 1;
  const uint32_t logn = p0.logn, es_log = logn - s0 - R, gl = logn - R;
  for(size_t t2 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t2 < n_groups / W; t2 += (size_t)gridDim.x * blockDim.x) {
    const size_t   t    = (size_t)W * t2;
    const size_t   poly = t >> gl;
    uint32_t       limb = 0;
    if constexpr(MULTI) limb = (uint32_t)poly / limbs.polys_per_limb;
    const ntt_cuda_params_t &p = ring_plan_of<MULTI>(p0, limbs, limb);
    const FpC      c{p.q_fd, p.qinv_fd, NTT_FP_MAGIC};
    const double   in_bias = -(4503599627370496.0 + (FWD ? 2.0 * p.q_fd : p.q_fd));
    const uint32_t g = (uint32_t)(t & (((size_t)1 << gl) - 1)), i = g >> es_log, j = g & ((1u << es_log) - 1u);
    uint64_t *     base = a + (poly << logn) + ((size_t)i << (logn - s0)) + j;
    double         x[W][n];
#pragma unroll
    for(int k = 0; k < n; k++) {
      if constexpr(V2) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(base + ((uint32_t)k << es_log));
        x[0][k]            = fp_from_u64(v.x, in_bias);
        x[1][k]            = fp_from_u64(v.y, in_bias);
      } else {
        x[0][k] = fp_from_u64(base[(uint32_t)k << es_log], in_bias);
      }
    }
    /* twiddle t = 2^u - 1 + sub of the network is entry 2^(s0+u) + i*2^u + sub of the plan's table (reference order) */
    const double2 *g_fd = (const double2 *)(FWD ? p.fwd_fd : p.inv_fd);
    auto           twf  = [&](int tt) {
      const int u = strided_ilog2(tt + 1), sub = tt + 1 - (1 << u);
      return __ldg(g_fd + (((size_t)1 << (s0 + u)) + ((size_t)i << u) + sub));
    };
#pragma unroll
    for(int w = 0; w < W; w++) {
      if constexpr(FWD) fp_network_fwd<R, FpSelStrided<0, Q50, R>>(x[w], c, twf);
      else fp_network_inv<R, LAST, FpSelStrided<LAST ? 1 : 2, Q50, R>>(x[w], c, p, twf);
    }
#pragma unroll
    for(int k = 0; k < n; k++) {
      if constexpr(V2) {
        ulonglong2 v;
        v.x = (!FWD && LAST) ? fp_to_u64(x[0][k], c, p.q) : fp_to_u64(fp_fold(x[0][k], c), c, p.q);
        v.y = (!FWD && LAST) ? fp_to_u64(x[1][k], c, p.q) : fp_to_u64(fp_fold(x[1][k], c), c, p.q);
        *reinterpret_cast<ulonglong2 *>(base + ((uint32_t)k << es_log)) = v;
      } else {
        base[(uint32_t)k << es_log] = (!FWD && LAST) ? fp_to_u64(x[0][k], c, p.q) : fp_to_u64(fp_fold(x[0][k], c), c, p.q);
      }
    }
  }
}

}  // namespace nttb200
