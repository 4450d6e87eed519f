/*
 * csrc/ntt_ring_fp.cuh -- the ring kernel with the butterflies on the FP64 pipe (q <= 2^50 - 2048).
 *
 * Why: on B200 the integer butterfly of ntt_device.cuh costs about 30 issue slots per warp (five 32x32->64
 * products dominate), the exact FP64 formulation below 8 FP64 instructions = 16 issue slots (an FP64 instruction
 * holds the scheduler's dispatch port for two cycles and nothing co-issues with it: tools/ubench_rf.cu,
 * profiles/r02_ubench_rf.txt).  Integer and FP64 work do not overlap, so the whole network runs in FP64.
 *
 * Exactness.  Coefficients are integers held in doubles, signed, |v| < 2^53.  Twiddles are stored centred, w in
 * (-q/2, q/2), with winv = RN(w/q) (|winv| <= 1/2, within 2^-55 of w/q).  For any integer y:
 *     c = (y*winv + M) - M                  (one fused rounding of y*w/q: to an integer, M = 1.5*2^52, |y*winv| < 2^51,
 *                                            i.e. |y| < 2^52 -- "plain" -- or to an EVEN integer, M = 3*2^52,
 *                                            |y*winv| < 2^52, i.e. |y| < 2^53 -- "coarse")
 *     h = RN(w*y),  l = fma(w, y, -h)       (error-free product: w*y = h + l exactly)
 *     d = fma(-c, q, h)                     (exact: |h - c*q| < 2^53)
 *     t = d + l                             (exact)  =>  t = w*y - c*q == w*y (mod q), an integer
 * so the residue is exact whatever c is; rounding only decides how large |t| gets:
 *     plain  |t| <= q*(1/2 + |y|*2^-55)          coarse  |t| <= q*(1 + |y|*2^-55)
 * (a multiplier that is NOT centred -- the data operand of the fused pointwise product -- keeps the older bounds:
 *  |y*winv| < 2^51 needs |y| < 2^51, error term |y|*2^-54 or, with winv rounded on the fly, |y|*2^-53)
 * Butterflies are X' = X + t, Y' = X - t (forward) and X' = X + Y, Y' = t(X - Y) (inverse) with no range
 * correction; a fold (v - rint(v/q)*q, 3 instructions) brings a value back to |v| <= q/2 + 6.  WHERE values are
 * folded and WHICH rounding a product uses is not written here: tools/gen_fp_schedule.py derives it with exact
 * rational bound propagation -- per stage for the forward transform, per POSITION of the register network for the
 * inverse, whose sums double while its products come back small -- and writes the tables of ntt_fp_schedule.h that
 * the networks below are instantiated from.  tests/test_fp64_arith_model.py re-checks the tables and the primitives
 * on the CPU (exact FMA emulation), tests/test_gpu_soak.py compares every row of 64 x 4096-polynomial batches per
 * modulus class and direction with the CPU restatement of the reference.
 * The last pass folds, adds q to negatives and converts back to u64: the output is the canonical residue in
 * [0,q), bit-identical to fwd_ntt_ref_harvey / inv_ntt_ref_harvey (include/ntt_reference.h:19-31,
 * src/ntt_reference.c:33-66).
 *
 * Data path, ring, TMA, swizzle, pass structure: identical to ntt_ring.cuh (see there).  Between passes the
 * slots hold doubles instead of u64.
 */
#pragma once
#include "ntt_fp_schedule.h"
#include "ntt_ring.cuh"

namespace nttb200 {

struct FpC {
  double q, qinv, magic;
};
#define NTT_FP_MAGIC 6755399441055744.0   /* 1.5 * 2^52: rounds |arg| < 2^51 to the nearest integer */
#define NTT_FP_MAGIC2 13510798882111488.0 /* 3 * 2^52:   rounds |arg| < 2^52 to the nearest EVEN integer */

/* v - rint(v/q)*q: |result| <= q/2 + 6 for |v| < 2^53 */
__device__ __forceinline__ double fp_fold(double v, const FpC &c)
{
  const double k = __dadd_rn(__fma_rn(v, c.qinv, c.magic), -c.magic);
  return __fma_rn(-k, c.q, v);
}
/* t == w*y (mod q), exact integer.  COARSE = false: |y*winv| < 2^51, |t| <= q*(1/2 + |y|*2^-55) for a centred table
 * multiplier; COARSE = true: |y*winv| < 2^52, the quotient is rounded to an even integer, |t| <= q*(1 + |y|*2^-55).
 * Six FP64 instructions either way (tools/gen_fp_schedule.py decides which one a butterfly gets). */
template <bool COARSE = false>
__device__ __forceinline__ double fp_mul(double y, double w, double winv, const FpC &c)
{
  constexpr double M  = COARSE ? NTT_FP_MAGIC2 : NTT_FP_MAGIC;
  const double     cc = __dadd_rn(__fma_rn(y, winv, M), -M);
  const double     h  = __dmul_rn(y, w);
  const double     l  = __fma_rn(y, w, -h);
  const double     d  = __fma_rn(-cc, c.q, h);
  return __dadd_rn(d, l);
}
/* u64 below 2^52 -> the same integer as a double (exponent splice + one DADD) */
__device__ __forceinline__ double fp_from_u64(uint64_t v, double neg_bias = -4503599627370496.0)
{
  /* neg_bias = -(2^52 + k): the result is v - k, exact as long as 2^52 + k is a double (k < 2^52) */
  return __dadd_rn(__hiloint2double((int)(hi32(v) | 0x43300000u), (int)lo32(v)), neg_bias);
}
/* |v| < q, integer  ->  canonical residue in [0,q) as u64.  One DADD puts v next to 1.5*2^52, where consecutive
 * integers are consecutive bit patterns; the rest is integer work: subtract the constant's bits and add q to
 * negatives (v < 0 <=> the high word is below the constant's, whose low word is 0).  Written in PTX so that it
 * stays at compare + two selects + add-with-carry. */
__device__ __forceinline__ uint64_t fp_to_u64(double v, const FpC &c, uint64_t q)
{
  const double   t   = __dadd_rn(v, c.magic);
  const uint32_t hi  = (uint32_t)__double2hiint(t), lo = (uint32_t)__double2loint(t);
  const uint32_t qlo = lo32(q), qhi = hi32(q);
  uint32_t       rlo, rhi;
  /* (a predicated 64-bit add of q instead of the two selects is one instruction shorter and measured 1.7 % slower) */
  asm("{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 alo, ahi;\n\t"
      "setp.lt.u32 p, %2, 0x43380000;\n\t"
      "selp.u32 alo, %4, 0, p;\n\t"
      "selp.u32 ahi, %5, 0xbcc80000, p;\n\t" /* -0x43380000 (+ q's high word for negatives) */
      "add.cc.u32 %0, %3, alo;\n\t"
      "addc.u32 %1, %2, ahi;\n\t"
      "}"
      : "=r"(rlo), "=r"(rhi)
      : "r"(hi), "r"(lo), "r"(qlo), "r"(qhi + 0xbcc80000u));
  return ((uint64_t)rhi << 32) | rlo;
}
/* |v| < q, integer -> v + q in (0, 2q) as u64: the lazy representative, no sign test (bias = 1.5*2^52 + q) */
__device__ __forceinline__ uint64_t fp_to_u64_lazy(double v, double bias)
{
  const double   t  = __dadd_rn(v, bias);
  const uint32_t hi = (uint32_t)__double2hiint(t), lo = (uint32_t)__double2loint(t);
  return ((uint64_t)(hi - 0x43380000u) << 32) | lo;
}

/*
 * Register networks.  Which positions are folded before a stage, which butterflies round coarsely and which
 * positions are folded after the last stage come from the generated tables of ntt_fp_schedule.h
 * (tools/gen_fp_schedule.py: exact-rational bound propagation, re-checked by tests/test_fp64_arith_model.py).
 *   forward (X' = X + t, Y' = X - t): all positions share one bound, so a fold is all-or-nothing per stage;
 *   inverse (X' = X + Y, Y' = t(X - Y)): sums double but products come back small, so only the few positions
 *   whose bound would break a limit are folded.
 * KIND 0 forward, 1 inverse (pass A ends with the N^-1 stage), 2 inverse of a chunk of a larger transform.
 * WHICH 0 / 1 / 2 = pass A / B / C.
 */
template <int KIND, bool Q50, int L, int WHICH>
struct FpSel {
  static __host__ __device__ constexpr FpPass get()
  {
    const FpSchedule &s = KIND == 0 ? FP_SCHED_FWD[Q50][L - 10]
                                    : (KIND == 1 ? FP_SCHED_INV[Q50][L - 10] : FP_SCHED_INV_NOFINAL[Q50][L - 10]);
    return WHICH == 0 ? s.a : (WHICH == 1 ? s.b : s.c);
  }
};
struct FpNoHook {
  __device__ __forceinline__ void operator()() const {}
};

template <int R, uint32_t MASK>
__device__ __forceinline__ void fp_fold_mask(double (&x)[1 << R], const FpC &c)
{
#pragma unroll
  for(int k = 0; k < (1 << R); k++) {
    if((MASK >> k) & 1u) x[k] = fp_fold(x[k], c);
  }
}

/* forward stage u of an R-stage network: pairs at distance n >> (u+1), sub-block `sub` uses twiddle 2^u-1+sub */
template <int R, int U, uint32_t COARSE, typename TWF>
__device__ __forceinline__ void fp_fwd_stage(double (&x)[1 << R], const FpC &c, TWF &twf)
{
  constexpr int n = 1 << R, d = n >> (U + 1);
#pragma unroll
  for(int sub = 0; sub < (1 << U); sub++) {
    const double2 tw = twf((1 << U) - 1 + sub);
#pragma unroll
    for(int k = 0; k < d; k++) {
      const int    lo = sub * 2 * d + k;
      const double t  = ((COARSE >> lo) & 1u) ? fp_mul<true>(x[lo + d], tw.x, tw.y, c) : fp_mul<false>(x[lo + d], tw.x, tw.y, c);
      x[lo + d]       = __dadd_rn(x[lo], -t);
      x[lo]           = __dadd_rn(x[lo], t);
    }
  }
}

/* MIDF: called once after the second stage (the forward kernel re-arms a ring slot there) */
template <int R, typename SEL, typename TWF, typename MIDF = FpNoHook>
__device__ __forceinline__ void fp_network_fwd(double (&x)[1 << R], const FpC &c, TWF twf, MIDF mid = MIDF())
{
  constexpr FpPass S = SEL::get();
  fp_fold_mask<R, S.fold_before[0]>(x, c);
  fp_fwd_stage<R, 0, S.coarse[0]>(x, c, twf);
  if constexpr(R > 1) {
    fp_fold_mask<R, S.fold_before[1]>(x, c);
    fp_fwd_stage<R, 1, S.coarse[1]>(x, c, twf);
  }
  mid();
  if constexpr(R > 2) {
    fp_fold_mask<R, S.fold_before[2]>(x, c);
    fp_fwd_stage<R, 2, S.coarse[2]>(x, c, twf);
  }
  if constexpr(R > 3) {
    fp_fold_mask<R, S.fold_before[3]>(x, c);
    fp_fwd_stage<R, 3, S.coarse[3]>(x, c, twf);
  }
  if constexpr(R > 4) {
    fp_fold_mask<R, S.fold_before[4]>(x, c);
    fp_fwd_stage<R, 4, S.coarse[4]>(x, c, twf);
  }
  fp_fold_mask<R, S.fold_end>(x, c);
}

/* inverse stage number ST in processing order (network stage u = R-1-ST): pairs at distance 2^ST.
 * LAST: global stage 0, both outputs are products carrying N^-1 (harvey_bkw_butterfly_final,
 * fast_mul_operators.h:94-106); the schedule guarantees plain rounding there, so |t| < q and the results are
 * converted without another fold. */
template <int R, int ST, uint32_t COARSE, bool LAST, typename TWF>
__device__ __forceinline__ void fp_inv_stage(double (&x)[1 << R], const FpC &c, const ntt_cuda_params_t &p, TWF &twf)
{
  constexpr int U = R - 1 - ST, d = 1 << ST;
  if constexpr(LAST) {
    static_assert(U == 0 && COARSE == 0, "the N^-1 stage is the last one and rounds plainly");
    const double2 a = make_double2(p.ninv_fd[0], p.ninv_fd[1]), b = make_double2(p.ninv_w1_fd[0], p.ninv_w1_fd[1]);
#pragma unroll
    for(int k = 0; k < d; k++) {
      const double s = __dadd_rn(x[k], x[k + d]), df = __dadd_rn(x[k], -x[k + d]);
      x[k]           = fp_mul<false>(s, a.x, a.y, c);
      x[k + d]       = fp_mul<false>(df, b.x, b.y, c);
    }
  } else {
#pragma unroll
    for(int sub = 0; sub < (1 << U); sub++) {
      const double2 tw = twf((1 << U) - 1 + sub);
#pragma unroll
      for(int k = 0; k < d; k++) {
        const int    lo = sub * 2 * d + k;
        const double df = __dadd_rn(x[lo], -x[lo + d]);
        x[lo]           = __dadd_rn(x[lo], x[lo + d]);
        x[lo + d]       = ((COARSE >> lo) & 1u) ? fp_mul<true>(df, tw.x, tw.y, c) : fp_mul<false>(df, tw.x, tw.y, c);
      }
    }
  }
}

template <int R, bool FINAL, typename SEL, typename TWF>
__device__ __forceinline__ void fp_network_inv(double (&x)[1 << R], const FpC &c, const ntt_cuda_params_t &p, TWF twf)
{
  constexpr FpPass S = SEL::get();
  fp_fold_mask<R, S.fold_before[0]>(x, c);
  fp_inv_stage<R, 0, S.coarse[0], FINAL && R == 1>(x, c, p, twf);
  if constexpr(R > 1) {
    fp_fold_mask<R, S.fold_before[1]>(x, c);
    fp_inv_stage<R, 1, S.coarse[1], FINAL && R == 2>(x, c, p, twf);
  }
  if constexpr(R > 2) {
    fp_fold_mask<R, S.fold_before[2]>(x, c);
    fp_inv_stage<R, 2, S.coarse[2], FINAL && R == 3>(x, c, p, twf);
  }
  if constexpr(R > 3) {
    fp_fold_mask<R, S.fold_before[3]>(x, c);
    fp_inv_stage<R, 3, S.coarse[3], FINAL && R == 4>(x, c, p, twf);
  }
  if constexpr(R > 4) {
    fp_fold_mask<R, S.fold_before[4]>(x, c);
    fp_inv_stage<R, 4, S.coarse[4], FINAL && R == 5>(x, c, p, twf);
  }
  fp_fold_mask<R, S.fold_end>(x, c);
}

/* -DNTT_RING_TRACE: CTA 0 records clock64() at the phase boundaries of its first 64 polynomials (read back with
 * ntt_cuda_trace_read); how the per-pass cycle counts in DESIGN.md were obtained. */
#ifdef NTT_RING_TRACE
__device__ long long g_trace[16 * 64 * 8]; /* [warp][poly][event] for CTA 0 */
#define TRACE(ev) do { if(blockIdx.x == 0 && lane == 0 && k < 64) g_trace[(warp * 64 + k) * 8 + (ev)] = clock64(); } while(0)
#else
#define TRACE(ev)
#endif

/* MODE (forward only):
 *   RING_MUL   multiply the transform pointwise by `p_other` (a transform of the same shape, canonical residues;
 *              chunk c of this array meets chunk (c & other_mask) of p_other, so one polynomial can be broadcast
 *              over the batch) before it is written -- the NTT-domain product of a negacyclic polynomial multiply,
 *              fused into the second forward transform;
 *   RING_LAZY  leave the output in [0,2q) like fwd_ntt_ref_harvey_lazy leaves it in [0,4q) (src/ntt_reference.c:11-31):
 *              the last fold is kept (the values must fit the window) but the sign correction is not.
 * p_out: the array itself (the inverse writes its results directly). */
enum { RING_PLAIN = 0, RING_MUL = 1, RING_LAZY = 2 };
/*
 * MULTI: one launch over the chunks of SEVERAL plans (RNS limbs: same N, one modulus and one set of tables each;
 * limb l owns polys_per_limb consecutive polynomials of the array).  The per-limb parameters travel as a kernel
 * argument (constant bank).  Every CTA serves ONE limb for its whole life: the grid is (ctas_per_limb, limbs), CTA
 * (x, y) belongs to limb y and takes every ctas_per_limb-th chunk of that limb in the order
 * (chunk-in-polynomial, polynomial), so its twiddle cache changes 2^(logn-14) times and its constants are loop
 * invariants (see where q and 1/q are read in the kernel).  One launch gives every CTA dozens of chunks to pipeline
 * where a launch per limb gives it three or four.
 */
template <bool MULTI>
__device__ __forceinline__ const ntt_cuda_params_t &ring_plan_of(const ntt_cuda_params_t &p0, const RingLimbs<MULTI> &limbs,
                                                                  uint32_t limb)
{
  if constexpr(MULTI) return limbs.e[limb];
  else return p0;
}
template <int L, bool FWD, int MODE = RING_PLAIN, bool Q50 = false, bool MULTI = false>
__global__ void __launch_bounds__(RingCfg<L>::THREADS, RingCfg<L>::CTAS)
  k_ring_fp(const __grid_constant__ ntt_cuda_params_t p0, const __grid_constant__ CUtensorMap tmap,
            const __grid_constant__ CUtensorMap tmap2, size_t n_chunks, uint64_t *__restrict__ p_out,
            const uint64_t *__restrict__ p_other, size_t other_mask, const __grid_constant__ RingLimbs<MULTI> limbs)
{
  constexpr bool MUL = MODE == RING_MUL, LAZY = MODE == RING_LAZY;
  static_assert(!MULTI || MODE == RING_PLAIN, "the multi-plan launch serves plain transforms");
  static_assert(FWD || MODE == RING_PLAIN, "the fused product and the lazy output belong to the forward kernel");
  using C = RingCfg<L>;
  constexpr int NB = C::NB, RA = C::RA, SLOTS = C::SLOTS, T = C::THREADS, HALF = NB / 2;
#ifndef NTT_PIPE_C1_MINL
#define NTT_PIPE_C1_MINL 14
#endif
  constexpr bool PIPE_C1 = !FWD && L >= NTT_PIPE_C1_MINL; /* inverse: see the barrier before the column pass */
  extern __shared__ uint8_t smem_raw[];
  const uint32_t ring     = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *      ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  double2 *      tw_s     = reinterpret_cast<double2 *>(ring_ptr + SLOTS * 4096); /* NTW x 16 bytes */
  const uint32_t bars     = ring + SLOTS * 4096 + C::TW_BYTES_FP;
#ifdef NTT_NO_TWC0
  constexpr bool TWC0 = false;
#else
  constexpr bool TWC0 = FWD && C::NTW_C0 > 0; /* pass C's first-stage twiddles come from shared memory */
#endif
  const uint32_t cta_bar  = bars + 8u * (2u * C::NBAR);

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t s1        = p0.logn - L;
  /* which chunks this CTA works on: every gridDim-th chunk (single plan), or every ctas_per_limb-th chunk of its limb
   * in the order (chunk-in-polynomial, polynomial) (several plans) */
  uint32_t my_limb = 0, cta_in_limb = 0, cpl = 1, ppl = 1;
  if constexpr(MULTI) {
    cpl         = limbs.ctas_per_limb;
    ppl         = limbs.polys_per_limb;
    my_limb     = blockIdx.y; /* grid (ctas_per_limb, limbs): no division, the index stays on the uniform datapath */
    cta_in_limb = blockIdx.x;
  }
  const uint32_t limb_chunks = ppl << s1; /* MULTI: chunks of one limb */
  size_t my_polys = MULTI ? (limb_chunks > cta_in_limb ? (limb_chunks - cta_in_limb + cpl - 1) / cpl : 0)
                          : ((n_chunks > blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  /* single plan, chunks of a LARGER transform (s1 > 0): the work sequence is ordered (chunk-in-polynomial, polynomial)
   * and every CTA takes a CONTIGUOUS range of it, so that it keeps its chunk-in-polynomial -- and with it the 48 KiB
   * twiddle cache -- for as long as possible; in plain chunk order (every gridDim-th chunk) it changed on every chunk
   * from N = 2^17 on (148 CTAs, 2^s1 chunks per polynomial) and the refill cost 15 % of the kernel.  Batches of 2^31
   * chunks or more keep the plain order. */
  const uint32_t n_polys32 = (uint32_t)(n_chunks >> s1);
  const bool     cp_major  = !MULTI && s1 != 0 && n_chunks < ((size_t)1 << 31);
  const size_t   range_lo  = cp_major ? (size_t)blockIdx.x * n_chunks / gridDim.x : 0;
  auto chunk_of = [&](size_t k) -> size_t {
    if(!MULTI) {
      if(!cp_major) return blockIdx.x + k * gridDim.x;
      const uint32_t i = (uint32_t)(range_lo + k), cpv = i / n_polys32, rest = i - cpv * n_polys32;
      return ((size_t)rest << s1) + cpv;
    }
    const uint32_t i = cta_in_limb + (uint32_t)k * cpl, cpv = i / ppl, rest = i - cpv * ppl;
    return ((size_t)(my_limb * ppl + rest) << s1) + cpv;
  };
  if(cp_major) my_polys = (size_t)(blockIdx.x + 1) * n_chunks / gridDim.x - range_lo;
  const size_t   my_blocks = my_polys * NB;
  const size_t   groups    = (size_t)1 << (p0.logn - 4);
  /* The plan of this CTA: the kernel's own (single plan), or its limb's entry of the argument table.  Its constants
   * are loop invariants and are set up here, outside the loop over the chunks. */
  const ntt_cuda_params_t &p = ring_plan_of<MULTI>(p0, limbs, my_limb);
  /* q and 1/q feed one operand of 288 of the forward kernel's 801 DFMAs.  Read as limbs.e[my_limb].q_fd they come out
   * of an indexed constant-bank load into ORDINARY registers and those DFMAs pay for a third register operand (3.2
   * issue cycles instead of 2, tools/ubench_rf.cu); read through a switch whose every case names a fixed table slot,
   * each load has a constant address, lands in a uniform register and stays there, as in the single-plan kernel. */
  double q_sel = p0.q_fd, qinv_sel = p0.qinv_fd;
  if constexpr(MULTI) {
    static_assert(RING_MAX_LIMBS == 48, "the switch below enumerates the table slots");
    switch(my_limb) {
#define NTT_SEL(i) case i: q_sel = limbs.e[i].q_fd; qinv_sel = limbs.e[i].qinv_fd; break;
#define NTT_SEL8(b) NTT_SEL(b) NTT_SEL(b + 1) NTT_SEL(b + 2) NTT_SEL(b + 3) NTT_SEL(b + 4) NTT_SEL(b + 5) NTT_SEL(b + 6) NTT_SEL(b + 7)
      NTT_SEL8(0) NTT_SEL8(8) NTT_SEL8(16) NTT_SEL8(24) NTT_SEL8(32) NTT_SEL8(40)
#undef NTT_SEL8
#undef NTT_SEL
      default: break;
    }
  }
  const FpC      c{q_sel, qinv_sel, NTT_FP_MAGIC};
  /* the conversion centres the input: forward [0,4q) -> [-2q,2q), inverse [0,2q) -> [-q,q) */
  const double   in_bias = -(4503599627370496.0 + (FWD ? 2.0 * p.q_fd : p.q_fd));
  const double   q_bias  = NTT_FP_MAGIC + p.q_fd; /* lazy output: v + q lands in (0, 2q) */
  const double2 *g_fd    = (const double2 *)(FWD ? p.fwd_fd : p.inv_fd);
  const double2 *g_ct    = (const double2 *)(FWD ? p.fwd_ct_fd : p.inv_ct_fd);

  auto slot_addr  = [&](size_t g) -> uint32_t { return ring + (uint32_t)(g % SLOTS) * 4096u; };
  auto issue_load = [&](size_t g) {
    if(g >= my_blocks) return;
    const size_t   k     = g / NB;
    const uint32_t b     = (uint32_t)(g % NB);
    const size_t   chunk = chunk_of(k);
    /* two barriers per polynomial in flight: blocks [0,HALF) and [HALF,NB).  The low half always sits in slots
     * that were free long ago; the high half may have been re-armed only when the previous polynomial was
     * stored, so the inverse starts on the low half while the rest is still landing. */
    const uint32_t bar = bars + 8u * (2u * (uint32_t)(k % C::NBAR) + (b >= (uint32_t)HALF ? 1u : 0u));
    mbar_arrive_expect_tx(bar, 4096u);
    tma_load_block(slot_addr(g), &tmap, (int)((chunk << (L - 4)) + b * 32u), bar);
  };

  /* inverse: all slots of a polynomial die together, so they are re-armed with four boxes of BOXB adjacent
   * blocks (tmap2: BOXB*32 rows, up to 32 KiB) -- four TMA instructions instead of NB on the critical path of
   * pass A; measured, a TMA load costs its issuing thread about 130 cycles whatever the box size. */
  auto issue_box = [&](size_t g) { /* g multiple of BOXB: blocks g .. g+BOXB-1, adjacent slots, one half-barrier */
    if(g >= my_blocks) return;
    const size_t   k     = g / NB;
    const uint32_t b     = (uint32_t)(g % NB);
    const size_t   chunk = chunk_of(k);
    const uint32_t bar = bars + 8u * (2u * (uint32_t)(k % C::NBAR) + (b >= (uint32_t)HALF ? 1u : 0u));
    mbar_arrive_expect_tx(bar, 4096u * C::BOXB);
    tma_load_block(slot_addr(g), &tmap2, (int)((chunk << (L - 4)) + b * 32u), bar);
  };

  if(tid == 0) {
    tma_prefetch_desc(&tmap);
    tma_prefetch_desc(&tmap2);
    /* forward: one arrival per block; inverse: one per box, two boxes per half */
    for(int i = 0; i < 2 * C::NBAR; i++) mbar_init(bars + 8u * i, FWD ? HALF : C::NBOX / 2);
    mbar_init(cta_bar, C::WARPS); /* the inverse's block-wide barrier: one arrival per warp (see below) */
    fence_barrier_init();
  }
  __syncthreads();
  if(FWD) {
    for(uint32_t g = tid; g < (uint32_t)SLOTS; g += T) issue_load(g);
  } else {
    for(uint32_t g = C::BOXB * tid; g < (uint32_t)SLOTS; g += C::BOXB * T) issue_box(g);
  }
  uint32_t cached_cp = 0xffffffffu;
  bool     c1_done   = false; /* inverse: this warp already ran pass C on the first block of polynomial k */
  uint32_t sl_next = 0; /* slot of block 0 of the next polynomial: (k * NB) mod SLOTS, kept in 32 bits */
  for(size_t k = 0; k < my_polys; k++) {
    const size_t   chunk = chunk_of(k);
    const uint32_t cp    = (uint32_t)(chunk & (((size_t)1 << s1) - 1));
    const uint32_t cache_key = cp;
    if(cache_key != cached_cp) {
      __syncthreads();
      for(uint32_t e = tid; e < (uint32_t)C::NTW; e += T) {
        uint32_t t, st, blk;
        if(e < (uint32_t)(NB - 1)) {
          t = e; st = s1; blk = cp;
        } else {
          const uint32_t r = e - (NB - 1);
          t = r % 31u; st = s1 + RA; blk = cp * NB + r / 31u;
        }
        const uint32_t u = 31u - __clz(t + 1u), sub = t + 1u - (1u << u);
        tw_s[e] = __ldg(g_fd + (((size_t)1 << (st + u)) + ((size_t)blk << u) + sub));
      }
      if constexpr(TWC0) {
        for(uint32_t e = tid; e < (uint32_t)C::NTW_C0; e += T) tw_s[C::NTW + e] = __ldg(g_ct + (size_t)cp * NB * 32 + e);
      }
      cached_cp = cache_key;
      __syncthreads();
    }
    const size_t   g0  = k * NB;
    /* SLOTS = 3 * HALF: a polynomial's low and high halves each sit in HALF consecutive slots, so a block address
     * is one of two uniform bases plus a compile-time offset */
    const uint32_t sl0 = sl_next, sh0 = sl0 + HALF >= (uint32_t)SLOTS ? sl0 + HALF - SLOTS : sl0 + HALF;
    sl_next            = sl0 + NB >= (uint32_t)SLOTS ? sl0 + NB - SLOTS : sl0 + NB;
    /* blocks g0+SLOTS .. g0+SLOTS+NB-1 get their slot only when this polynomial's blocks are stored; ask L2
     * for them now so that the late TMA loads find them on chip */
    if(tid < (uint32_t)NB) {
      const size_t g = g0 + SLOTS + tid;
      if(g < my_blocks) {
        const size_t ck = chunk_of(g / NB);
        tma_prefetch_block_l2(&tmap, (int)((ck << (L - 4)) + (uint32_t)(g % NB) * 32u));
      }
    }
    if(FWD && k > 0 && lane == 0) { /* deferred re-arm of the previous polynomial's second block (see below) */
      tma_wait_read_all();
      issue_load(g0 - NB + warp + HALF + SLOTS);
    }
    TRACE(0);
    const uint32_t bar_lo = bars + 16u * (uint32_t)(k % C::NBAR), bar_hi = bar_lo + 8u;
    const uint32_t parity = (uint32_t)((k / C::NBAR) & 1);
    if(FWD) {
      mbar_wait(bar_lo, parity);
      mbar_wait(bar_hi, parity);
    }
    TRACE(1);

    auto blk_slot = [&](uint32_t b) -> uint32_t { return b < (uint32_t)HALF ? sl0 + b : sh0 + (b - HALF); };

    /* pass A, forward: first pass; reads the raw u64 input from the slots, leaves doubles in place. */
    auto pass_a_fwd = [&]() {
      for(uint32_t j = tid; j < 512u; j += T) {
        const uint32_t off = slot_off(j);
        double         x[NB];
#pragma unroll
        for(int b = 0; b < NB; b++)
          x[b] = fp_from_u64(*reinterpret_cast<const uint64_t *>(ring_ptr + blk_slot(b) * 4096u + off), in_bias);
        fp_network_fwd<RA, FpSel<0, Q50, L, 0>>(x, c, [&](int t) { return tw_s[t]; });
#pragma unroll
        for(int b = 0; b < NB; b++)
          *reinterpret_cast<double *>(ring_ptr + blk_slot(b) * 4096u + off) = x[b];
      }
    };
    /* pass A, inverse: last pass.  Every thread first pulls its column(s) into registers; once all have, the
     * polynomial's slots are dead and are re-armed at once (the next polynomials' blocks get
     * a whole pass of lead time), and the results go from registers straight to global memory: column j of
     * block b is word b*512+j, so a warp writes 256 contiguous bytes per store instruction.  No TMA store,
     * no proxy fence, no drain wait on this side. */
    auto pass_a_inv = [&]() {
      constexpr int COLS = 512 / T; /* columns per thread: 1 at L = 14 */
      double        x[COLS][NB];
#pragma unroll
      for(int cidx = 0; cidx < COLS; cidx++) {
        const uint32_t off = slot_off(tid + cidx * T);
#pragma unroll
        for(int b = 0; b < NB; b++) x[cidx][b] = *reinterpret_cast<const double *>(ring_ptr + blk_slot(b) * 4096u + off);
      }
      /* the slots may be overwritten once EVERY warp has pulled its columns: warp 0 waits for that (named
       * barrier, the other warps only arrive and go straight on to their butterflies) and re-arms the ring */
      if(warp == 0) {
        named_sync(T);
        if(lane < (uint32_t)C::NBOX) issue_box(g0 + C::BOXB * lane + SLOTS);
        __syncwarp();
      } else {
        named_arrive(T);
      }
      TRACE(6);
      uint64_t *gout = p_out + (chunk << L);
      if(s1 == 0) {
        /* the chunk is the whole polynomial: global stage 0 with N^-1 is part of this pass and its products
         * (|v| < q) are final */
#pragma unroll
        for(int cidx = 0; cidx < COLS; cidx++) {
          fp_network_inv<RA, true, FpSel<1, Q50, L, 0>>(x[cidx], c, p, [&](int t) { return tw_s[t]; });
          const uint32_t j = tid + cidx * T;
#pragma unroll
          for(int b = 0; b < NB; b++) gout[(size_t)b * 512 + j] = fp_to_u64(x[cidx][b], c, p.q);
        }
      } else {
        /* strided inverse passes follow (they accept [0,2q)): fold and hand over the canonical residue */
#pragma unroll
        for(int cidx = 0; cidx < COLS; cidx++) {
          fp_network_inv<RA, false, FpSel<2, Q50, L, 0>>(x[cidx], c, p, [&](int t) { return tw_s[t]; });
          const uint32_t j = tid + cidx * T;
#pragma unroll
          for(int b = 0; b < NB; b++) gout[(size_t)b * 512 + j] = fp_to_u64(fp_fold(x[cidx][b], c), c, p.q);
        }
      }
    };

    const uint32_t hb = lane >> 4, jb = lane & 15u;
    const uint32_t blkB = warp + hb * HALF;
    auto pass_b = [&]() {
      uint8_t *      base = ring_ptr + blk_slot(blkB) * 4096u + ((jb & 1u) << 3);
      const uint32_t jc   = jb >> 1;
      const double2 *tw   = tw_s + (NB - 1) + blkB * 31;
      double         x[32];
#pragma unroll
      for(int kk = 0; kk < 32; kk++)
        x[kk] = *reinterpret_cast<const double *>(base + kk * 128 + ((jc ^ (uint32_t)(kk & 7)) << 4));
      if constexpr(FWD) fp_network_fwd<5, FpSel<0, Q50, L, 1>>(x, c, [&](int t) { return tw[t]; });
      else fp_network_inv<5, false, FpSel<1, Q50, L, 1>>(x, c, p, [&](int t) { return tw[t]; });
#pragma unroll
      for(int kk = 0; kk < 32; kk++)
        *reinterpret_cast<double *>(base + kk * 128 + ((jc ^ (uint32_t)(kk & 7)) << 4)) = x[kk];
    };

    /* pass C: 16 contiguous coefficients.  Forward: last pass, writes canonical u64.  Inverse: first pass,
     * reads the raw u64 input (contract [0,2q)). */
    auto pass_c_at = [&](uint32_t slot, uint32_t blk, uint32_t cp, size_t chunk, bool rearm_first) {
      uint8_t *base = ring_ptr + slot * 4096u + lane * 128u;
      double   x[16];
#pragma unroll
      for(int cc = 0; cc < 8; cc++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(base + (((uint32_t)cc ^ (lane & 7u)) << 4));
        x[2 * cc]     = FWD ? __longlong_as_double((long long)v.x) : fp_from_u64(v.x, in_bias);
        x[2 * cc + 1] = FWD ? __longlong_as_double((long long)v.y) : fp_from_u64(v.y, in_bias);
      }
      const double2 *tw = g_ct + ((size_t)cp * NB + blk) * 32 + lane;
      const double2 *tw0 = tw_s + C::NTW + blk * 32 + lane;
      auto           twf = [&](int t) {
        if(TWC0 && t == 0) return *tw0;
        if constexpr(MULTI) return ldg_keep(tw + (size_t)t * groups); /* many plans: pin the tables in L2 (ntt_ring.cuh) */
        else return __ldg(tw + (size_t)t * groups);
      };
      if constexpr(FWD) {
        fp_network_fwd<4, FpSel<0, Q50, L, 2>>(x, c, twf, [&]() {
          /* the first block's store has had this block's loads and two stages of butterflies to drain: its slot is
           * re-armed here without stalling the warp (the block it receives is needed by the next polynomial's pass A) */
          if(rearm_first && lane == 0) {
            tma_wait_read_all();
            issue_load(g0 + warp + SLOTS);
          }
        });
      } else {
        fp_network_inv<4, false, FpSel<1, Q50, L, 2>>(x, c, p, twf);
      }
#pragma unroll
      for(int cc = 0; cc < 8; cc++) {
        ulonglong2 v;
        if(FWD && MUL) {
          /* other operand: same position of the other transform; o/q is rounded on the fly (it only steers the
           * quotient estimate).  The product of a folded value with a residue below q is below 0.51q. */
          const ulonglong2 o = __ldg(reinterpret_cast<const ulonglong2 *>(
            p_other + ((chunk & other_mask) << L) + (size_t)blk * 512 + lane * 16 + 2 * cc));
          const double o0 = fp_from_u64(o.x), o1 = fp_from_u64(o.y);
          v.x = fp_to_u64(fp_mul(fp_fold(x[2 * cc], c), o0, __dmul_rn(o0, c.qinv), c), c, p.q);
          v.y = fp_to_u64(fp_mul(fp_fold(x[2 * cc + 1], c), o1, __dmul_rn(o1, c.qinv), c), c, p.q);
        } else if(FWD && LAZY) {
          v.x = fp_to_u64_lazy(fp_fold(x[2 * cc], c), q_bias);
          v.y = fp_to_u64_lazy(fp_fold(x[2 * cc + 1], c), q_bias);
        } else if(FWD) {
          v.x = fp_to_u64(fp_fold(x[2 * cc], c), c, p.q);
          v.y = fp_to_u64(fp_fold(x[2 * cc + 1], c), c, p.q);
        } else {
          v.x = (uint64_t)__double_as_longlong(x[2 * cc]);
          v.y = (uint64_t)__double_as_longlong(x[2 * cc + 1]);
        }
        *reinterpret_cast<ulonglong2 *>(base + (((uint32_t)cc ^ (lane & 7u)) << 4)) = v;
      }
    };

    auto pass_c = [&](uint32_t blk, bool rearm_first) { pass_c_at(blk_slot(blk), blk, cp, chunk, rearm_first); };

    auto store_block = [&](uint32_t b) {
      tma_store_block(&tmap, (int)((chunk << (L - 4)) + b * 32u), ring + blk_slot(b) * 4096u);
      tma_commit();
    };

    if(FWD) {
      pass_a_fwd();
      TRACE(2);
      __syncthreads();
      TRACE(3);
      pass_b();
      __syncwarp();
      TRACE(4);
      pass_c(warp, false);
      fence_proxy_async();
      __syncwarp();
      if(lane == 0) store_block(warp);
      TRACE(5);
      pass_c(warp + HALF, true);
      fence_proxy_async();
      __syncwarp();
      TRACE(6);
      /* the second block's slot is re-armed at the top of the next iteration (its new block is only needed a whole
       * polynomial later), when the store has long drained */
      if(lane == 0) store_block(warp + HALF);
      __syncwarp();
      TRACE(7);
    } else {
      if(!c1_done) {
        mbar_wait(bar_lo, parity);
        pass_c(warp, false);
      }
      c1_done = false;
      TRACE(2);
      mbar_wait(bar_hi, parity);
      pass_c(warp + HALF, false);
      __syncwarp();
      TRACE(3);
      pass_b();
      TRACE(4);
      /* Split block-wide barrier with independent work in between: every warp ARRIVES once its blocks are through
       * pass B, then runs pass C on the first block of its NEXT polynomial (resident since the previous re-arm,
       * warp-private), and only then WAITS -- by which time the other warps have usually arrived, so the barrier
       * costs no idle time (measured at N = 2^14: inverse 0.407 -> 0.373 ms per 4096; the same reordering around a
       * plain __syncthreads gains nothing).  mbarrier because bar.arrive + bar.sync would count the warp twice.
       * PIPE_C1 is off where it measured slower (the 8- and 4-warp CTAs of L = 13 / 12). */
      if(PIPE_C1) {
        __syncwarp(); /* orders the other lanes' pass-B stores before lane 0's releasing arrive */
        if(lane == 0) mbar_arrive(cta_bar);
        bool prerun = k + 1 < my_polys;
        if(prerun) {
          mbar_wait(bars + 16u * (uint32_t)((k + 1) % C::NBAR), (uint32_t)(((k + 1) / C::NBAR) & 1));
          const size_t nchunk = chunk_of(k + 1);
          pass_c_at(sl_next + warp, warp, (uint32_t)(nchunk & (((size_t)1 << s1) - 1)), nchunk, false);
          c1_done = true;
        }
        mbar_wait(cta_bar, (uint32_t)(k & 1)); /* acquire: every lane waits itself */
      } else {
        __syncthreads();
      }
      TRACE(5);
      pass_a_inv();
      TRACE(7);
    }
  }
  tma_wait_all();
}

}  // namespace nttb200
