/*
 * csrc/ntt_polymul_fp.cuh -- negacyclic polynomial multiply in ONE kernel (BASELINE config 4):
 * c = INTT( NTT(a) .* NTT(b) ) for N = 2^13, q <= 2^50 - 2048, both operands resident in shared memory.
 *
 * What it composes (reference semantics): fwd_ntt_ref_harvey on a and on b (include/ntt_reference.h:19-31), a
 * pointwise product mod q, inv_ntt_ref_harvey (src/ntt_reference.c:33-66).  The NTT-domain arrays never exist in
 * global memory: per product the kernel reads a and b once (2*N*8 bytes) and writes c once (N*8).
 *
 * Geometry: the ring of ntt_ring.cuh at its L = 14 size -- 512 threads, 48 slots of 4 KiB -- where a "chunk" is the
 * PAIR (a, b): a's 16 blocks of 512 coefficients sit in the low half of the chunk's slots, b's 16 blocks in the high
 * half.  Both arrive by TMA boxes (8 blocks per box, one mbarrier per operand and pair in flight); the next pair's a
 * is resident while this one is worked on, the rest follows when this pair's slots die.
 *
 *   pass A  forward stages 0-3 across blocks, thread j = column j: once for a, once for b (two 16-value networks)
 *           __syncthreads
 *   pass B  forward stages 4-8 inside a block: the two half-warps of warp w own a's block w and b's block w (same
 *           twiddles), 32 values at stride 16 per thread
 *           __syncwarp
 *   pass C  forward stages 9-12 on 16 contiguous coefficients, a's and b's block w side by side in registers with
 *           ONE set of per-thread twiddles from L2; fold both; product (the second operand's quotient factor is
 *           rounded on the fly); inverse stages 12-9 on the product, still in registers -- the thread that holds 16
 *           contiguous coefficients of NTT(a) and NTT(b) holds the same 16 of the product
 *           product -> a's slot, __syncwarp
 *   pass B' inverse stages 8-4 on the product block.  There is ONE product block per warp where the forward passes
 *           had two, so the 32-value network of a column group is split over both half-warps: lane (jb, h) loads
 *           all 32 values, stage 8 pairs positions (2i, 2i+1) -- half h = 0 keeps the sums, half h = 1 the products --
 *           and stages 7-4 run on each lane's own 16 values (positions 2i + h) with all 32 lanes busy
 *           __syncthreads
 *   pass A' inverse stages 3-0 across the 16 product blocks with the N^-1 stage, thread j = column j; the columns are
 *           pulled into registers, the pair's 32 slots are re-armed with four TMA boxes, results go from registers
 *           straight to c (a warp writes 256 contiguous bytes per store).
 *
 * Range schedules: the forward passes of the L = 13 transform, a fold, the product (|p| <= 0.5625 q, inside the
 * inverse's input bound of q), then the L = 13 inverse passes of ntt_fp_schedule.h.
 * c may alias a or b (a pair is completely on chip before its product is written), and a == b squares.
 */
#pragma once
#include "ntt_ring_fp.cuh"

namespace nttb200 {

struct PolymulCfg {
  static constexpr int L       = 13;               /* transform size */
  static constexpr int PB      = 16;               /* blocks per operand */
  static constexpr int NB      = 32;               /* blocks per pair */
  static constexpr int T       = 512;
  static constexpr int SLOTS   = 48;
  static constexpr int NBAR    = 4;
  static constexpr int BOXB    = 8;                /* blocks per TMA box: 256 rows of 128 bytes */
  static constexpr int NTW     = PB - 1 + PB * 31; /* per direction: pass A + pass B twiddles */
  static constexpr int TW_BYTES = ((2 * NTW * 16 + 127) / 128) * 128;
  static constexpr int SMEM    = SLOTS * 4096 + 1024 + TW_BYTES + 128;
  static_assert(SMEM <= 232448, "shared memory budget");
};

/* forward stage U of two R-stage networks that share their twiddles */
template <int R, int U, uint32_t COARSE, typename TWF>
__device__ __forceinline__ void fp_fwd_stage2(double (&xa)[1 << R], double (&xb)[1 << R], const FpC &c, TWF &twf)
{
  constexpr int n = 1 << R, d = n >> (U + 1);
#pragma unroll
  for(int sub = 0; sub < (1 << U); sub++) {
    const double2 tw = twf((1 << U) - 1 + sub);
#pragma unroll
    for(int k = 0; k < d; k++) {
      const int lo = sub * 2 * d + k;
      {
        const double t = ((COARSE >> lo) & 1u) ? fp_mul<true>(xa[lo + d], tw.x, tw.y, c) : fp_mul<false>(xa[lo + d], tw.x, tw.y, c);
        xa[lo + d]     = __dadd_rn(xa[lo], -t);
        xa[lo]         = __dadd_rn(xa[lo], t);
      }
      {
        const double t = ((COARSE >> lo) & 1u) ? fp_mul<true>(xb[lo + d], tw.x, tw.y, c) : fp_mul<false>(xb[lo + d], tw.x, tw.y, c);
        xb[lo + d]     = __dadd_rn(xb[lo], -t);
        xb[lo]         = __dadd_rn(xb[lo], t);
      }
    }
  }
}

template <typename SEL, typename TWF>
__device__ __forceinline__ void fp_network_fwd2_r4(double (&xa)[16], double (&xb)[16], const FpC &c, TWF twf)
{
  constexpr FpPass S = SEL::get();
  fp_fold_mask<4, S.fold_before[0]>(xa, c);
  fp_fold_mask<4, S.fold_before[0]>(xb, c);
  fp_fwd_stage2<4, 0, S.coarse[0]>(xa, xb, c, twf);
  fp_fold_mask<4, S.fold_before[1]>(xa, c);
  fp_fold_mask<4, S.fold_before[1]>(xb, c);
  fp_fwd_stage2<4, 1, S.coarse[1]>(xa, xb, c, twf);
  fp_fold_mask<4, S.fold_before[2]>(xa, c);
  fp_fold_mask<4, S.fold_before[2]>(xb, c);
  fp_fwd_stage2<4, 2, S.coarse[2]>(xa, xb, c, twf);
  fp_fold_mask<4, S.fold_before[3]>(xa, c);
  fp_fold_mask<4, S.fold_before[3]>(xb, c);
  fp_fwd_stage2<4, 3, S.coarse[3]>(xa, xb, c, twf);
  fp_fold_mask<4, S.fold_end>(xa, c);
  fp_fold_mask<4, S.fold_end>(xb, c);
}

/* inverse schedules of this kernel (ntt_fp_schedule.h, FP_SCHED_INV_POLYMUL): WHICH 0 / 1 / 2 = pass A / B / C */
template <bool Q50, int WHICH>
struct FpSelPm {
  static __host__ __device__ constexpr FpPass get()
  {
    const FpSchedule &s = FP_SCHED_INV_POLYMUL[Q50];
    return WHICH == 0 ? s.a : (WHICH == 1 ? s.b : s.c);
  }
};
/* bit i of the result = bit 2i of a 32-position mask (the masks of the split pass B are symmetric in 2i / 2i+1) */
__host__ __device__ constexpr uint32_t pm_half_mask(uint32_t m)
{
  uint32_t r = 0;
  for(int i = 0; i < 16; i++) r |= ((m >> (2 * i)) & 1u) << i;
  return r;
}

template <bool Q50>
__global__ void __launch_bounds__(PolymulCfg::T, 1)
  k_polymul_fp(const __grid_constant__ ntt_cuda_params_t p, const __grid_constant__ CUtensorMap tmap_a,
               const __grid_constant__ CUtensorMap tmap_b, size_t n_pairs, uint64_t *__restrict__ p_out)
{
  using C = PolymulCfg;
  constexpr int L = C::L, PB = C::PB, NB = C::NB, SLOTS = C::SLOTS, T = C::T, HALF = PB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t ring     = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *      ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  double2 *      tw_f     = reinterpret_cast<double2 *>(ring_ptr + SLOTS * 4096); /* forward: pass A, pass B */
  double2 *      tw_i     = tw_f + C::NTW;                                          /* inverse */
  const uint32_t bars     = ring + SLOTS * 4096 + C::TW_BYTES;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t   my_pairs  = (n_pairs > blockIdx.x) ? (n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const size_t   my_blocks = my_pairs * NB;
  const size_t   groups    = (size_t)1 << (L - 4);
  const FpC      c{p.q_fd, p.qinv_fd, NTT_FP_MAGIC};
  const double   in_bias = -(4503599627370496.0 + 2.0 * p.q_fd); /* forward input [0,4q) -> [-2q,2q) */
  const double2 *g_ctf = (const double2 *)p.fwd_ct_fd, *g_cti = (const double2 *)p.inv_ct_fd;

  auto slot_addr = [&](size_t g) -> uint32_t { return ring + (uint32_t)(g % SLOTS) * 4096u; };
  /* g multiple of BOXB: blocks g .. g+BOXB-1 of the CTA's sequence (pair g/NB; blocks below PB are a's) */
  auto issue_box = [&](size_t g) {
    if(g >= my_blocks) return;
    const size_t   k    = g / NB;
    const uint32_t b    = (uint32_t)(g % NB);
    const size_t   pair = blockIdx.x + k * gridDim.x;
    const bool     hi   = b >= (uint32_t)HALF;
    const uint32_t bar  = bars + 8u * (2u * (uint32_t)(k % C::NBAR) + (hi ? 1u : 0u));
    mbar_arrive_expect_tx(bar, 4096u * C::BOXB);
    tma_load_block(slot_addr(g), hi ? &tmap_b : &tmap_a, (int)((pair << (L - 4)) + (b - (hi ? HALF : 0)) * 32u), bar);
  };

  if(tid == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for(int i = 0; i < 2 * C::NBAR; i++) mbar_init(bars + 8u * i, 2); /* two boxes per operand */
    fence_barrier_init();
  }
  /* twiddles of passes A and B, both directions (the chunk is the whole polynomial: they never change) */
  for(uint32_t e = tid; e < 2u * C::NTW; e += T) {
    const bool     inv = e >= (uint32_t)C::NTW;
    const uint32_t ee  = inv ? e - C::NTW : e;
    uint32_t       t, st, blk;
    if(ee < (uint32_t)(PB - 1)) {
      t = ee; st = 0; blk = 0;
    } else {
      const uint32_t r = ee - (PB - 1);
      t = r % 31u; st = 4; blk = r / 31u;
    }
    const uint32_t u = 31u - __clz(t + 1u), sub = t + 1u - (1u << u);
    const double2 *g = (const double2 *)(inv ? p.inv_fd : p.fwd_fd);
    tw_f[e]          = __ldg(g + (((size_t)1 << (st + u)) + ((size_t)blk << u) + sub));
  }
  __syncthreads();
  for(uint32_t g = C::BOXB * tid; g < (uint32_t)SLOTS; g += C::BOXB * T) issue_box(g);

  uint32_t sl_next = 0;
  for(size_t k = 0; k < my_pairs; k++) {
    const size_t   pair = blockIdx.x + k * gridDim.x;
    const size_t   g0   = k * NB;
    const uint32_t sl0 = sl_next, sh0 = sl0 + HALF >= (uint32_t)SLOTS ? sl0 + HALF - SLOTS : sl0 + HALF;
    sl_next            = sl0 + NB >= (uint32_t)SLOTS ? sl0 + NB - SLOTS : sl0 + NB;
    const uint32_t bar_lo = bars + 16u * (uint32_t)(k % C::NBAR), bar_hi = bar_lo + 8u;
    const uint32_t parity = (uint32_t)((k / C::NBAR) & 1);

    /* ---- pass A, forward: column tid of a's blocks, then of b's ---------------------------------------- */
    {
      const uint32_t off = slot_off(tid);
      double         x[PB];
      mbar_wait(bar_lo, parity);
#pragma unroll
      for(int b = 0; b < PB; b++)
        x[b] = fp_from_u64(*reinterpret_cast<const uint64_t *>(ring_ptr + (sl0 + b) * 4096u + off), in_bias);
      fp_network_fwd<4, FpSel<0, Q50, L, 0>>(x, c, [&](int t) { return tw_f[t]; });
#pragma unroll
      for(int b = 0; b < PB; b++) *reinterpret_cast<double *>(ring_ptr + (sl0 + b) * 4096u + off) = x[b];
      mbar_wait(bar_hi, parity);
#pragma unroll
      for(int b = 0; b < PB; b++)
        x[b] = fp_from_u64(*reinterpret_cast<const uint64_t *>(ring_ptr + (sh0 + b) * 4096u + off), in_bias);
      fp_network_fwd<4, FpSel<0, Q50, L, 0>>(x, c, [&](int t) { return tw_f[t]; });
#pragma unroll
      for(int b = 0; b < PB; b++) *reinterpret_cast<double *>(ring_ptr + (sh0 + b) * 4096u + off) = x[b];
    }
    __syncthreads();

    const uint32_t hb = lane >> 4, jb = lane & 15u;
    const uint32_t slot_a = sl0 + warp, slot_b = sh0 + warp;
    /* ---- pass B, forward: half-warp 0 on a's block `warp`, half-warp 1 on b's block `warp` ------------------ */
    {
      uint8_t *      base = ring_ptr + (hb ? slot_b : slot_a) * 4096u + ((jb & 1u) << 3);
      const uint32_t jc   = jb >> 1;
      const double2 *tw   = tw_f + (PB - 1) + warp * 31;
      double         x[32];
#pragma unroll
      for(int kk = 0; kk < 32; kk++)
        x[kk] = *reinterpret_cast<const double *>(base + kk * 128 + ((jc ^ (uint32_t)(kk & 7)) << 4));
      fp_network_fwd<5, FpSel<0, Q50, L, 1>>(x, c, [&](int t) { return tw[t]; });
#pragma unroll
      for(int kk = 0; kk < 32; kk++)
        *reinterpret_cast<double *>(base + kk * 128 + ((jc ^ (uint32_t)(kk & 7)) << 4)) = x[kk];
    }
    __syncwarp();

    /* ---- pass C forward on both operands, product, pass C inverse: all in registers ------------------------- */
    {
      uint8_t *base_a = ring_ptr + slot_a * 4096u + lane * 128u;
      uint8_t *base_b = ring_ptr + slot_b * 4096u + lane * 128u;
      double   xa[16], xb[16];
#pragma unroll
      for(int cc = 0; cc < 8; cc++) {
        const uint32_t   o  = (((uint32_t)cc ^ (lane & 7u)) << 4);
        const ulonglong2 va = *reinterpret_cast<const ulonglong2 *>(base_a + o);
        const ulonglong2 vb = *reinterpret_cast<const ulonglong2 *>(base_b + o);
        xa[2 * cc]     = __longlong_as_double((long long)va.x);
        xa[2 * cc + 1] = __longlong_as_double((long long)va.y);
        xb[2 * cc]     = __longlong_as_double((long long)vb.x);
        xb[2 * cc + 1] = __longlong_as_double((long long)vb.y);
      }
      const double2 *twf = g_ctf + (size_t)warp * 32 + lane;
      fp_network_fwd2_r4<FpSel<0, Q50, L, 2>>(xa, xb, c, [&](int t) { return __ldg(twf + (size_t)t * groups); });
#pragma unroll
      for(int i = 0; i < 16; i++) {
        /* fold both; the product of two folded values is below 0.5625 q in magnitude (the quotient factor of
         * the second operand, RN(b * RN(1/q)), is within 2^-53 of b/q because |b| <= q/2 + 6) */
        const double a = fp_fold(xa[i], c), b = fp_fold(xb[i], c);
        xa[i]          = fp_mul<false>(a, b, __dmul_rn(b, c.qinv), c);
      }
      const double2 *twi = g_cti + (size_t)warp * 32 + lane;
      fp_network_inv<4, false, FpSelPm<Q50, 2>>(xa, c, p, [&](int t) { return __ldg(twi + (size_t)t * groups); });
#pragma unroll
      for(int cc = 0; cc < 8; cc++) {
        ulonglong2 v;
        v.x = (uint64_t)__double_as_longlong(xa[2 * cc]);
        v.y = (uint64_t)__double_as_longlong(xa[2 * cc + 1]);
        *reinterpret_cast<ulonglong2 *>(base_a + (((uint32_t)cc ^ (lane & 7u)) << 4)) = v;
      }
    }
    __syncwarp();

    /* ---- pass B, inverse, on the product block, split over both half-warps (see the header) ---------------------- */
    {
      constexpr FpPass SB   = FpSelPm<Q50, 1>::get();
      uint8_t *        base = ring_ptr + slot_a * 4096u + ((jb & 1u) << 3);
      const uint32_t   jc   = jb >> 1;
      const double2 *  tw   = tw_i + (PB - 1) + warp * 31;
      double           xin[32], x[16];
#pragma unroll
      for(int kk = 0; kk < 32; kk++)
        xin[kk] = *reinterpret_cast<const double *>(base + kk * 128 + ((jc ^ (uint32_t)(kk & 7)) << 4));
      fp_fold_mask<5, SB.fold_before[0]>(xin, c);
      if(hb == 0) {
#pragma unroll
        for(int i = 0; i < 16; i++) x[i] = __dadd_rn(xin[2 * i], xin[2 * i + 1]);
      } else {
#pragma unroll
        for(int i = 0; i < 16; i++) {
          const double2 t  = tw[15 + i]; /* network stage u = 4: sub-block i uses twiddle 2^4 - 1 + i */
          const double  df = __dadd_rn(xin[2 * i], -xin[2 * i + 1]);
          x[i] = ((SB.coarse[0] >> (2 * i)) & 1u) ? fp_mul<true>(df, t.x, t.y, c) : fp_mul<false>(df, t.x, t.y, c);
        }
      }
      /* stages 1..4 of the 5-stage network on positions 2i + h = stages 0..3 of a 4-stage network on i, same
       * twiddle indices (the sub-block of position 2i + h at distance 2^s is the sub-block of i at 2^(s-1)) */
      auto twf = [&](int t) { return tw[t]; };
      fp_fold_mask<4, pm_half_mask(SB.fold_before[1])>(x, c);
      fp_inv_stage<4, 0, pm_half_mask(SB.coarse[1]), false>(x, c, p, twf);
      fp_fold_mask<4, pm_half_mask(SB.fold_before[2])>(x, c);
      fp_inv_stage<4, 1, pm_half_mask(SB.coarse[2]), false>(x, c, p, twf);
      fp_fold_mask<4, pm_half_mask(SB.fold_before[3])>(x, c);
      fp_inv_stage<4, 2, pm_half_mask(SB.coarse[3]), false>(x, c, p, twf);
      fp_fold_mask<4, pm_half_mask(SB.fold_before[4])>(x, c);
      fp_inv_stage<4, 3, pm_half_mask(SB.coarse[4]), false>(x, c, p, twf);
      fp_fold_mask<4, pm_half_mask(SB.fold_end)>(x, c);
      __syncwarp(); /* every lane has read all 32 positions before any of them is overwritten */
#pragma unroll
      for(int i = 0; i < 16; i++) {
        const uint32_t kk = 2u * i + hb;
        *reinterpret_cast<double *>(base + kk * 128 + ((jc ^ (kk & 7u)) << 4)) = x[i];
      }
    }
    __syncthreads();

    /* ---- pass A, inverse, with the N^-1 stage; slots re-armed as soon as every thread holds its column -------- */
    {
      const uint32_t off = slot_off(tid);
      double         x[PB];
#pragma unroll
      for(int b = 0; b < PB; b++) x[b] = *reinterpret_cast<const double *>(ring_ptr + (sl0 + b) * 4096u + off);
      if(warp == 0) {
        named_sync(T);
        if(lane < 4) issue_box(g0 + C::BOXB * lane + SLOTS);
        __syncwarp();
      } else {
        named_arrive(T);
      }
      fp_network_inv<4, true, FpSelPm<Q50, 0>>(x, c, p, [&](int t) { return tw_i[t]; });
      uint64_t *gout = p_out + (pair << L);
#pragma unroll
      for(int b = 0; b < PB; b++) gout[(size_t)b * 512 + tid] = fp_to_u64(x[b], c, p.q);
    }
  }
}

}  // namespace nttb200
