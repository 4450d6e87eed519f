/*
 * csrc/ntt_ring.cuh -- the headline kernel: persistent CTAs, TMA slot ring, three register-tiled passes.
 *
 * A chunk of 2^L coefficients (L = 10 .. 14; a whole polynomial when N = 2^L) is NB = 2^(L-9) "blocks"
 * of 512 coefficients (4 KiB).  Shared memory is a ring of SLOTS 4-KiB slots; block g of the CTA's work
 * sequence lives in slot g mod SLOTS.  Blocks arrive by TMA (cp.async.bulk.tensor, SWIZZLE_128B, one
 * mbarrier per polynomial-in-flight), results leave by TMA store straight out of the slot, and the slot is
 * re-armed with the block SLOTS positions further down the sequence.  With SLOTS = 48 at L = 14 (one and a half
 * polynomials) half of the next polynomial is resident before the current one finishes and the rest follows block by block, so HBM traffic overlaps the
 * butterflies without any register staging.
 *
 * Forward schedule for one chunk (stage numbers local to the chunk; the inverse mirrors it):
 *   pass A  stages 0 .. RA-1 (RA = L-9): butterflies ACROSS blocks; thread j owns column j of every block
 *           (element j + 512*k, k < NB): one LDS.64/STS.64 per element, conflict free.
 *           __syncthreads
 *   pass B  stages RA .. RA+4: inside a block, 32 elements at stride 16 per thread; the two half-warps of
 *           warp w own blocks w and w + NB/2.
 *           __syncwarp   (blocks are warp-private from here on)
 *   pass C  stages RA+5 .. L-1: 16 contiguous coefficients (one swizzled 128-byte row) per thread, final
 *           reduction to [0,q), LDS.128/STS.128; then lane 0 issues the TMA store of the finished block.
 *
 * Butterflies use the lazy split multiplier of ntt_device.cuh (8 IMAD + 1 SHF + 4 IADD3 per forward
 * butterfly): no conditional subtraction anywhere before the final reduction.  Twiddles of pass C are per-thread distinct; they come from a table laid out
 * [t][group] (t = 2^u-1+sub) so that a warp reads 32 consecutive entries (ntt_cuda_params_t::*_ct_*).
 *
 * Reference semantics restated: src/ntt_reference.c:11-31 (forward), :33-66 (inverse),
 * include/ntt_reference.h:19-31 (final reduction).
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "ntt_cuda.h"
#include "ntt_device.cuh"

namespace nttb200 {

/* ---- PTX wrappers: mbarrier + TMA -------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
/* Plain try_wait loop.  A suspend-time hint (the form with a third operand) was measured and does not pay: the
 * inverse kernel spends about 340 instructions per thread and polynomial polling (TRYWAIT, BRA, YIELD), but they
 * are issued by warps that have nothing else to do, and with the hint both kernels were 0.3 % slower
 * (profiles/r02_kernel_experiments.txt). */
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
    "{\n\t"
    ".reg .pred p;\n\t"
    "WAIT_%=:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra DONE_%=;\n\t"
    "bra WAIT_%=;\n\t"
    "DONE_%=:\n\t"
    "}" ::"r"(bar),
    "r"(parity)
    : "memory");
}
/* named barrier 1: arrive without waiting / wait for `count` threads (producer-consumer hand-off inside the CTA) */
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_arrive(uint32_t count) { asm volatile("bar.arrive 1, %0;" ::"r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(uint32_t count) { asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory"); }
__device__ __forceinline__ void fence_barrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
/* L2 residency (north star: twiddle tables "kept L2-resident"): the coefficient stream is read once and written once,
 * so its TMA loads and stores carry an evict-first policy, and the per-thread twiddle loads of the last pass an
 * evict-last one -- the tables (1 MiB per plan and direction) then survive the 128 KiB-per-transform stream in the
 * 126 MB L2 whatever the batch size.  Measured (profiles/r02g_*): with ONE plan the tables stay resident anyway and
 * the hinted table load costs the forward kernel 2 % (one more register pair and a volatile load in a schedule that is
 * sensitive to both), so single-plan kernels hint the stream only; with 48 plans in one launch (RNS limbs, 47 MB of
 * last-pass tables in flight) 60 % of the table reads missed L2 and came from DRAM -- there the table loads carry the
 * keep policy too.  -DNTT_L2HINT=0 builds without any hint (A/B timing). */
#ifndef NTT_L2HINT
#define NTT_L2HINT 1
#endif
__device__ __forceinline__ uint64_t l2_policy_stream()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_keep()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
/* 16-byte read-only load of a table entry with the keep policy */
__device__ __forceinline__ double2 ldg_keep(const double2 *ptr)
{
#if NTT_L2HINT
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(l2_policy_keep()));
  return v;
#else
  return __ldg(ptr);
#endif
}
/* global -> shared tile load (box {32 x u32, 32 rows} = 4 KiB), completion counted on `bar` */
__device__ __forceinline__ void tma_load_block(uint32_t dst, const CUtensorMap *tm, int row, uint32_t bar)
{
#if NTT_L2HINT
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::
      "r"(dst),
    "l"(tm), "r"(0), "r"(row), "r"(bar), "l"(l2_policy_stream())
    : "memory");
#else
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
      "r"(dst),
    "l"(tm), "r"(0), "r"(row), "r"(bar)
    : "memory");
#endif
}
/* shared -> global tile store, tracked by the issuing thread's bulk async-group */
__device__ __forceinline__ void tma_store_block(const CUtensorMap *tm, int row, uint32_t src)
{
#if NTT_L2HINT
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(tm),
               "r"(0), "r"(row), "r"(src), "l"(l2_policy_stream())
               : "memory");
#else
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(0),
               "r"(row), "r"(src)
               : "memory");
#endif
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
/* all of this thread's store groups have finished READING shared memory (slots may be overwritten) */
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
/* ... and have finished writing global memory */
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
/* pull one 4-KiB block into L2 ahead of its (later) TMA load */
__device__ __forceinline__ void tma_prefetch_block_l2(const CUtensorMap *tm, int row)
{
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(0), "r"(row) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm)
{
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

/* ---- geometry ---------------------------------------------------------------------------------------------- */

template <int L>
struct RingCfg {
  static_assert(L >= 10 && L <= 14, "ring kernel covers chunks of 2^10 .. 2^14");
  static constexpr int RA       = L - 9;                  /* stages of pass A */
  static constexpr int NB       = 1 << RA;                /* 512-element blocks per chunk */
  static constexpr int WARPS    = NB / 2;
  static constexpr int THREADS  = WARPS * 32;
  /* resident CTAs per SM: 16 warps per SM, except at L = 11 / 10 (two warps / one warp per CTA), where that many CTAs
   * would leave each of them less shared memory than a ring deeper than one polynomial needs: six CTAs with eight slots
   * each at L = 11, twelve CTAs with three slots each at L = 10 */
  static constexpr int CTAS     = L == 11 ? 6 : (L == 10 ? 12 : 512 / THREADS);
  static constexpr int NBAR     = 4;                      /* polynomials in flight (mbarrier ring) */
  /* shared-memory copy of the twiddles of passes A and B: NB-1 entries for pass A, 31 per block for pass B */
  static constexpr int NTW      = NB - 1 + NB * 31;
  static constexpr int TW_BYTES = ((NTW * 24 + 127) / 128) * 128;
  /* 228 KiB per SM, 1 KiB reserved per resident CTA, at most 227 KiB per CTA */
  static constexpr int PER_CTA  = 233472 / CTAS - 1024;
  static constexpr int BUDGET   = PER_CTA < 232448 ? PER_CTA : 232448;
  /* the inverse re-arms a dead polynomial with NBOX = 4 big TMA boxes of BOXB = NB/4 adjacent blocks each (two boxes of
   * one block at L = 10; BOXB*32 rows <= 256): the ring depth is a multiple of BOXB so that a box never wraps */
  static constexpr int BOXB     = NB >= 4 ? NB / 4 : 1;
  static constexpr int NBOX     = NB / BOXB;
  static constexpr int SLOTS    = ((BUDGET - TW_BYTES - 1024 - 128) / 4096) / (NB / 2) * (NB / 2); /* 48 / 24 / 12 / 8 / 3 for L = 14 / 13 / 12 / 11 / 10 */
  static constexpr int SMEM     = SLOTS * 4096 + 1024 /* alignment slack */ + TW_BYTES + 128 /* 8 load barriers + the CTA barrier */;
  /* FP64 kernel: 16-byte twiddle entries; at L = 14 the cache also holds the FIRST-stage twiddle of pass C for every
   * 16-coefficient group (NB*32 entries), so that the last pass starts from shared memory while its other 14
   * per-thread twiddles are still on their way from L2 */
  static constexpr int NTW_C0      = (L == 14) ? NB * 32 : 0;
  static constexpr int TW_BYTES_FP = (((NTW + NTW_C0) * 16 + 127) / 128) * 128 > TW_BYTES ? (((NTW + NTW_C0) * 16 + 127) / 128) * 128 : TW_BYTES;
  static constexpr int SMEM_FP     = SLOTS * 4096 + 1024 + TW_BYTES_FP + 128;
  static_assert(SMEM_FP <= BUDGET, "FP64 ring kernel: shared memory budget");
  static_assert(SLOTS > NB && 4 * NB > SLOTS && SLOTS % BOXB == 0, "ring depth vs. mbarrier reuse distance");
  static_assert(SLOTS % (NB / 2) == 0, "half a polynomial must not wrap around the ring (blk_slot)");
};

/* byte offset of coefficient o (0..511) inside a 4-KiB slot under TMA SWIZZLE_128B (rows of 128 bytes) */
__device__ __forceinline__ uint32_t slot_off(uint32_t o)
{
  const uint32_t row = o >> 4, c16 = (o >> 1) & 7u;
  return (row << 7) + (((c16 ^ (row & 7u)) << 4) | ((o & 1u) << 3));
}


/* ---- per-pass register networks (twiddle pointers are 64-bit so sub-block offsets fold into immediates) --- */

__device__ __forceinline__ Mulc ld_tw(const uint4 *wu, const uint2 *qq, int off)
{
  const uint4 a = __ldg(wu + off);
  const uint2 b = __ldg(qq + off);
  return Mulc{a.x, a.y, a.z, a.w, b.x, b.y};
}
/* same entry from the shared-memory twiddle cache (plain loads: LDS.128 + LDS.64) */
__device__ __forceinline__ Mulc ld_tw_s(const uint4 *wu, const uint2 *qq, int off)
{
  const uint4 a = wu[off];
  const uint2 b = qq[off];
  return Mulc{a.x, a.y, a.z, a.w, b.x, b.y};
}

/*
 * R-stage network on x[0..2^R) for a group whose first stage is global stage s0 and whose block index at that
 * stage is blk0.  Forward: stages s0..s0+R-1, stage s0+u pairs x[k], x[k+d], d = 2^(R-1-u), sub-block `sub`
 * uses twiddle 2^(s0+u) + (blk0<<u) + sub.  Inverse: same stages backwards, Gentleman-Sande, bound constants
 * from p.inv_c[].  TOP_RENORM: this pass may be asked (p.inv_renorm_mask) to pull values below 3q first.
 */
template <int R, bool FWD>
__device__ __forceinline__ void ring_network(uint64_t (&x)[1 << R], const ntt_cuda_params_t &p, uint32_t s0,
                                             const uint4 *tw_wu, const uint2 *tw_qq)
{
  /* tw_wu / tw_qq: shared-memory twiddles of this group, entry 2^u-1+sub for stage s0+u, sub-block sub */
  constexpr int n = 1 << R;
  if(FWD) {
#pragma unroll
    for(int u = 0; u < R; u++) {
      const int d = n >> (u + 1);
#pragma unroll
      for(int sub = 0; sub < (1 << u); sub++) {
        const Mulc m = ld_tw_s(tw_wu, tw_qq, (1 << u) - 1 + sub);
#pragma unroll
        for(int k = 0; k < d; k++) bfly_fwd<false>(x[sub * 2 * d + k], x[sub * 2 * d + k + d], m, p, p.c10q);
      }
    }
  } else {
#pragma unroll
    for(int u = R - 1; u >= 0; u--) {
      const int      d  = n >> (u + 1);
      const uint32_t s  = s0 + u;
      const uint64_t cb = p.inv_c[s];
      if(u == R - 1 && ((p.inv_renorm_mask >> s) & 1u)) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
#pragma unroll
        for(int k = 0; k < n; k++) x[k] = reduce_2q(x[k], rc);
      }
      if(u == 0 && s == 0) {
        const Mulc a = mulc_from(p.ninv), b = mulc_from(p.ninv_w1);
#pragma unroll
        for(int k = 0; k < d; k++) bfly_inv_final<false>(x[k], x[k + d], a, b, p, cb);
      } else {
#pragma unroll
        for(int sub = 0; sub < (1 << u); sub++) {
          const Mulc m = ld_tw_s(tw_wu, tw_qq, (1 << u) - 1 + sub);
#pragma unroll
          for(int k = 0; k < d; k++) bfly_inv<false>(x[sub * 2 * d + k], x[sub * 2 * d + k + d], m, p, cb);
        }
      }
    }
  }
}

/* Pass C network: the last four stages on 16 contiguous coefficients, twiddles from the [15][groups]
 * table: entry t = 2^u-1+sub of group gidx sits at t*groups + gidx. */
template <bool FWD>
__device__ __forceinline__ void ring_network_c(uint64_t (&x)[16], const ntt_cuda_params_t &p, uint32_t s0,
                                               size_t gidx, size_t groups)
{
  const uint4 *wu = (const uint4 *)(FWD ? p.fwd_ct_wu : p.inv_ct_wu) + gidx;
  const uint2 *qq = (const uint2 *)(FWD ? p.fwd_ct_qq : p.inv_ct_qq) + gidx;
  if(FWD) {
#pragma unroll
    for(int u = 0; u < 4; u++) {
      const int d = 8 >> u;
#pragma unroll
      for(int sub = 0; sub < (1 << u); sub++) {
        const size_t t = (size_t)((1 << u) - 1 + sub) * groups;
        const Mulc   m = ld_tw(wu + t, qq + t, 0);
#pragma unroll
        for(int k = 0; k < d; k++) bfly_fwd<false>(x[sub * 2 * d + k], x[sub * 2 * d + k + d], m, p, p.c10q);
      }
    }
  } else {
#pragma unroll
    for(int u = 3; u >= 0; u--) {
      const int      d  = 8 >> u;
      const uint32_t s  = s0 + u;
      const uint64_t cb = p.inv_c[s];
      if(u == 3 && ((p.inv_renorm_mask >> s) & 1u)) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
#pragma unroll
        for(int k = 0; k < 16; k++) x[k] = reduce_2q(x[k], rc);
      }
#pragma unroll
      for(int sub = 0; sub < (1 << u); sub++) {
        const size_t t = (size_t)((1 << u) - 1 + sub) * groups;
        const Mulc   m = ld_tw(wu + t, qq + t, 0);
#pragma unroll
        for(int k = 0; k < d; k++) bfly_inv<false>(x[sub * 2 * d + k], x[sub * 2 * d + k + d], m, p, cb);
      }
    }
  }
}

/* ---- the kernel ----------------------------------------------------------------------------------------------- */

template <int L, bool FWD>
__global__ void __launch_bounds__(RingCfg<L>::THREADS, RingCfg<L>::CTAS)
  k_ring(const __grid_constant__ ntt_cuda_params_t p, const __grid_constant__ CUtensorMap tmap, size_t n_chunks,
         uint64_t *__restrict__ p_out)
{
  using C = RingCfg<L>;
  constexpr int NB = C::NB, RA = C::RA, SLOTS = C::SLOTS, T = C::THREADS, HALF = NB / 2;
  extern __shared__ uint8_t smem_raw[];
  /* slots need 1024-byte alignment for SWIZZLE_128B; barriers sit behind the ring */
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *      ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  uint4 *        tw_wu = reinterpret_cast<uint4 *>(ring_ptr + SLOTS * 4096);       /* NTW x 16 bytes */
  uint2 *        tw_qq = reinterpret_cast<uint2 *>(ring_ptr + SLOTS * 4096 + C::NTW * 16); /* NTW x 8 bytes */
  const uint32_t bars  = ring + SLOTS * 4096 + C::TW_BYTES;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t s1       = p.logn - L;                       /* stages already done by strided passes */
  const size_t   my_polys = (n_chunks > blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const size_t   my_blocks = my_polys * NB;
  const size_t   groups   = (size_t)1 << (p.logn - 4);        /* 16-coefficient runs per polynomial */

  auto slot_addr = [&](size_t g) -> uint32_t { return ring + (uint32_t)(g % SLOTS) * 4096u; };
  /* arm + issue the TMA load of block g of this CTA's sequence (no-op past the end) */
  auto issue_load = [&](size_t g) {
    if(g >= my_blocks) return;
    const size_t   k     = g / NB;
    const uint32_t b     = (uint32_t)(g % NB);
    const size_t   chunk = blockIdx.x + k * gridDim.x;
    const uint32_t bar   = bars + 8u * (uint32_t)(k % C::NBAR);
    mbar_arrive_expect_tx(bar, 4096u);
    tma_load_block(slot_addr(g), &tmap, (int)((chunk << (L - 4)) + b * 32u), bar);
  };

  if(tid == 0) {
    tma_prefetch_desc(&tmap);
    for(int i = 0; i < C::NBAR; i++) mbar_init(bars + 8u * i, NB);
    fence_barrier_init();
  }
  __syncthreads();
  /* prologue: fill the ring */
  for(uint32_t g = tid; g < (uint32_t)SLOTS; g += T) issue_load(g);

  uint32_t cached_cp = 0xffffffffu;
  uint32_t sl_next = 0; /* slot of block 0 of the next polynomial: (k * NB) mod SLOTS, kept in 32 bits */
  for(size_t k = 0; k < my_polys; k++) {
    const size_t   chunk = blockIdx.x + k * gridDim.x;
    const uint32_t cp    = (uint32_t)(chunk & (((size_t)1 << s1) - 1)); /* chunk index inside its polynomial */
    if(cp != cached_cp) {
      /* (re)fill the pass A / pass B twiddle cache for this chunk position: pass A entry 2^u-1+sub is global
       * 2^(s1+u) + (cp<<u) + sub; pass B, block b: 2^(s1+RA+u) + ((cp*NB+b)<<u) + sub.  When N = 2^L this
       * happens once per CTA. */
      const uint4 *gwu = (const uint4 *)(FWD ? p.fwd_wu : p.inv_wu);
      const uint2 *gqq = (const uint2 *)(FWD ? p.fwd_qq : p.inv_qq);
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
        const size_t   src = ((size_t)1 << (st + u)) + ((size_t)blk << u) + sub;
        tw_wu[e] = __ldg(gwu + src);
        tw_qq[e] = __ldg(gqq + src);
      }
      cached_cp = cp;
      __syncthreads();
    }
    const size_t   g0    = k * NB;
    /* SLOTS = 3 * HALF: a polynomial's low and high halves each sit in HALF consecutive slots, so a block address
     * is one of two uniform bases plus a compile-time offset */
    const uint32_t sl0 = sl_next, sh0 = sl0 + HALF >= (uint32_t)SLOTS ? sl0 + HALF - SLOTS : sl0 + HALF;
    sl_next            = sl0 + NB >= (uint32_t)SLOTS ? sl0 + NB - SLOTS : sl0 + NB;
    mbar_wait(bars + 8u * (uint32_t)(k % C::NBAR), (uint32_t)((k / C::NBAR) & 1));

    /* slot of block b of this polynomial (uniform arithmetic: one compare per block) */
    auto blk_slot = [&](uint32_t b) -> uint32_t { return b < (uint32_t)HALF ? sl0 + b : sh0 + (b - HALF); };

    /* ---------- pass A (forward first, inverse last): across blocks ---------- */
    auto pass_a = [&]() {
      for(uint32_t j = tid; j < 512u; j += T) {
        const uint32_t off = slot_off(j);
        uint64_t       x[NB];
#pragma unroll
        for(int b = 0; b < NB; b++) x[b] = *reinterpret_cast<const uint64_t *>(ring_ptr + blk_slot(b) * 4096u + off);
        ring_network<RA, FWD>(x, p, s1, tw_wu, tw_qq);
        if(!FWD && s1 == 0) {
          const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
#pragma unroll
          for(int b = 0; b < NB; b++) x[b] = reduce_full(x[b], rc);
        }
#pragma unroll
        for(int b = 0; b < NB; b++) *reinterpret_cast<uint64_t *>(ring_ptr + blk_slot(b) * 4096u + off) = x[b];
      }
    };

    /* inverse pass A: columns into registers, one __syncthreads, slots re-armed at once, results written from
     * registers to global memory (256 contiguous bytes per warp store) -- see ntt_ring_fp.cuh for the why */
    auto pass_a_inv = [&]() {
      constexpr int COLS = 512 / T;
      uint64_t      x[COLS][NB];
#pragma unroll
      for(int cidx = 0; cidx < COLS; cidx++) {
        const uint32_t off = slot_off(tid + cidx * T);
#pragma unroll
        for(int b = 0; b < NB; b++)
          x[cidx][b] = *reinterpret_cast<const uint64_t *>(ring_ptr + blk_slot(b) * 4096u + off);
      }
      __syncthreads();
      if(lane == 0) {
        issue_load(g0 + warp + SLOTS);
        issue_load(g0 + warp + HALF + SLOTS);
      }
      uint64_t *gout = p_out + (chunk << L);
#pragma unroll
      for(int cidx = 0; cidx < COLS; cidx++) {
        ring_network<RA, false>(x[cidx], p, s1, tw_wu, tw_qq);
        const uint32_t j = tid + cidx * T;
        const Red      rc{p.q, p.negq, p.red_shift, p.red_mu};
#pragma unroll
        for(int b = 0; b < NB; b++) gout[(size_t)b * 512 + j] = (s1 == 0) ? reduce_full(x[cidx][b], rc) : x[cidx][b];
      }
    };

    /* ---------- pass B: inside a block, 32 coefficients at stride 16; half-warp h owns block warp + h*HALF ---- */
    const uint32_t hb   = lane >> 4, jb = lane & 15u;
    const uint32_t blkB = warp + hb * HALF;
    auto pass_b = [&]() {
      uint8_t *      base = ring_ptr + blk_slot(blkB) * 4096u + ((jb & 1u) << 3);
      const uint32_t jc   = jb >> 1;
      uint64_t       x[32];
#pragma unroll
      for(int kk = 0; kk < 32; kk++)
        x[kk] = *reinterpret_cast<const uint64_t *>(base + kk * 128 + ((jc ^ (uint32_t)(kk & 7)) << 4));
      ring_network<5, FWD>(x, p, s1 + RA, tw_wu + (NB - 1) + blkB * 31, tw_qq + (NB - 1) + blkB * 31);
#pragma unroll
      for(int kk = 0; kk < 32; kk++)
        *reinterpret_cast<uint64_t *>(base + kk * 128 + ((jc ^ (uint32_t)(kk & 7)) << 4)) = x[kk];
    };

    /* ---------- pass C: one swizzled 128-byte row (16 contiguous coefficients) per lane ---------- */
    auto pass_c = [&](uint32_t blk) {
      uint8_t *base = ring_ptr + blk_slot(blk) * 4096u + lane * 128u;
      uint64_t x[16];
#pragma unroll
      for(int c = 0; c < 8; c++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(base + (((uint32_t)c ^ (lane & 7u)) << 4));
        x[2 * c]           = v.x;
        x[2 * c + 1]       = v.y;
      }
      ring_network_c<FWD>(x, p, s1 + RA + 5, ((size_t)cp * NB + blk) * 32 + lane, groups);
      if(FWD) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
#pragma unroll
        for(int i = 0; i < 16; i++) x[i] = reduce_full(x[i], rc);
      }
#pragma unroll
      for(int c = 0; c < 8; c++) {
        ulonglong2 v;
        v.x = x[2 * c];
        v.y = x[2 * c + 1];
        *reinterpret_cast<ulonglong2 *>(base + (((uint32_t)c ^ (lane & 7u)) << 4)) = v;
      }
    };

    /* store block b of this polynomial from its slot (called by one lane after fence + warp sync) */
    auto store_block = [&](uint32_t b) {
      tma_store_block(&tmap, (int)((chunk << (L - 4)) + b * 32u), ring + blk_slot(b) * 4096u);
      tma_commit();
    };

    if(FWD) {
      pass_a();
      __syncthreads();
      pass_b();
      __syncwarp();
      pass_c(warp);
      fence_proxy_async();
      __syncwarp();
      if(lane == 0) store_block(warp);
      pass_c(warp + HALF);
      fence_proxy_async();
      __syncwarp();
      if(lane == 0) {
        store_block(warp + HALF);
        /* both slots are re-armed once the stores have drained them */
        tma_wait_read_all();
        issue_load(g0 + warp + SLOTS);
        issue_load(g0 + warp + HALF + SLOTS);
      }
      __syncwarp();
    } else {
      pass_c(warp);
      pass_c(warp + HALF);
      __syncwarp();
      pass_b();
      __syncthreads();
      pass_a_inv();
    }
  }
  /* the kernel may not exit while its bulk stores are still in flight */
  tma_wait_all();
}

}  // namespace nttb200
