/*
 * csrc/ntt_device.cuh -- device arithmetic for the negacyclic NTT on sm_100a.
 *
 * The B200 integer datapath is 32 bits wide: IMAD and IMAD.WIDE.U32 issue at 64 lanes/clk/SM, IMAD.HI at
 * half of that (measured, profiles/r01_ubench_pipes.txt).  Everything below is therefore expressed in
 * IMAD.WIDE / IMAD.LO and 64-bit add chains; no mul.hi.
 *
 * Two multiplier forms implement the reference's Shoup product fast_mul_mod_q2
 * (include/internal/fast_mul_operators.h:49-54):
 *
 *  (1) LAZY SPLIT form (default; any q with (4+10*log2N)*q < 2^64, i.e. every q below 2^56):
 *      a multiplier w is stored as  w, u = w*2^32 mod q, wq = floor(w*2^30/q), uq = floor(u*2^30/q).
 *      For y = y1*2^32 + y0:   w*y == w*y0 + u*y1 =: V (mod q), V < q*2^33.
 *      S = floor((wq*y0 + uq*y1) / 2^31) satisfies V/(2q) - 5 < S <= V/(2q), so
 *      r = V - S*2q lies in [0, 10q) and r == w*y (mod q).  Cost: 8 IMAD (2 for S, 6 for r mod 2^64)
 *      + 1 funnel shift, against 10 IMAD + carry adds for the textbook 64-bit Shoup product.
 *      Because r < 10q regardless of y, butterflies need NO conditional subtraction: values simply
 *      grow by 10q per forward stage and are brought back once at the end (reduce_full).
 *
 *  (2) EXACT form (q up to 2^62): w and c = floor(w*2^64/q); r = w*y - hi64(c*y)*q in [0,2q), the
 *      reference's arithmetic bit for bit, with Harvey's per-stage conditional subtractions
 *      (fast_mul_operators.h:72-106).
 *
 * Parity is defined after full reduction to [0,q) (tests/test_correctness.c:267-269), so the lazy
 * representation is free to differ from the reference's as long as the residues agree.
 */
#pragma once
#include <cstdint>

#include "ntt_cuda.h"

namespace nttb200 {

struct Mulc {  // one multiplier, register-resident
  uint32_t w0, w1, u0, u1, wq, uq;
};

__device__ __forceinline__ uint32_t lo32(uint64_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi32(uint64_t v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi)
{
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t mul_wide(uint32_t a, uint32_t b)
{
  uint64_t r;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c)
{
  uint64_t r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

/* r == m*y (mod q), r in [0,10q).  n2q = 2^64 - 2q.  8 full-rate IMAD + 1 SHF.
 * A = wq*y0 + uq*y1 < 2^63 (30-bit companions); S = floor(A/2^31) obeys V/(2q) - 5 < S <= V/(2q).
 * Shifting by 31 rather than taking the high word keeps both halves of A live, so ptxas cannot fuse the
 * second IMAD.WIDE into a half-rate IMAD.HI. */
__device__ __forceinline__ uint64_t mul_lazy_acc(uint64_t y, const Mulc &m, uint64_t n2q, uint64_t acc)
{
  const uint32_t y0 = lo32(y), y1 = hi32(y);
  const uint64_t a  = mad_wide(m.uq, y1, mul_wide(m.wq, y0));
  const uint32_t S  = __funnelshift_r(lo32(a), hi32(a), 31);
  uint64_t       r  = mad_wide(m.w0, y0, acc);
  r                 = mad_wide(m.u0, y1, r);
  r                 = mad_wide(S, lo32(n2q), r);
  uint32_t rh       = hi32(r);
  rh                = mad_lo(m.w1, y0, rh);
  rh                = mad_lo(m.u1, y1, rh);
  rh                = mad_lo(S, hi32(n2q), rh);
  return pack64(lo32(r), rh);
}
__device__ __forceinline__ uint64_t mul_lazy(uint64_t y, const Mulc &m, uint64_t n2q)
{
  return mul_lazy_acc(y, m, n2q, 0);
}

/* hi64(a*b) from four IMAD.WIDE (exact) */
__device__ __forceinline__ uint64_t mulhi64(uint64_t a, uint64_t b)
{
  const uint32_t a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
  const uint64_t p00 = mul_wide(a0, b0);
  const uint64_t p01 = mad_wide(a0, b1, (uint64_t)hi32(p00));
  const uint64_t p10 = mad_wide(a1, b0, (uint64_t)lo32(p01));
  return mad_wide(a1, b1, (uint64_t)hi32(p01) + (uint64_t)hi32(p10));
}

/* exact Shoup product: r = w*y - hi64(c*y)*q in [0,2q); (u0,u1) hold c = floor(w*2^64/q) */
__device__ __forceinline__ uint64_t mul_exact(uint64_t y, const Mulc &m, uint64_t q)
{
  const uint64_t w = pack64(m.w0, m.w1), c = pack64(m.u0, m.u1);
  return w * y - mulhi64(c, y) * q;
}

/* v - m if v >= m else v, for v, m < 2^63 (sign test of the difference: 2 adds, 1 compare, 2 selects) */
__device__ __forceinline__ uint64_t csub(uint64_t v, uint64_t m)
{
  const uint64_t t = v - m;
  return ((int64_t)t < 0) ? v : t;
}

struct Red {  // reduction constants
  uint64_t q, negq;
  uint32_t shift, mu;
};

/* v < 2^(bitlen(q)+22)  ->  v mod q up to one multiple: result in [0,2q).
 * shift = max(0, bitlen(q)-9), mu = floor(2^(32+shift)/q) < 2^24: the quotient estimate
 * Q = floor(floor(v/2^shift) * mu / 2^32) is floor(v/q) or one less.  1 funnel shift + 3 IMAD. */
__device__ __forceinline__ uint64_t reduce_2q(uint64_t v, const Red &c)
{
  const uint32_t vt = (uint32_t)(v >> c.shift);
  const uint32_t Q  = hi32(mul_wide(vt, c.mu));
  uint64_t       r  = mad_wide(Q, lo32(c.negq), v);
  return pack64(lo32(r), mad_lo(Q, hi32(c.negq), hi32(r)));
}
/* v -> v mod q in [0,q) */
__device__ __forceinline__ uint64_t reduce_full(uint64_t v, const Red &c) { return csub(reduce_2q(v, c), c.q); }

__device__ __forceinline__ Mulc load_mulc(const uint4 *wu, const uint2 *qq, uint32_t idx)
{
  const uint4 a = __ldg(wu + idx);
  const uint2 b = __ldg(qq + idx);
  return Mulc{a.x, a.y, a.z, a.w, b.x, b.y};
}
__device__ __forceinline__ Mulc load_mulc_exact(const uint4 *wu, uint32_t idx)
{
  const uint4 a = __ldg(wu + idx);
  return Mulc{a.x, a.y, a.z, a.w, 0u, 0u};
}
__device__ __forceinline__ Mulc mulc_from(const ntt_cuda_mulc_t &m)
{
  return Mulc{m.w0, m.w1, m.u0, m.u1, m.wq, m.uq};
}

/* ---- butterflies -------------------------------------------------------------------------------- */

/* forward (Cooley-Tukey) butterfly, harvey_fwd_butterfly fast_mul_operators.h:72-81.
 * lazy: X' = X + T, Y' = X - T + 10q with T in [0,10q): both outputs < X + 10q.  (ptxas sums the three
 * 64-bit partial products with one 3-input add pair either way, so T is formed first: 6 add instructions.) */
template <bool EXACT>
__device__ __forceinline__ void bfly_fwd(uint64_t &x, uint64_t &y, const Mulc &m, const ntt_cuda_params_t &p,
                                         uint64_t c10q)
{
  if(EXACT) {
    const uint64_t q2 = p.q << 1;
    const uint64_t x1 = csub(x, q2);
    const uint64_t t  = mul_exact(y, m, p.q);
    x                 = x1 + t;
    y                 = x1 - t + q2;
  } else {
    const uint64_t t = mul_lazy(y, m, p.neg2q);
    y                = x - t + c10q;
    x                = x + t;
  }
}

/* inverse (Gentleman-Sande) butterfly, harvey_bkw_butterfly fast_mul_operators.h:83-92.
 * lazy: X' = X + Y (< 2B q), Y' = m * (X - Y + B q) in [0,10q), where cb = B*q bounds the inputs. */
template <bool EXACT>
__device__ __forceinline__ void bfly_inv(uint64_t &x, uint64_t &y, const Mulc &m, const ntt_cuda_params_t &p,
                                         uint64_t cb)
{
  if(EXACT) {
    const uint64_t q2 = p.q << 1;
    const uint64_t s  = csub(x + y, q2);
    const uint64_t d  = x - y + q2;
    x                 = s;
    y                 = mul_exact(d, m, p.q);
  } else {
    const uint64_t d = x - y + cb;
    x                = x + y;
    y                = mul_lazy(d, m, p.neg2q);
  }
}

/* last inverse stage with N^-1 folded in, harvey_bkw_butterfly_final fast_mul_operators.h:94-106.
 * Outputs are left lazy (< 10q, or < 2q exact); the caller applies the final reduction. */
template <bool EXACT>
__device__ __forceinline__ void bfly_inv_final(uint64_t &x, uint64_t &y, const Mulc &ninv, const Mulc &ninv_w,
                                               const ntt_cuda_params_t &p, uint64_t cb)
{
  if(EXACT) {
    const uint64_t q2 = p.q << 1;
    const uint64_t s  = x + y;
    const uint64_t d  = x - y + q2;
    x                 = mul_exact(s, ninv, p.q);
    y                 = mul_exact(d, ninv_w, p.q);
  } else {
    const uint64_t s = x + y;
    const uint64_t d = x - y + cb;
    x                = mul_lazy(s, ninv, p.neg2q);
    y                = mul_lazy(d, ninv_w, p.neg2q);
  }
}

}  // namespace nttb200
