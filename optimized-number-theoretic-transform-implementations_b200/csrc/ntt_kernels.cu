/*
 * csrc/ntt_kernels.cu -- sm_100a kernels and their extern "C" launchers (see ntt_cuda.h).
 *
 * Kernel inventory
 *   k_chunk<L,...>     one CTA transforms one contiguous chunk of 2^L coefficients (L <= 14, i.e. up to a
 *                      whole N = 2^14 polynomial) held in shared memory: up to three register-tiled
 *                      passes (radix 2^5, 2^5, 2^4) with one __syncthreads between passes.
 *   k_strided<R,...>   register-only radix-2^R pass over global memory for the stages whose butterfly
 *                      distance exceeds a chunk (N >= 2^15): the "two-kernel split" of the north star.
 *                      Each thread owns 2^R coefficients at stride N/2^(s0+R); warps read and write
 *                      consecutive addresses.
 *   k_build_tables, k_gen_roots   on-device twiddle table generation
 *                      (replaces calc_w / calc_w_con, include/internal/pre_compute.h:38-77).
 *   k_pointwise        NTT-domain product.
 *
 * Stage/twiddle schedule (restates src/ntt_reference.c:19-30 and :43-53): forward stage s has 2^s blocks
 * of 2t coefficients, t = N/2^(s+1); block i multiplies its upper half by w[2^s + i].  The inverse runs
 * the stages backwards with Gentleman-Sande butterflies and the tables of psi^-1, and folds N^-1 into
 * stage 0.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include <cuda.h>

#include <cstdlib>

#include <mutex>

#include "ntt_cuda.h"
#include "ntt_device.cuh"
#include "ntt_launch.h"

using namespace nttb200;

/* ------------------------------------------------------------------------------------------------ */
/* error plumbing                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

static thread_local char g_err[512] = "";

static int fail(const char *what, cudaError_t e)
{
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -1;
}
static int fail_msg(const char *what)
{
  snprintf(g_err, sizeof(g_err), "%s", what);
  return -1;
}
#define CU(call)                             \
  do {                                       \
    cudaError_t e_ = (call);                 \
    if(e_ != cudaSuccess) return fail(#call, e_); \
  } while(0)

extern "C" const char *ntt_cuda_error(void) { return g_err; }
namespace nttb200 {
int nl_fail(const char *what, cudaError_t e) { return fail(what, e); }
int nl_fail_msg(const char *what) { return fail_msg(what); }
}  // namespace nttb200

extern "C" int ntt_cuda_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

/* ------------------------------------------------------------------------------------------------ */
/* shared-memory layout                                                                              */
/* ------------------------------------------------------------------------------------------------ */

/* Coefficient e of a chunk lives at 8-byte slot swz(e): the 16-byte unit index (e>>1) has its low three
 * bits XORed with bits 3..5 of itself, i.e. exactly the TMA SWIZZLE_128B pattern over 128-byte rows.
 * Effect: a thread reading 16 consecutive coefficients (one 128-byte row) with LDS.128 hits a different
 * bank group than its seven neighbours, and strided passes still see whole rows. */
__device__ __forceinline__ uint32_t swz(uint32_t e) { return e ^ (((e >> 4) & 7u) << 1); }

/* ------------------------------------------------------------------------------------------------ */
/* register-tiled radix-2^R butterfly network                                                        */
/* ------------------------------------------------------------------------------------------------ */

/* Twiddle fetch for global stage s, block blk */
template <bool EXACT>
__device__ __forceinline__ Mulc tw(const ntt_cuda_params_t &p, bool fwd, uint32_t s, uint32_t blk)
{
  const uint32_t idx = (1u << s) + blk;
  const uint4 *  wu  = (const uint4 *)(fwd ? p.fwd_wu : p.inv_wu);
  if(EXACT) return load_mulc_exact(wu, idx);
  const uint2 *qq = (const uint2 *)(fwd ? p.fwd_qq : p.inv_qq);
  return load_mulc(wu, qq, idx);
}

/*
 * x[0..2^R) are the coefficients of one group: positions base + k*es of a block that starts stage s0.
 * Forward: stages s0 .. s0+R-1; stage s0+u pairs x[k], x[k+d] with d = 2^(R-1-u); the 2^u sub-blocks use
 * twiddle block index (blk0 << u) + sub.  Inverse: the same stages in the opposite order.
 */
template <int R, bool FWD, bool EXACT>
__device__ __forceinline__ void radix_network(uint64_t (&x)[1 << R], const ntt_cuda_params_t &p, uint32_t s0,
                                              uint32_t blk0)
{
  constexpr int n   = 1 << R;
  const uint64_t c10 = p.c10q;
  if(FWD) {
#pragma unroll
    for(int u = 0; u < R; u++) {
      const int d = n >> (u + 1);
#pragma unroll
      for(int sub = 0; sub < (1 << u); sub++) {
        const Mulc m = tw<EXACT>(p, true, s0 + u, (blk0 << u) + sub);
#pragma unroll
        for(int k = 0; k < d; k++) bfly_fwd<EXACT>(x[sub * 2 * d + k], x[sub * 2 * d + k + d], m, p, c10);
      }
    }
  } else {
#pragma unroll
    for(int u = R - 1; u >= 0; u--) {
      const int      d  = n >> (u + 1);
      const uint32_t s  = s0 + u;
      const uint64_t cb = p.inv_c[s];
      /* renormalisation is only ever scheduled on the first stage a pass runs (u == R-1) */
      if(!EXACT && u == R - 1 && ((p.inv_renorm_mask >> s) & 1u)) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
#pragma unroll
        for(int k = 0; k < n; k++) x[k] = reduce_2q(x[k], rc);
      }
      if(u == 0 && s == 0) {
        /* global stage 0 (only reachable with blk0 == 0, u == 0): N^-1 folded in */
        const Mulc a = mulc_from(p.ninv), b = mulc_from(p.ninv_w1);
#pragma unroll
        for(int k = 0; k < d; k++) bfly_inv_final<EXACT>(x[k], x[k + d], a, b, p, cb);
      } else {
#pragma unroll
        for(int sub = 0; sub < (1 << u); sub++) {
          const Mulc m = tw<EXACT>(p, false, s, (blk0 << u) + sub);
#pragma unroll
          for(int k = 0; k < d; k++) bfly_inv<EXACT>(x[sub * 2 * d + k], x[sub * 2 * d + k + d], m, p, cb);
        }
      }
    }
  }
}

template <bool EXACT>
__device__ __forceinline__ uint64_t finish(uint64_t v, const ntt_cuda_params_t &p)
{
  if(EXACT) return csub(csub(v, p.q << 1), p.q); /* 4q < 2^64 and q < 2^62: differences stay below 2^63 */
  const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
  return reduce_full(v, rc);
}

/* ------------------------------------------------------------------------------------------------ */
/* chunk kernel                                                                                      */
/* ------------------------------------------------------------------------------------------------ */

template <int L>
struct ChunkCfg {
  static constexpr int RC      = L < 4 ? L : 4;                     /* last pass: contiguous 2^RC per thread */
  static constexpr int RB      = (L - RC) < 5 ? (L - RC) : 5;       /* middle pass */
  static constexpr int RA      = L - RC - RB;                       /* first pass */
  static constexpr int RMAX    = RA > RB ? (RA > RC ? RA : RC) : (RB > RC ? RB : RC);
  static constexpr int GROUPS  = 1 << (L - RMAX);                   /* groups in the widest pass */
  static constexpr int THREADS = GROUPS < 32 ? 32 : GROUPS;
  /* resident CTAs per SM the register allocator should leave room for (128 registers per thread) */
  static constexpr int MINB    = (512 / THREADS) < 16 ? (512 / THREADS) : 16;
  static_assert(RA <= 5, "chunk too large");
};

/* One pass over the chunk in shared memory: local stages ls0 .. ls0+R-1 (forward order). */
template <int L, int R, bool FWD, bool EXACT, bool FINISH, int THREADS>
__device__ __forceinline__ void smem_pass(uint64_t *sm, const ntt_cuda_params_t &p, uint32_t ls0, uint32_t s1,
                                          uint32_t chunk_in_poly)
{
  if(R == 0) return;
  constexpr int n      = 1 << R;
  constexpr int groups = 1 << (L - R);
  const uint32_t es_log = L - ls0 - R; /* log2 of the element stride inside a group */
  for(uint32_t g = threadIdx.x; g < (uint32_t)groups; g += THREADS) {
    const uint32_t i    = g >> es_log;
    const uint32_t j    = g & ((1u << es_log) - 1u);
    const uint32_t base = (i << (L - ls0)) + j;
    uint64_t       x[n];
    if(R == 4 && ls0 + R == (uint32_t)L) {
      /* contiguous 16 coefficients = one swizzled 128-byte row: eight 16-byte accesses */
      const uint32_t row = base >> 4;
#pragma unroll
      for(int c = 0; c < 8; c++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sm + (row << 4) + (((uint32_t)c ^ (row & 7u)) << 1));
        x[2 * c]           = v.x;
        x[2 * c + 1]       = v.y;
      }
    } else {
#pragma unroll
      for(int k = 0; k < n; k++) x[k] = sm[swz(base + ((uint32_t)k << es_log))];
    }
    radix_network<R, FWD, EXACT>(x, p, s1 + ls0, (chunk_in_poly << ls0) + i);
    if(FINISH) {
#pragma unroll
      for(int k = 0; k < n; k++) x[k] = finish<EXACT>(x[k], p);
    }
    if(R == 4 && ls0 + R == (uint32_t)L) {
      const uint32_t row = base >> 4;
#pragma unroll
      for(int c = 0; c < 8; c++) {
        ulonglong2 v;
        v.x = x[2 * c];
        v.y = x[2 * c + 1];
        *reinterpret_cast<ulonglong2 *>(sm + (row << 4) + (((uint32_t)c ^ (row & 7u)) << 1)) = v;
      }
    } else {
#pragma unroll
      for(int k = 0; k < n; k++) sm[swz(base + ((uint32_t)k << es_log))] = x[k];
    }
  }
}

/*
 * Forward: the chunk is block `chunk_in_poly` of global stage s1 (s1 = logn - L stages were already done by
 * k_strided); runs local stages 0..L-1 and the final reduction.  Inverse: runs local stages L-1..0; if
 * s1 == 0 that includes global stage 0 and the final reduction, otherwise values stay lazy for k_strided.
 */
template <int L, bool FWD, bool EXACT>
__global__ void __launch_bounds__(ChunkCfg<L>::THREADS, ChunkCfg<L>::MINB) k_chunk(const __grid_constant__ ntt_cuda_params_t p,
                                                                  uint64_t *__restrict__ a, size_t n_chunks)
{
  using C = ChunkCfg<L>;
  extern __shared__ __align__(1024) uint64_t sm[];
  constexpr uint32_t n   = 1u << L;
  const uint32_t     s1  = p.logn - L;
  constexpr int      T   = C::THREADS;

  for(size_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    uint64_t *     g  = a + chunk * n;
    const uint32_t cp = (uint32_t)(chunk & ((1u << s1) - 1u));

    /* global -> shared, 16 bytes per access, swizzled */
    if(L >= 1) {
      for(uint32_t u = threadIdx.x; u < n / 2; u += T) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(g + 2 * u);
        *reinterpret_cast<ulonglong2 *>(sm + 2 * (u ^ ((u >> 3) & 7u))) = v;
      }
    }
    __syncthreads();

    if(FWD) {
      smem_pass<L, C::RA, true, EXACT, false, T>(sm, p, 0, s1, cp);
      if(C::RA) __syncthreads();
      smem_pass<L, C::RB, true, EXACT, false, T>(sm, p, C::RA, s1, cp);
      if(C::RB) __syncthreads();
      smem_pass<L, C::RC, true, EXACT, true, T>(sm, p, C::RA + C::RB, s1, cp);
    } else {
      if(s1 == 0) {
        /* the pass that contains global stage 0 also applies the final reduction */
        if(C::RA) {
          smem_pass<L, C::RC, false, EXACT, false, T>(sm, p, C::RA + C::RB, s1, cp);
          __syncthreads();
          smem_pass<L, C::RB, false, EXACT, false, T>(sm, p, C::RA, s1, cp);
          __syncthreads();
          smem_pass<L, C::RA, false, EXACT, true, T>(sm, p, 0, s1, cp);
        } else if(C::RB) {
          smem_pass<L, C::RC, false, EXACT, false, T>(sm, p, C::RA + C::RB, s1, cp);
          __syncthreads();
          smem_pass<L, C::RB, false, EXACT, true, T>(sm, p, C::RA, s1, cp);
        } else {
          smem_pass<L, C::RC, false, EXACT, true, T>(sm, p, C::RA + C::RB, s1, cp);
        }
      } else {
        smem_pass<L, C::RC, false, EXACT, false, T>(sm, p, C::RA + C::RB, s1, cp);
        if(C::RB) __syncthreads();
        smem_pass<L, C::RB, false, EXACT, false, T>(sm, p, C::RA, s1, cp);
        if(C::RA) __syncthreads();
        smem_pass<L, C::RA, false, EXACT, false, T>(sm, p, 0, s1, cp);
      }
    }
    __syncthreads();

    for(uint32_t u = threadIdx.x; u < n / 2; u += T) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sm + 2 * (u ^ ((u >> 3) & 7u)));
      *reinterpret_cast<ulonglong2 *>(g + 2 * u) = v;
    }
    __syncthreads();
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* strided global pass                                                                               */
/* ------------------------------------------------------------------------------------------------ */

/*
 * Global stages s0 .. s0+R-1 of every polynomial (forward order; the inverse runs them backwards).
 * Thread <-> group g of a polynomial: es = N >> (s0+R), block i = g / es, offset j = g % es, coefficients
 * at i*(es<<R) + j + k*es.  Consecutive threads have consecutive j, so each of the 2^R loads/stores of a
 * warp covers 256 contiguous bytes.
 */
/* OUT: 0 = leave lazy, 1 = final reduction to [0,q), 2 = bring below 2q (what the FP64 chunk kernel accepts) */
template <int R, bool FWD, bool EXACT, int OUT>
__global__ void __launch_bounds__(256) k_strided(const __grid_constant__ ntt_cuda_params_t p,
                                                 uint64_t *__restrict__ a, uint32_t s0, size_t first_group,
                                                 size_t n_groups)
{
  /* groups [first_group, first_group + n_groups) of the array that starts at `a`; a multi-GPU tail pass hands
   * in a virtual base so that only the groups of its own block are touched */
  constexpr int  n      = 1 << R;
  const uint32_t logn   = p.logn;
  const uint32_t es_log = logn - s0 - R;
  const uint32_t gl     = logn - R; /* log2 groups per polynomial */
  for(size_t t = first_group + (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < first_group + n_groups;
      t += (size_t)gridDim.x * blockDim.x) {
    const size_t   poly = t >> gl;
    const uint32_t g    = (uint32_t)(t & ((1u << gl) - 1u));
    const uint32_t i    = g >> es_log;
    const uint32_t j    = g & ((1u << es_log) - 1u);
    uint64_t *     base = a + (poly << logn) + ((size_t)i << (logn - s0)) + j;
    uint64_t       x[n];
#pragma unroll
    for(int k = 0; k < n; k++) x[k] = base[(size_t)k << es_log];
    radix_network<R, FWD, EXACT>(x, p, s0, i);
#pragma unroll
    for(int k = 0; k < n; k++) {
      uint64_t v = x[k];
      if(OUT == 1) v = finish<EXACT>(v, p);
      if(OUT == 2 && !EXACT) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
        v = reduce_2q(v, rc);
      }
      base[(size_t)k << es_log] = v;
    }
  }
}

/* Two adjacent groups per thread (16-byte loads and stores; both groups sit in the same block and share their
 * twiddles): the pass is HBM-bound and the 8-byte version reaches 4.8 TB/s (profiles/r02e_ncu_strided.txt).
 * Needs an even group stride (es >= 2), even first_group / n_groups and 16-byte aligned data; R <= 4. */
template <int R, bool FWD, bool EXACT, int OUT>
__global__ void __launch_bounds__(256) k_strided_v2(const __grid_constant__ ntt_cuda_params_t p,
                                                    uint64_t *__restrict__ a, uint32_t s0, size_t first_group,
                                                    size_t n_groups)
{
  constexpr int  n      = 1 << R;
  const uint32_t logn   = p.logn;
  const uint32_t es_log = logn - s0 - R;
  const uint32_t gl     = logn - R;
  for(size_t t2 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t2 < n_groups / 2; t2 += (size_t)gridDim.x * blockDim.x) {
    const size_t   t    = first_group + 2 * t2;
    const size_t   poly = t >> gl;
    const uint32_t g    = (uint32_t)(t & ((1u << gl) - 1u));
    const uint32_t i    = g >> es_log;
    const uint32_t j    = g & ((1u << es_log) - 1u);
    uint64_t *     base = a + (poly << logn) + ((size_t)i << (logn - s0)) + j;
    uint64_t       x0[n], x1[n];
#pragma unroll
    for(int k = 0; k < n; k++) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(base + ((size_t)k << es_log));
      x0[k] = v.x;
      x1[k] = v.y;
    }
    radix_network<R, FWD, EXACT>(x0, p, s0, i);
    radix_network<R, FWD, EXACT>(x1, p, s0, i);
#pragma unroll
    for(int k = 0; k < n; k++) {
      ulonglong2 v = make_ulonglong2(x0[k], x1[k]);
      if(OUT == 1) {
        v.x = finish<EXACT>(v.x, p);
        v.y = finish<EXACT>(v.y, p);
      }
      if(OUT == 2 && !EXACT) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
        v.x = reduce_2q(v.x, rc);
        v.y = reduce_2q(v.y, rc);
      }
      *reinterpret_cast<ulonglong2 *>(base + ((size_t)k << es_log)) = v;
    }
  }
}

/* The same pass over the polynomials of SEVERAL plans in one launch (RNS limbs, lazy path): polynomial `poly` of
 * the array belongs to plan poly / polys_per_limb, whose parameters come from the argument table. */
template <int R, bool FWD, int OUT>
__global__ void __launch_bounds__(256) k_strided_multi(const __grid_constant__ RingLimbs<true> limbs,
                                                       uint64_t *__restrict__ a, uint32_t s0, size_t n_groups)
{
  constexpr int  n      = 1 << R;
  const uint32_t logn   = limbs.e[0].logn;
  const uint32_t es_log = logn - s0 - R;
  const uint32_t gl     = logn - R;
  for(size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_groups; t += (size_t)gridDim.x * blockDim.x) {
    const size_t             poly = t >> gl;
    const ntt_cuda_params_t &p    = limbs.e[(uint32_t)poly / limbs.polys_per_limb];
    const uint32_t           g    = (uint32_t)(t & ((1u << gl) - 1u));
    const uint32_t           i    = g >> es_log;
    const uint32_t           j    = g & ((1u << es_log) - 1u);
    uint64_t *               base = a + (poly << logn) + ((size_t)i << (logn - s0)) + j;
    uint64_t                 x[n];
#pragma unroll
    for(int k = 0; k < n; k++) x[k] = base[(size_t)k << es_log];
    radix_network<R, FWD, false>(x, p, s0, i);
#pragma unroll
    for(int k = 0; k < n; k++) {
      uint64_t v = x[k];
      if(OUT == 1) v = finish<false>(v, p);
      if(OUT == 2) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
        v = reduce_2q(v, rc);
      }
      base[(size_t)k << es_log] = v;
    }
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* table generation                                                                                  */
/* ------------------------------------------------------------------------------------------------ */

typedef unsigned __int128 u128;

__device__ __forceinline__ uint64_t mulmod128(uint64_t a, uint64_t b, uint64_t q)
{
  return (uint64_t)(((u128)a * b) % q);
}

/* d_w[bitrev_m(i)] = root^i mod q  (calc_w, pre_compute.h:38-51) -- one modpow per entry */
__global__ void k_gen_roots(uint64_t *__restrict__ d_w, uint64_t root, uint32_t logn, uint64_t q)
{
  const uint64_t n = 1ull << logn;
  for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = 1 % q, b = root % q, e = i;
    while(e) {
      if(e & 1) r = mulmod128(r, b, q);
      b = mulmod128(b, b, q);
      e >>= 1;
    }
    const uint64_t rev = logn ? (__brevll(i) >> (64 - logn)) : 0;
    d_w[rev]           = r;
  }
}

/* reference-format w[i] -> device multiplier (wu, qq) and, optionally, floor(w*2^64/q) (calc_w_con) */
__global__ void k_build_tables(const uint64_t *__restrict__ d_w, uint4 *__restrict__ wu, uint2 *__restrict__ qq,
                               uint64_t *__restrict__ con, uint64_t n, uint64_t q, int lazy)
{
  for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t w = d_w[i] % q;
    const uint64_t c = (uint64_t)((((u128)w) << 64) / q);
    if(con) con[i] = c;
    if(lazy) {
      const uint64_t u = (uint64_t)((((u128)w) << 32) % q);
      wu[i]            = make_uint4((uint32_t)w, (uint32_t)(w >> 32), (uint32_t)u, (uint32_t)(u >> 32));
      qq[i]            = make_uint2((uint32_t)((((u128)w) << 30) / q), (uint32_t)((((u128)u) << 30) / q));
    } else {
      wu[i] = make_uint4((uint32_t)w, (uint32_t)(w >> 32), (uint32_t)c, (uint32_t)(c >> 32));
      qq[i] = make_uint2(0u, 0u);
    }
  }
}

/* c = a .* b mod q, exact for any q < 2^62 and inputs < 2^64 (128-bit product, Barrett by 2^128/q) */
/* no __restrict__: callers alias c with a and/or b (in-place products, squaring).  b_mask: all ones, or N-1 to
 * broadcast ONE polynomial b over the batch. */
__global__ void k_pointwise(uint64_t *c, const uint64_t *a, const uint64_t *b, size_t n, size_t b_mask, uint64_t q,
                            uint64_t mu_hi, uint64_t mu_lo)
{
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint64_t x = a[i], y = b[i & b_mask];
    const uint64_t ph = mulhi64(x, y), pl = x * y;
    /* Q = floor(P * mu / 2^128) with mu = floor(2^128 / q) = mu_hi*2^64 + mu_lo; error <= 2 */
    const uint64_t t1 = mulhi64(pl, mu_hi);
    const uint64_t t2 = mulhi64(ph, mu_lo);
    const uint64_t m  = ph * mu_hi; /* low 64 bits of ph*mu_hi: P < q*2^64 keeps Q below 2^64 */
    uint64_t       Q  = m + t1 + t2;
    uint64_t       r  = pl - Q * q;
    r                 = csub(r, q << 1);
    r                 = csub(r, q);
    /* cross-term carries can leave one more q */
    r    = csub(r, q);
    c[i] = r;
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* launchers                                                                                         */
/* ------------------------------------------------------------------------------------------------ */

struct DevGuard {
  int prev = -1;
  bool ok  = true;
  explicit DevGuard(int dev)
  {
    if(cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if(prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DevGuard()
  {
    if(prev >= 0) cudaSetDevice(prev);
  }
};

static int sm_count(int device)
{
  static int cache[64];
  if(device < 0 || device >= 64) return 148;
  if(!cache[device]) {
    int n = 0;
    if(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    cache[device] = n;
  }
  return cache[device];
}

namespace nttb200 {
int nl_sm_count(int device) { return sm_count(device); }
}  // namespace nttb200

extern "C" int ntt_cuda_malloc(int device, void **d_ptr, size_t bytes)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  CU(cudaMalloc(d_ptr, bytes ? bytes : 1));
  return 0;
}
extern "C" int ntt_cuda_free(int device, void *d_ptr)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  CU(cudaFree(d_ptr));
  return 0;
}
extern "C" int ntt_cuda_host_alloc(void **h_ptr, size_t bytes)
{
  CU(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
extern "C" int ntt_cuda_host_free(void *h_ptr)
{
  CU(cudaFreeHost(h_ptr));
  return 0;
}
extern "C" int ntt_cuda_h2d(int device, void *d_dst, const void *h_src, size_t bytes, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
extern "C" int ntt_cuda_d2h(int device, void *h_dst, const void *d_src, size_t bytes, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
extern "C" int ntt_cuda_d2d(int device, void *d_dst, const void *d_src, size_t bytes, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  CU(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
extern "C" int ntt_cuda_sync(int device, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(stream) {
    CU(cudaStreamSynchronize((cudaStream_t)stream));
  } else {
    CU(cudaDeviceSynchronize());
  }
  return 0;
}
extern "C" int ntt_cuda_stream_create(int device, void **stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  cudaStream_t s;
  CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (void *)s;
  return 0;
}
extern "C" int ntt_cuda_stream_destroy(int device, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  CU(cudaStreamDestroy((cudaStream_t)stream));
  return 0;
}

extern "C" int ntt_cuda_gen_root_table(int device, uint64_t *d_w, uint64_t root, uint64_t N, uint64_t q,
                                       void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  uint32_t logn = 0;
  while((1ull << logn) < N) logn++;
  const int blocks = (int)((N + 255) / 256 < 4096 ? (N + 255) / 256 : 4096);
  k_gen_roots<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_w, root, logn, q);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int ntt_cuda_build_tables(int device, const ntt_cuda_params_t *p, const uint64_t *d_w, void *d_wu,
                                     void *d_qq, uint64_t *d_con_out, uint64_t N, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  const int blocks = (int)((N + 255) / 256 < 4096 ? (N + 255) / 256 : 4096);
  k_build_tables<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_w, (uint4 *)d_wu, (uint2 *)d_qq, d_con_out, N, p->q,
                                                          (int)p->lazy);
  CU(cudaGetLastError());
  return 0;
}

/* ---- FP64 twiddles ----------------------------------------------------------------------------------- */

__global__ void k_build_fd(const uint64_t *__restrict__ d_w, double2 *__restrict__ fd, double2 *__restrict__ ct,
                           uint32_t logn, uint64_t q)
{
  /* multipliers are stored CENTRED, w in (-q/2, q/2): |w/q| <= 1/2 doubles the operand range of the quotient
   * estimate and halves its error term (ntt_ring_fp.cuh; the residue class is what matters, not the representative) */
  const size_t n  = (size_t)1 << logn;
  const double qd = (double)q;
  auto centred    = [&](uint64_t w) { return w > (q >> 1) ? -(double)(q - w) : (double)w; }; /* exact: q < 2^50 */
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double w = centred(d_w[i] % q);
    fd[i]          = make_double2(w, __ddiv_rn(w, qd));
  }
  if(ct) {
    const size_t G = n >> 4;
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 15 * G; i += (size_t)gridDim.x * blockDim.x) {
      const uint32_t t = (uint32_t)(i / G);
      const size_t   g = i % G;
      const uint32_t u = 31u - __clz(t + 1u), sub = t + 1u - (1u << u);
      const double   w = centred(d_w[((size_t)1 << (logn - 4 + u)) + (g << u) + sub] % q);
      ct[i]            = make_double2(w, __ddiv_rn(w, qd));
    }
  }
}

extern "C" int ntt_cuda_build_fd_tables(int device, const ntt_cuda_params_t *p, const uint64_t *d_w, void *d_fd,
                                        void *d_ct_fd, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(d_ct_fd && p->logn < 4) return fail_msg("pass-C tables need N >= 16");
  const size_t n      = (size_t)1 << p->logn;
  const int    blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  k_build_fd<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_w, (double2 *)d_fd, (double2 *)d_ct_fd, p->logn, p->q);
  CU(cudaGetLastError());
  return 0;
}

/* ---- pass-C tables ---------------------------------------------------------------------------------- */

/* ct[t*G + g] = entry 2^(logn-4+u) + g*2^u + sub of the stage tables, t = 2^u-1+sub, G = N/16 */
__global__ void k_build_ctables(const uint4 *__restrict__ wu, const uint2 *__restrict__ qq, uint4 *__restrict__ ct_wu,
                                uint2 *__restrict__ ct_qq, uint32_t logn)
{
  const size_t G = (size_t)1 << (logn - 4);
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 15 * G; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t t = (uint32_t)(i / G);
    const size_t   g = i % G;
    const uint32_t u = 31u - __clz(t + 1u), sub = t + 1u - (1u << u);
    const size_t   src = ((size_t)1 << (logn - 4 + u)) + (g << u) + sub;
    ct_wu[i] = wu[src];
    ct_qq[i] = qq[src];
  }
}

extern "C" int ntt_cuda_build_ctables(int device, const ntt_cuda_params_t *p, const void *d_wu, const void *d_qq,
                                      void *d_ct_wu, void *d_ct_qq, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(p->logn < 4) return fail_msg("pass-C tables need N >= 16");
  const size_t total  = (size_t)15 << (p->logn - 4);
  const int    blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  k_build_ctables<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4 *)d_wu, (const uint2 *)d_qq, (uint4 *)d_ct_wu,
                                                           (uint2 *)d_ct_qq, p->logn);
  CU(cudaGetLastError());
  return 0;
}

/* ---- ring kernel launch ------------------------------------------------------------------------------ */

typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tmap_encode_fn tmap_encoder()
{
  static tmap_encode_fn fn = nullptr;
  if(!fn) {
    void *                          sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) == cudaSuccess &&
       qr == cudaDriverEntryPointSuccess)
      fn = (tmap_encode_fn)sym;
  }
  return fn;
}

/* The coefficient array seen as rows of 128 bytes (32 x u32); one TMA box = `rows` rows (32 rows = one
 * 512-coefficient block), written to shared memory with the 128-byte swizzle the passes are laid out for.
 * cuTensorMapEncodeTiled costs about 10 us per call, which shows when a config launches many small batches (48 RNS
 * limbs of 32 polynomials: 96 encodes per transform), so the last few descriptors are kept per thread, keyed on
 * (pointer, words, rows) -- the descriptor depends on nothing else. */
static int make_block_tmap(CUtensorMap *tm, uint64_t *d_a, size_t total_words, unsigned rows = 32)
{
  struct Entry {
    uint64_t *  ptr;
    size_t      words;
    unsigned    rows;
    CUtensorMap tm;
  };
  constexpr int                NCACHE = 128;
  static thread_local Entry    cache[NCACHE];
  static thread_local unsigned next = 0, used = 0;
  for(unsigned i = 0; i < used; i++) {
    if(cache[i].ptr == d_a && cache[i].words == total_words && cache[i].rows == rows) {
      *tm = cache[i].tm;
      return 0;
    }
  }
  tmap_encode_fn enc = tmap_encoder();
  if(!enc) return fail_msg("cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2]    = {32, (cuuint64_t)(total_words / 16)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2]     = {32, rows};
  const cuuint32_t estr[2]    = {1, 1};
  const CUresult   r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_a, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if(r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -1;
  }
  cache[next] = Entry{d_a, total_words, rows, *tm};
  next        = (next + 1) % NCACHE;
  if(used < NCACHE) used++;
  return 0;
}
namespace nttb200 {
int nl_make_block_tmap(CUtensorMap *tm, uint64_t *d_a, size_t total_words, unsigned rows)
{
  return make_block_tmap(tm, d_a, total_words, rows);
}
}  // namespace nttb200

/* kernel-selection switches (benchmarks and A/B parity tests): environment at first use, or ntt_cuda_configure */
static int g_ring_on = -1, g_fp64_on = -1, g_polymul_on = -1;
static bool ring_enabled()
{
  if(g_ring_on < 0) {
    const char *e = getenv("NTT_B200_NO_RING");
    g_ring_on     = (e && e[0] == '1') ? 0 : 1;
  }
  return g_ring_on == 1;
}
extern "C" int ntt_cuda_configure(const char *key, int value)
{
  if(key && !strcmp(key, "ring")) { g_ring_on = value ? 1 : 0; return 0; }
  if(key && !strcmp(key, "fp64")) { g_fp64_on = value ? 1 : 0; return 0; }
  if(key && !strcmp(key, "polymul")) { g_polymul_on = value ? 1 : 0; return 0; }
  return fail_msg("unknown configuration key");
}

static bool fp64_enabled()
{
  if(g_fp64_on < 0) {
    const char *e = getenv("NTT_B200_NO_FP64");
    g_fp64_on     = (e && e[0] == '1') ? 0 : 1;
  }
  return g_fp64_on == 1;
}
/* does this (plan, direction) run its chunk stage on the FP64 ring kernel? */
static bool use_fp64(const ntt_cuda_params_t &p, bool fwd)
{
  return fp64_enabled() && p.fp64 && (fwd ? p.fwd_ct_fd : p.inv_ct_fd) != nullptr;
}

/* Does the chunk stage of this transform run on the FP64 ring kernel (same conditions as try_ring)?  Then the strided
 * passes around it run in FP64 too (ntt_strided_fp.cuh): they exchange canonical residues with it.
 * NTT_B200_NO_FP_STRIDED=1 keeps the integer passes (A/B timing). */
static bool fp_strided_path(const ntt_cuda_params_t &p, int L, const uint64_t *d_a, bool fwd)
{
  static int on = -1;
  if(on < 0) {
    const char *e = getenv("NTT_B200_NO_FP_STRIDED");
    on            = (e && e[0] == '1') ? 0 : 1;
  }
  return on && ring_enabled() && p.lazy && L >= 12 && L <= 14 && use_fp64(p, fwd) && (fwd ? p.fwd_ct_wu : p.inv_ct_wu) &&
         (fwd ? p.fwd_fd : p.inv_fd) && ((uintptr_t)d_a & 127) == 0;
}

/* Several small launches run side by side (RNS limbs on their own streams): each then takes only as many CTAs
 * as gives every CTA a few chunks to pipeline, and leaves the other SMs to its neighbours.  0 = whole GPU. */
static thread_local size_t g_min_chunks_per_cta = 0;
namespace nttb200 {
size_t nl_min_chunks_per_cta() { return g_min_chunks_per_cta; }
}  // namespace nttb200

/* true if the ring kernel handled the chunk stage (lazy path, chunk of 2^10..2^14 -- 2^10 and 2^11 on the FP64 kernel only --, pass-C tables present).
 * `o` (forward only): fused pointwise product / lazy output, honoured by the FP64 kernel only -- if they are asked
 * for and the FP64 kernel cannot run, nothing is launched and *done stays false. */
template <bool FWD>
static int try_ring(int device, int L, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks, cudaStream_t st,
                    bool *done, const RingOpts *o = nullptr)
{
  *done = false;
  if(!ring_enabled() || !p.lazy || L < 10 || L > 14) return 0;
  if(L <= 11 && !use_fp64(p, FWD)) return 0; /* chunks of 2^10 / 2^11: FP64 ring kernel only */
  if(!(FWD ? p.fwd_ct_wu : p.inv_ct_wu)) return 0;
  if(((uintptr_t)d_a & 127) != 0) return 0; /* TMA wants 128-byte aligned rows (cudaMalloc gives 256) */
  if(o && o->d_other && !use_fp64(p, FWD)) return 0; /* the fused product exists on the FP64 kernel only */
  *done = true;
  if(use_fp64(p, FWD)) {
    const RingOpts none;
    const RingOpts &ro = o ? *o : none;
    switch(L) {
      case 10: return ring_fp_launch_10(FWD, device, p, d_a, n_chunks, st, ro);
      case 11: return ring_fp_launch_11(FWD, device, p, d_a, n_chunks, st, ro);
      case 12: return ring_fp_launch_12(FWD, device, p, d_a, n_chunks, st, ro);
      case 13: return ring_fp_launch_13(FWD, device, p, d_a, n_chunks, st, ro);
      default: return ring_fp_launch_14(FWD, device, p, d_a, n_chunks, st, ro);
    }
  }
  return ring_int_launch(L, FWD, device, p, d_a, n_chunks, st);
}

/* ---- transform dispatch ---------------------------------------------------------------------------- */

template <int L, bool FWD, bool EXACT>
static int launch_chunk(int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks, cudaStream_t st)
{
  using C            = ChunkCfg<L>;
  const size_t smem  = (size_t)8 << L;
  auto         kern  = k_chunk<L, FWD, EXACT>;
  static bool  ready[64] = {false};
  if(!ready[device & 63]) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ready[device & 63] = true;
  }
  int per_sm = 1;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS, smem));
  if(per_sm < 1) per_sm = 1;
  size_t grid = (size_t)sm_count(device) * per_sm;
  if(grid > n_chunks) grid = n_chunks;
  kern<<<(unsigned)grid, C::THREADS, smem, st>>>(p, d_a, n_chunks);
  CU(cudaGetLastError());
  return 0;
}

template <bool FWD, bool EXACT>
static int dispatch_chunk(int device, int L, const ntt_cuda_params_t &p, uint64_t *d_a, size_t n_chunks,
                          cudaStream_t st)
{
  switch(L) {
#define CASE(l) \
  case l: return launch_chunk<l, FWD, EXACT>(device, p, d_a, n_chunks, st);
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
    CASE(14)
#undef CASE
    default: return fail_msg("unsupported chunk size");
  }
}

template <int R, bool FWD, bool EXACT, int OUT>
static int launch_strided(int device, const ntt_cuda_params_t &p, uint64_t *d_a, uint32_t s0, size_t first_group,
                          size_t n_groups, cudaStream_t st)
{
  const size_t cap = (size_t)sm_count(device) * 32;
  if constexpr(R <= 4) {
    /* measured (profiles/r02_kernel_experiments.txt): a gain at radix 4 (N = 2^16: 0.540 -> 0.503 ms per 1024),
     * a loss at radix 16 (registers); NTT_B200_STRIDED_V2_MAXR overrides the largest radix exponent it is used for */
    static int v2_maxr = -1;
    if(v2_maxr < 0) {
      const char *e = getenv("NTT_B200_STRIDED_V2_MAXR");
      v2_maxr       = e ? atoi(e) : 2;
    }
    const uint32_t es_log = p.logn - s0 - R;
    if(R <= v2_maxr && es_log >= 1 && (first_group & 1) == 0 && (n_groups & 1) == 0 && ((uintptr_t)d_a & 15) == 0) {
      size_t grid = (n_groups / 2 + 255) / 256;
      if(grid > cap) grid = cap;
      k_strided_v2<R, FWD, EXACT, OUT><<<(unsigned)grid, 256, 0, st>>>(p, d_a, s0, first_group, n_groups);
      CU(cudaGetLastError());
      return 0;
    }
  }
  size_t grid = (n_groups + 255) / 256;
  if(grid > cap) grid = cap;
  k_strided<R, FWD, EXACT, OUT><<<(unsigned)grid, 256, 0, st>>>(p, d_a, s0, first_group, n_groups);
  CU(cudaGetLastError());
  return 0;
}

template <bool FWD, bool EXACT, int OUT>
static int dispatch_strided_groups(int device, int R, const ntt_cuda_params_t &p, uint64_t *d_a, uint32_t s0,
                                   size_t first_group, size_t n_groups, cudaStream_t st)
{
  switch(R) {
    case 1: return launch_strided<1, FWD, EXACT, OUT>(device, p, d_a, s0, first_group, n_groups, st);
    case 2: return launch_strided<2, FWD, EXACT, OUT>(device, p, d_a, s0, first_group, n_groups, st);
    case 3: return launch_strided<3, FWD, EXACT, OUT>(device, p, d_a, s0, first_group, n_groups, st);
    case 4: return launch_strided<4, FWD, EXACT, OUT>(device, p, d_a, s0, first_group, n_groups, st);
    case 5: return launch_strided<5, FWD, EXACT, OUT>(device, p, d_a, s0, first_group, n_groups, st);
    default: return fail_msg("unsupported strided radix");
  }
}

template <bool FWD, bool EXACT, int OUT>
static int dispatch_strided(int device, int R, const ntt_cuda_params_t &p, uint64_t *d_a, uint32_t s0, size_t batch,
                            cudaStream_t st)
{
  return dispatch_strided_groups<FWD, EXACT, OUT>(device, R, p, d_a, s0, 0, batch << (p.logn - R), st);
}

/* k_strided_multi with two adjacent groups per thread (see k_strided_v2) */
/* (R <= 2: capped at 64 registers -- four resident CTAs; at 70 registers the forward pass ran at 5.0 instead of 5.9 TB/s) */
template <int R, bool FWD, int OUT>
__global__ void __launch_bounds__(256, R <= 2 ? 4 : 1) k_strided_multi_v2(const __grid_constant__ RingLimbs<true> limbs,
                                                          uint64_t *__restrict__ a, uint32_t s0, size_t n_groups)
{
  constexpr int  n      = 1 << R;
  const uint32_t logn   = limbs.e[0].logn;
  const uint32_t es_log = logn - s0 - R;
  const uint32_t gl     = logn - R;
  for(size_t t2 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t2 < n_groups / 2; t2 += (size_t)gridDim.x * blockDim.x) {
    const size_t             t    = 2 * t2;
    const size_t             poly = t >> gl;
    const ntt_cuda_params_t &p    = limbs.e[(uint32_t)poly / limbs.polys_per_limb];
    const uint32_t           g    = (uint32_t)(t & ((1u << gl) - 1u));
    const uint32_t           i    = g >> es_log;
    const uint32_t           j    = g & ((1u << es_log) - 1u);
    uint64_t *               base = a + (poly << logn) + ((size_t)i << (logn - s0)) + j;
    uint64_t                 x0[n], x1[n];
#pragma unroll
    for(int k = 0; k < n; k++) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(base + ((size_t)k << es_log));
      x0[k] = v.x;
      x1[k] = v.y;
    }
    radix_network<R, FWD, false>(x0, p, s0, i);
    radix_network<R, FWD, false>(x1, p, s0, i);
#pragma unroll
    for(int k = 0; k < n; k++) {
      ulonglong2 v = make_ulonglong2(x0[k], x1[k]);
      if(OUT == 1) {
        v.x = finish<false>(v.x, p);
        v.y = finish<false>(v.y, p);
      }
      if(OUT == 2) {
        const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
        v.x = reduce_2q(v.x, rc);
        v.y = reduce_2q(v.y, rc);
      }
      *reinterpret_cast<ulonglong2 *>(base + ((size_t)k << es_log)) = v;
    }
  }
}

template <int R, bool FWD, int OUT>
static int launch_strided_multi(int device, const RingLimbs<true> &lb, uint64_t *d_a, uint32_t s0, size_t total_polys,
                                cudaStream_t st)
{
  const size_t n_groups = total_polys << (lb.e[0].logn - R);
  const size_t cap      = (size_t)sm_count(device) * 32;
  if constexpr(R <= 2) {
    if(lb.e[0].logn - s0 - R >= 1 && ((uintptr_t)d_a & 15) == 0) {
      size_t grid = (n_groups / 2 + 255) / 256;
      if(grid > cap) grid = cap;
      k_strided_multi_v2<R, FWD, OUT><<<(unsigned)grid, 256, 0, st>>>(lb, d_a, s0, n_groups);
      CU(cudaGetLastError());
      return 0;
    }
  }
  size_t grid = (n_groups + 255) / 256;
  if(grid > cap) grid = cap;
  k_strided_multi<R, FWD, OUT><<<(unsigned)grid, 256, 0, st>>>(lb, d_a, s0, n_groups);
  CU(cudaGetLastError());
  return 0;
}
template <bool FWD, int OUT>
static int dispatch_strided_multi(int device, int R, const RingLimbs<true> &lb, uint64_t *d_a, uint32_t s0,
                                  size_t total_polys, cudaStream_t st)
{
  switch(R) {
    case 1: return launch_strided_multi<1, FWD, OUT>(device, lb, d_a, s0, total_polys, st);
    case 2: return launch_strided_multi<2, FWD, OUT>(device, lb, d_a, s0, total_polys, st);
    case 3: return launch_strided_multi<3, FWD, OUT>(device, lb, d_a, s0, total_polys, st);
    case 4: return launch_strided_multi<4, FWD, OUT>(device, lb, d_a, s0, total_polys, st);
    case 5: return launch_strided_multi<5, FWD, OUT>(device, lb, d_a, s0, total_polys, st);
    default: return fail_msg("unsupported strided radix");
  }
}

/* How the stages of a 2^logn transform are split: `ns` strided passes of radix 2^r[i] cover the first
 * S1 = sum r[i] stages, the chunk kernel covers the remaining L = logn - S1 (<= 14). */
struct Split {
  int L;
  int ns;
  int r[4];
};
static Split make_split(int logn)
{
  /* as few stages as possible go to the strided passes: measured at N = 2^16, (2 strided + 2^14 chunks) takes
   * 0.58 ms per 1024 transforms against 0.68 ms for (4 strided + 2^12 chunks) */
  Split s{};
  int   rest = logn > 14 ? logn - 14 : 0; /* stages that do not fit a chunk */
  s.L        = logn - rest;
  s.ns       = 0;
  while(rest > 0) {
    /* balanced radices, at most 2^5 per pass */
    const int passes = (rest + 4) / 5;
    const int r      = (rest + passes - 1) / passes;
    s.r[s.ns++]      = r;
    rest -= r;
  }
  return s;
}

/* Inverse lazy bookkeeping (see ntt_cuda.h): walks the passes in the order the inverse runs them and
 * records, per global stage s, the bound constant inv_c[s] = B_s*q and where values must first be pulled
 * back below 2q so that neither x+y nor x-y+B_s*q can reach 2^63. */
extern "C" int ntt_cuda_plan_inverse_bounds(ntt_cuda_params_t *p)
{
  const int logn = (int)p->logn;
  if(logn < 1 || logn > NTT_MAX_STAGES) return fail_msg("logn out of range");
  memset(p->inv_c, 0, sizeof(p->inv_c));
  p->inv_renorm_mask = 0;
  const Split sp     = make_split(logn);
  int         s1     = 0;
  for(int k = 0; k < sp.ns; k++) s1 += sp.r[k];
  /* (top stage, radix) of every pass in inverse processing order */
  int tops[8], rads[8], np = 0;
  {
    const int L  = sp.L;
    const int RC = L < 4 ? L : 4, RB = (L - RC) < 5 ? (L - RC) : 5, RA = L - RC - RB;
    if(RC) { tops[np] = s1 + L - 1; rads[np++] = RC; }
    if(RB) { tops[np] = s1 + RA + RB - 1; rads[np++] = RB; }
    if(RA) { tops[np] = s1 + RA - 1; rads[np++] = RA; }
    int s0 = s1;
    for(int k = sp.ns - 1; k >= 0; k--) {
      s0 -= sp.r[k];
      tops[np]   = s0 + sp.r[k] - 1;
      rads[np++] = sp.r[k];
    }
  }
  /* values must stay below 2^63 (x+y, x-y+B*q) and below 2^(bitlen(q)+22) (precondition of reduce_2q) */
  long double lim = 9223372036854775808.0L / (long double)p->q; /* in units of q */
  if(lim > 2097152.0L) lim = 2097152.0L;                         /* 2^21 */
  long double B = 2.0L;                                          /* input contract of the inverse: [0,2q) */
  for(int k = 0; k < np; k++) {
    if(B * (long double)(1u << rads[k]) >= lim) {
      p->inv_renorm_mask |= 1u << tops[k];
      B = 2.0L;
    }
    if(B * (long double)(1u << rads[k]) >= lim) return fail_msg("q too large for the lazy inverse");
    for(int u = 0; u < rads[k]; u++) {
      const int s = tops[k] - u;
      p->inv_c[s] = (uint64_t)B * p->q;
      B           = (2 * B > 10.0L) ? 2 * B : 10.0L;
    }
  }
  return 0;
}

/* d_other != nullptr: also multiply pointwise by d_other (fused into the chunk kernel); *fused tells the caller
 * whether that happened (it does on the FP64 ring kernel), otherwise nothing was multiplied. */
/* phases (RNS batches issue all strided passes before any chunk kernel): bit 0 = strided passes, bit 1 = chunk kernel */
enum { PH_STRIDED = 1, PH_CHUNK = 2, PH_ALL = 3 };
template <bool EXACT>
static int forward_impl(int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t batch, cudaStream_t st,
                        const ntt_cuda_fwd_opts_t *opts = nullptr, bool *fused = nullptr, int phases = PH_ALL)
{
  if(fused) *fused = false;
  const Split sp = make_split((int)p.logn);
  uint32_t    s0 = 0;
  /* the FP64 chunk kernel wants inputs below 2^52: the last strided pass then hands over values below 2q */
  const bool fp_next = !EXACT && ring_enabled() && p.lazy && sp.L >= 12 && use_fp64(p, true) && ((uintptr_t)d_a & 127) == 0;
  const bool fp_pass = !EXACT && fp_strided_path(p, sp.L, d_a, true);
  for(int k = 0; k < sp.ns; k++) {
    if(phases & PH_STRIDED) {
      const int rc = fp_pass ? strided_fp_launch(true, false, sp.r[k], device, p, d_a, s0, batch, st)
                             : ((fp_next && k == sp.ns - 1) ? dispatch_strided<true, EXACT, 2>(device, sp.r[k], p, d_a, s0, batch, st)
                                                            : dispatch_strided<true, EXACT, 0>(device, sp.r[k], p, d_a, s0, batch, st));
      if(rc) return -1;
    }
    s0 += sp.r[k];
  }
  if(!(phases & PH_CHUNK)) return 0;
  if(!EXACT) {
    bool done = false;
    if(opts && (opts->d_other || opts->lazy_out)) {
      RingOpts ro;
      ro.d_other    = opts->d_other;
      ro.other_mask = opts->other_broadcast ? (((size_t)1 << s0) - 1) : ~(size_t)0;
      ro.lazy_out   = opts->lazy_out != 0 && !opts->d_other;
      if(try_ring<true>(device, sp.L, p, d_a, batch << s0, st, &done, &ro)) return -1;
      if(done) {
        if(fused) *fused = true;
        return 0;
      }
    }
    if(try_ring<true>(device, sp.L, p, d_a, batch << s0, st, &done)) return -1;
    if(done) return 0;
  }
  return dispatch_chunk<true, EXACT>(device, sp.L, p, d_a, batch << s0, st);
}

template <bool EXACT>
static int inverse_impl(int device, const ntt_cuda_params_t &p, uint64_t *d_a, size_t batch, cudaStream_t st,
                        int phases = PH_ALL)
{
  const Split sp = make_split((int)p.logn);
  uint32_t    s1 = 0;
  for(int k = 0; k < sp.ns; k++) s1 += sp.r[k];
  if(phases & PH_CHUNK) {
    bool done = false;
    if(!EXACT && try_ring<false>(device, sp.L, p, d_a, batch << s1, st, &done)) return -1;
    if(!done && dispatch_chunk<false, EXACT>(device, sp.L, p, d_a, batch << s1, st)) return -1;
  }
  if(!(phases & PH_STRIDED)) return 0;
  const bool fp_pass = !EXACT && fp_strided_path(p, sp.L, d_a, false);
  uint32_t   s0      = s1;
  for(int k = sp.ns - 1; k >= 0; k--) {
    s0 -= sp.r[k];
    const int rc = fp_pass ? strided_fp_launch(false, k == 0, sp.r[k], device, p, d_a, s0, batch, st)
                           : ((k == 0) ? dispatch_strided<false, EXACT, 1>(device, sp.r[k], p, d_a, s0, batch, st)
                                       : dispatch_strided<false, EXACT, 0>(device, sp.r[k], p, d_a, s0, batch, st));
    if(rc) return -1;
  }
  return 0;
}

/* Which kernels a transform of this plan launches (what the dispatch above would do for 128-byte aligned data):
 * buf receives e.g. "k_strided<2> + k_ring_fp<14,fwd>", *launches their number.  For benchmarks and reports. */
extern "C" int ntt_cuda_describe(const ntt_cuda_params_t *p, int inverse, char *buf, size_t n, int *launches)
{
  if(!p || !buf || n == 0) return fail_msg("describe: NULL argument");
  const Split sp  = make_split((int)p->logn);
  const bool  fwd = !inverse;
  char        chunk[64];
  const bool  ring = ring_enabled() && p->lazy && sp.L <= 14 && (sp.L >= 12 || (sp.L >= 10 && use_fp64(*p, fwd))) &&
                    (fwd ? p->fwd_ct_wu : p->inv_ct_wu);
  if(ring && use_fp64(*p, fwd)) snprintf(chunk, sizeof(chunk), "k_ring_fp<%d,%s>", sp.L, fwd ? "fwd" : "inv");
  else if(ring) snprintf(chunk, sizeof(chunk), "k_ring<%d,%s>", sp.L, fwd ? "fwd" : "inv");
  else snprintf(chunk, sizeof(chunk), "k_chunk<%d,%s,%s>", sp.L, fwd ? "fwd" : "inv", p->lazy ? "lazy" : "exact");
  size_t off = 0;
  buf[0]     = 0;
  /* (128-byte aligned data assumed, as for the chunk kernel) */
  const char *sk = fp_strided_path(*p, sp.L, nullptr, fwd) ? "k_strided_fp" : "k_strided";
  if(fwd) {
    for(int k = 0; k < sp.ns && off < n; k++) off += (size_t)snprintf(buf + off, n - off, "%s<%d> + ", sk, sp.r[k]);
    if(off < n) off += (size_t)snprintf(buf + off, n - off, "%s", chunk);
  } else {
    if(off < n) off += (size_t)snprintf(buf + off, n - off, "%s", chunk);
    for(int k = sp.ns - 1; k >= 0 && off < n; k--) off += (size_t)snprintf(buf + off, n - off, " + %s<%d>", sk, sp.r[k]);
  }
  if(launches) *launches = sp.ns + 1;
  return 0;
}

extern "C" int ntt_cuda_forward(int device, const ntt_cuda_params_t *p, uint64_t *d_a, size_t batch, void *stream)
{
  if(batch == 0) return 0;
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(p->logn < 1 || p->logn > NTT_MAX_STAGES) return fail_msg("logn out of range");
  return p->lazy ? forward_impl<false>(device, *p, d_a, batch, (cudaStream_t)stream)
                 : forward_impl<true>(device, *p, d_a, batch, (cudaStream_t)stream);
}

/* forward transform with options (ntt_cuda_fwd_opts_t): a pointwise product fused before the store and / or a lazy
 * output.  *fused_out = 1 if the chunk kernel honoured them; 0 means a plain, fully reduced transform was run and
 * the caller still has to multiply (ntt_cuda_pointwise). */
extern "C" int ntt_cuda_forward_ex(int device, const ntt_cuda_params_t *p, uint64_t *d_a, size_t batch, void *stream,
                                   const ntt_cuda_fwd_opts_t *opts, int *fused_out)
{
  if(fused_out) *fused_out = 0;
  if(batch == 0) return 0;
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(p->logn < 1 || p->logn > NTT_MAX_STAGES) return fail_msg("logn out of range");
  bool      fused = false;
  const int rc    = p->lazy ? forward_impl<false>(device, *p, d_a, batch, (cudaStream_t)stream, opts, &fused)
                            : forward_impl<true>(device, *p, d_a, batch, (cudaStream_t)stream);
  if(fused_out) *fused_out = fused ? 1 : 0;
  return rc;
}

extern "C" int ntt_cuda_forward_mul(int device, const ntt_cuda_params_t *p, uint64_t *d_a, const uint64_t *d_other,
                                    size_t batch, void *stream, int *fused_out)
{
  ntt_cuda_fwd_opts_t o;
  memset(&o, 0, sizeof(o));
  o.d_other = d_other;
  return ntt_cuda_forward_ex(device, p, d_a, batch, stream, &o, fused_out);
}

/* c = a * b in Z_q[X]/(X^N+1), `batch` products, in ONE kernel (both operands in shared memory, nothing of the NTT
 * domain in global memory).  *done = 0 if this plan / these pointers are not served (N != 2^13, modulus outside the
 * FP64 range, ring kernels switched off, a or b not 128-byte aligned): the caller then composes the product from
 * transforms.  c may alias a or b; a == b squares. */
extern "C" int ntt_cuda_polymul(int device, const ntt_cuda_params_t *p, uint64_t *d_c, uint64_t *d_a, uint64_t *d_b,
                                size_t batch, void *stream, int *done)
{
  *done = 0;
  if(batch == 0) return 0;
  if(p->logn != 13 || !ring_enabled() || !p->lazy || !use_fp64(*p, true) || !use_fp64(*p, false)) return 0;
  if((((uintptr_t)d_a | (uintptr_t)d_b) & 127) != 0 || ((uintptr_t)d_c & 15) != 0) return 0;
  if(g_polymul_on < 0) {
    const char *e = getenv("NTT_B200_NO_FUSED_POLYMUL");
    g_polymul_on  = (e && e[0] == '1') ? 0 : 1;
  }
  if(!g_polymul_on) return 0;
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  *done = 1;
  return polymul_fp_launch(device, *p, d_a, d_b, d_c, batch, (cudaStream_t)stream);
}

extern "C" int ntt_cuda_inverse(int device, const ntt_cuda_params_t *p, uint64_t *d_a, size_t batch, void *stream)
{
  if(batch == 0) return 0;
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(p->logn < 1 || p->logn > NTT_MAX_STAGES) return fail_msg("logn out of range");
  return p->lazy ? inverse_impl<false>(device, *p, d_a, batch, (cudaStream_t)stream)
                 : inverse_impl<true>(device, *p, d_a, batch, (cudaStream_t)stream);
}

/*
 * Tail pass of a transform of size N = 2^logn spread over 2^glog devices: the last glog
 * stages (distances 2^(glog-1) .. 1) on the contiguous block `block` (N >> glog words at d_block).  Forward:
 * input in [0,4q), output fully reduced.  Inverse: the same stages backwards (they come first), input in
 * [0,2q), output below 2q for the complete local inverse that follows the exchange.
 */
/* The inverse tail runs stages logn-1 .. logn-glog as ONE register network, whatever pass structure the size-N plan
 * was laid out for, so the plan's inv_c[] / inv_renorm_mask (computed for that structure: a renormalisation may be
 * scheduled at stage logn-5, which this network would skip) do not apply.  Recompute the bounds for the tail's own
 * stages: B = 2 (input contract), then 10, 20, 40, 80; no renormalisation inside. */
static int tail_inverse_params(const ntt_cuda_params_t *p, uint32_t glog, ntt_cuda_params_t *out)
{
  *out = *p;
  if(!p->lazy) return 0;
  long double lim = 9223372036854775808.0L / (long double)p->q;
  if(lim > 2097152.0L) lim = 2097152.0L;
  long double B = 2.0L;
  uint32_t    mask = p->inv_renorm_mask;
  for(uint32_t u = 0; u < glog; u++) {
    const uint32_t s = p->logn - 1 - u;
    if(2 * B >= lim) return fail_msg("q too large for the lazy inverse tail");
    out->inv_c[s] = (uint64_t)B * p->q;
    mask &= ~(1u << s);
    B = (2 * B > 10.0L) ? 2 * B : 10.0L;
  }
  out->inv_renorm_mask = mask;
  return 0;
}

extern "C" int ntt_cuda_tail(int device, const ntt_cuda_params_t *p_in, uint64_t *d_block, uint32_t glog, uint32_t block,
                             int inverse, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(glog < 1 || glog > 5 || 2 * glog > p_in->logn || block >= (1u << glog)) return fail_msg("bad tail geometry");
  ntt_cuda_params_t tp;
  if(inverse) {
    if(tail_inverse_params(p_in, glog, &tp)) return -1;
  }
  const ntt_cuda_params_t *p = inverse ? &tp : p_in;
  const uint32_t s0       = p->logn - glog;
  const size_t   n_groups = (size_t)1 << (p->logn - 2 * glog);       /* groups of 2^glog words in one block */
  const size_t   first    = (size_t)block * n_groups;
  uint64_t *     vbase    = d_block - ((size_t)block << (p->logn - glog)); /* virtual start of the polynomial */
  cudaStream_t   st       = (cudaStream_t)stream;
  if(p->lazy) {
    return inverse ? dispatch_strided_groups<false, false, 2>(device, (int)glog, *p, vbase, s0, first, n_groups, st)
                   : dispatch_strided_groups<true, false, 1>(device, (int)glog, *p, vbase, s0, first, n_groups, st);
  }
  return inverse ? dispatch_strided_groups<false, true, 0>(device, (int)glog, *p, vbase, s0, first, n_groups, st)
                 : dispatch_strided_groups<true, true, 1>(device, (int)glog, *p, vbase, s0, first, n_groups, st);
}

/* ------------------------------------------------------------------------------------------------ */
/* tail stages fused with the exchange over peer memory (NVLink)                                      */
/* ------------------------------------------------------------------------------------------------ */

/*
 * The distributed transform's exchange step without a collective: every rank maps the other ranks' slice
 * buffers (CUDA IPC) and the tail kernel does the all-to-all itself.  Group kk of block `rank` is the G words
 * { slice_p[rank*piece + kk] : p = 0..G-1 } (exchange_cyclic_to_blocks in fourstep.py), so
 *   forward:  x[p] is LOADED from peer p's slice (consecutive threads read consecutive words of the same peer:
 *             256 contiguous bytes per warp and peer), the last g stages run in registers, the fully reduced
 *             group is stored to this rank's block;
 *   inverse:  the group is read from this rank's block, the first g inverse stages run, x[p] is STORED into
 *             peer p's slice.
 * Peer data must bypass L1 (it is written by another GPU between launches): ld/st .relaxed.sys.
 */
struct PeerPtrs {
  uint64_t *p[32];
};
__device__ __forceinline__ uint64_t ld_sys(const uint64_t *a)
{
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(uint64_t *a, uint64_t v)
{
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}

/* batch > 1: every slice buffer holds `batch` slices of N/G words one after the other and d_block `batch` blocks;
 * one launch (and one barrier) then serves the whole batch -- at N = 2^22 the exchange of a single polynomial is
 * latency-bound, a batch amortises the launches and the barrier. */
template <int R, bool FWD, bool EXACT>
__global__ void __launch_bounds__(256) k_tail_peer(const __grid_constant__ ntt_cuda_params_t p,
                                                   const __grid_constant__ PeerPtrs peers, uint64_t *__restrict__ block,
                                                   uint32_t rank, uint32_t piece_log, size_t total)
{
  constexpr int  n       = 1 << R;
  const uint32_t s0      = p.logn - R;
  const size_t   piece   = (size_t)1 << piece_log;
  const uint32_t loc_log = p.logn - R; /* log2 words of one slice / one block */
  for(size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t b = t >> piece_log, kk = t & (piece - 1);
    const size_t grp = (size_t)rank * piece + kk;        /* global group index of this block = index in every slice */
    const size_t src = (b << loc_log) + grp;
    uint64_t *   dst = block + (b << loc_log) + (kk << R);
    uint64_t     x[n];
    if(FWD) {
#pragma unroll
      for(int k = 0; k < n; k++) x[k] = ld_sys(peers.p[k] + src);
    } else {
#pragma unroll
      for(int k = 0; k < n; k++) x[k] = dst[k];
    }
    radix_network<R, FWD, EXACT>(x, p, s0, (uint32_t)grp);
#pragma unroll
    for(int k = 0; k < n; k++) {
      uint64_t v = x[k];
      if(FWD) {
        dst[k] = finish<EXACT>(v, p);
      } else {
        if(!EXACT) {
          const Red rc{p.q, p.negq, p.red_shift, p.red_mu};
          v = reduce_2q(v, rc);
        }
        st_sys(peers.p[k] + src, v);
      }
    }
  }
}

template <int R, bool FWD, bool EXACT>
static int launch_tail_peer(int device, const ntt_cuda_params_t &p, const PeerPtrs &pp, uint64_t *d_block, uint32_t rank,
                            uint32_t piece_log, size_t batch, cudaStream_t st)
{
  const size_t total = batch << piece_log;
  size_t       grid  = (total + 255) / 256;
  const size_t cap   = (size_t)sm_count(device) * 8;
  if(grid > cap) grid = cap;
  k_tail_peer<R, FWD, EXACT><<<(unsigned)grid, 256, 0, st>>>(p, pp, d_block, rank, piece_log, total);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int ntt_cuda_tail_peer(int device, const ntt_cuda_params_t *p_in, uint64_t *const *peer_slices,
                                  uint64_t *d_block, uint32_t glog, uint32_t rank, size_t batch, int inverse,
                                  void *stream)
{
  if(batch == 0) return 0;
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(glog < 1 || glog > 5 || 2 * glog > p_in->logn || rank >= (1u << glog)) return fail_msg("bad tail geometry");
  ntt_cuda_params_t tp;
  if(inverse) {
    if(tail_inverse_params(p_in, glog, &tp)) return -1;
  }
  const ntt_cuda_params_t *p = inverse ? &tp : p_in;
  PeerPtrs pp{};
  for(uint32_t k = 0; k < (1u << glog); k++) {
    if(!peer_slices[k]) return fail_msg("peer slice pointer is NULL");
    pp.p[k] = peer_slices[k];
  }
  const uint32_t piece = p->logn - 2 * glog; /* log2 groups per block */
  cudaStream_t   st    = (cudaStream_t)stream;
#define TAILPEER(R)                                                                                                   \
  case R:                                                                                                             \
    if(p->lazy)                                                                                                       \
      return inverse ? launch_tail_peer<R, false, false>(device, *p, pp, d_block, rank, piece, batch, st)             \
                     : launch_tail_peer<R, true, false>(device, *p, pp, d_block, rank, piece, batch, st);             \
    return inverse ? launch_tail_peer<R, false, true>(device, *p, pp, d_block, rank, piece, batch, st)                \
                   : launch_tail_peer<R, true, true>(device, *p, pp, d_block, rank, piece, batch, st);
  switch(glog) {
    TAILPEER(1)
    TAILPEER(2)
    TAILPEER(3)
    TAILPEER(4)
    TAILPEER(5)
  }
#undef TAILPEER
  return fail_msg("unsupported tail radix");
}

/*
 * Barrier between the ranks of a distributed transform, on the GPU timeline: thread k stores `epoch` into slot
 * `rank` of peer k's flag array (after a system-scope fence, so everything earlier kernels of this stream wrote is
 * visible to the peer before the flag is), then waits until slot k of its own array has reached `epoch`.  Flags
 * only grow; the array is 128 words: flags, word 64 = timeout report, word 65 = device-side epoch counter.  The wait is bounded (about two seconds of clock64) and reports a timeout instead of hanging the GPU.
 */
__global__ void k_peer_barrier(const __grid_constant__ PeerPtrs flags, uint32_t rank, uint32_t world, uint32_t epoch,
                               uint32_t *my_flags, int *timed_out)
{
  const uint32_t k = threadIdx.x;
  /* epoch == 0: count the calls on the device (word 65 of the flag array), so that a captured CUDA graph can be
   * replayed -- every rank runs the same sequence of barriers, the counters agree */
  __shared__ uint32_t s_epoch;
  if(k == 0) s_epoch = epoch ? epoch : ++my_flags[65];
  __syncthreads();
  epoch = s_epoch;
  if(k >= world) return;
  __threadfence_system();
  uint32_t *remote = reinterpret_cast<uint32_t *>(flags.p[k]) + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const long long t0 = clock64();
  uint32_t        v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + k) : "memory");
    if((int)(v - epoch) >= 0) break;
    __nanosleep(64);
  } while(clock64() - t0 < 4000000000ll);
  if((int)(v - epoch) < 0) *timed_out = 1;
  __threadfence_system();
}

extern "C" int ntt_cuda_peer_barrier(int device, void *const *peer_flags, void *my_flags, uint32_t rank, uint32_t world,
                                     uint32_t epoch, int *d_timed_out, void *stream)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(world < 1 || world > 32 || rank >= world) return fail_msg("bad peer barrier geometry");
  PeerPtrs pp{};
  for(uint32_t k = 0; k < world; k++) pp.p[k] = (uint64_t *)peer_flags[k];
  k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(pp, rank, world, epoch, (uint32_t *)my_flags, d_timed_out);
  CU(cudaGetLastError());
  return 0;
}

/* CUDA IPC: export a cudaMalloc'ed buffer / map a peer's buffer into this process (64-byte opaque handle). */
extern "C" int ntt_cuda_ipc_export(int device, void *d_ptr, void *handle64)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, d_ptr));
  return 0;
}
extern "C" int ntt_cuda_ipc_open(int device, const void *handle64, void **d_ptr)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int ntt_cuda_ipc_close(int device, void *d_ptr)
{
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  CU(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}

/*
 * RNS batch: limb l (its own modulus and tables, plist[l]) transforms `batch_per_limb` polynomials at
 * d_a + l * batch_per_limb * N.  One launch per limb would leave each CTA one or two chunks, nothing to pipeline
 * and a ragged tail, so the limbs are issued round-robin on a few internal streams, each launch sized to a
 * fraction of the GPU (see g_min_chunks_per_cta); `stream` forks into them and joins again.
 */
/* All limbs in ONE launch per kernel (strided pass(es) + FP64 chunk kernel over the whole limb-major array): what an
 * RNS batch of N >= 2^14 in the FP64 range gets.  A launch per limb gives every persistent CTA three or four chunks
 * and leaves SMs idle between limbs; one launch gives it dozens (config 3: 0.99 -> see profiles). */
static int rns_multi_launch(int device, const ntt_cuda_params_t *const *plist, size_t limbs, uint64_t *d_a,
                            size_t batch_per_limb, int inverse, cudaStream_t st)
{
  const ntt_cuda_params_t &p0 = *plist[0];
  const Split              sp = make_split((int)p0.logn);
  static thread_local RingLimbs<true> lb;
  for(size_t l = 0; l < limbs; l++) lb.e[l] = *plist[l];
  lb.polys_per_limb = (uint32_t)batch_per_limb;
  const size_t total = limbs * batch_per_limb;
  /* rns_multi_ok has checked that every limb runs the FP64 chunk kernel: the strided passes can be FP64 as well */
  bool fp_pass = true;
  for(size_t l = 0; l < limbs; l++) fp_pass = fp_pass && fp_strided_path(*plist[l], sp.L, d_a, !inverse);
  uint32_t     s1    = 0;
  for(int k = 0; k < sp.ns; k++) s1 += sp.r[k];
  if(!inverse) {
    uint32_t s0 = 0;
    for(int k = 0; k < sp.ns; k++) {
      const int rc = fp_pass ? strided_fp_launch_multi(true, false, sp.r[k], device, lb, d_a, s0, total, st)
                             : ((k == sp.ns - 1) ? dispatch_strided_multi<true, 2>(device, sp.r[k], lb, d_a, s0, total, st)
                                                 : dispatch_strided_multi<true, 0>(device, sp.r[k], lb, d_a, s0, total, st));
      if(rc) return -1;
      s0 += sp.r[k];
    }
    return ring_fp_launch_multi_14(true, device, plist, limbs, batch_per_limb, d_a, st);
  }
  if(ring_fp_launch_multi_14(false, device, plist, limbs, batch_per_limb, d_a, st)) return -1;
  uint32_t s0 = s1;
  for(int k = sp.ns - 1; k >= 0; k--) {
    s0 -= sp.r[k];
    const int rc = fp_pass ? strided_fp_launch_multi(false, k == 0, sp.r[k], device, lb, d_a, s0, total, st)
                           : ((k == 0) ? dispatch_strided_multi<false, 1>(device, sp.r[k], lb, d_a, s0, total, st)
                                       : dispatch_strided_multi<false, 0>(device, sp.r[k], lb, d_a, s0, total, st));
    if(rc) return -1;
  }
  return 0;
}

static bool rns_multi_ok(const ntt_cuda_params_t *const *plist, size_t limbs, const uint64_t *d_a, size_t batch_per_limb,
                         int inverse)
{
  static int on = -1;
  if(on < 0) {
    const char *e = getenv("NTT_B200_NO_RNS_MULTI");
    on            = (e && e[0] == '1') ? 0 : 1;
  }
  if(!on || limbs < 2 || limbs > (size_t)RING_MAX_LIMBS || !ring_enabled()) return false;
  if(((uintptr_t)d_a & 127) != 0 || batch_per_limb >= ((size_t)1 << 31)) return false;
  const ntt_cuda_params_t &p0 = *plist[0];
  if(p0.logn < 14 || ((limbs * batch_per_limb) << (p0.logn - 14)) >= ((size_t)1 << 31)) return false;
  for(size_t l = 0; l < limbs; l++) {
    const ntt_cuda_params_t &p = *plist[l];
    if(p.logn != p0.logn || !p.lazy || p.fp64 != p0.fp64 || !use_fp64(p, !inverse)) return false;
  }
  return true;
}

extern "C" int ntt_cuda_rns(int device, const ntt_cuda_params_t *const *plist, size_t limbs, uint64_t *d_a,
                            size_t batch_per_limb, int inverse, void *stream)
{
  if(limbs == 0 || batch_per_limb == 0) return 0;
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  if(rns_multi_ok(plist, limbs, d_a, batch_per_limb, inverse))
    return rns_multi_launch(device, plist, limbs, d_a, batch_per_limb, inverse, (cudaStream_t)stream);
  constexpr int       NS = 4;
  static cudaStream_t side[64][NS];
  static cudaEvent_t  fork_ev[64], join_ev[64][NS];
  static bool         made[64] = {false};
  static std::mutex   lock[64];
  const int           dv = device & 63;
  /* the side streams and events are shared per device: two host threads forking into them at the same time would
   * wait on each other's fork record (and could start before their own stream's earlier work) */
  std::lock_guard<std::mutex> hold(lock[dv]);
  if(!made[dv]) {
    for(int i = 0; i < NS; i++) {
      CU(cudaStreamCreateWithFlags(&side[dv][i], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&join_ev[dv][i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&fork_ev[dv], cudaEventDisableTiming));
    made[dv] = true;
  }
  cudaStream_t user = (cudaStream_t)stream;
  const size_t limb_words = batch_per_limb << plist[0]->logn;
  const int    lanes = limbs < (size_t)NS ? (int)limbs : NS;
  CU(cudaEventRecord(fork_ev[dv], user));
  for(int i = 0; i < lanes; i++) CU(cudaStreamWaitEvent(side[dv][i], fork_ev[dv], 0));
  int rc = 0;
  /* N >= 2^15: every limb is a strided pass plus a chunk kernel.  Issued limb by limb the small strided grids of one
   * limb queue behind the SM-filling chunk kernels of the others; issued phase by phase (all strided passes of a
   * stream first, then its chunk kernels -- the per-limb order on a stream is unchanged) the two kinds of kernel do
   * not interleave.  Chunk kernels of `lanes` limbs run side by side, each on its share of the SMs. */
  size_t share = ((size_t)sm_count(device) + lanes - 1) / lanes;
  const size_t chunks_per_limb = batch_per_limb << (plist[0]->logn > 14 ? plist[0]->logn - 14 : 0);
  g_min_chunks_per_cta = (chunks_per_limb + share - 1) / share;
  if(g_min_chunks_per_cta < 1) g_min_chunks_per_cta = 1;
  const int first_phase = inverse ? PH_CHUNK : PH_STRIDED, second_phase = inverse ? PH_STRIDED : PH_CHUNK;
  for(int phase = 0; phase < 2 && !rc; phase++) {
    const int ph = phase == 0 ? first_phase : second_phase;
    for(size_t l = 0; l < limbs && !rc; l++) {
      const ntt_cuda_params_t &p  = *plist[l];
      cudaStream_t             st = side[dv][l % lanes];
      uint64_t *               d  = d_a + l * limb_words;
      if(inverse) {
        rc = p.lazy ? inverse_impl<false>(device, p, d, batch_per_limb, st, ph) : inverse_impl<true>(device, p, d, batch_per_limb, st, ph);
      } else {
        rc = p.lazy ? forward_impl<false>(device, p, d, batch_per_limb, st, nullptr, nullptr, ph)
                    : forward_impl<true>(device, p, d, batch_per_limb, st, nullptr, nullptr, ph);
      }
    }
  }
  g_min_chunks_per_cta = 0;
  for(int i = 0; i < lanes; i++) {
    CU(cudaEventRecord(join_ev[dv][i], side[dv][i]));
    CU(cudaStreamWaitEvent(user, join_ev[dv][i], 0));
  }
  return rc;
}

static int pointwise_launch(int device, const ntt_cuda_params_t *p, uint64_t *d_c, const uint64_t *d_a,
                           const uint64_t *d_b, size_t n, size_t b_mask, void *stream)
{
  if(n == 0) return 0;
  DevGuard g(device);
  if(!g.ok) return fail_msg("cudaSetDevice failed");
  /* mu = floor(2^128 / q) via two 128/64 divisions on the host */
  const u128     top  = ~(u128)0;
  u128           mu   = top / p->q; /* floor((2^128-1)/q) == floor(2^128/q) for q not a power of two */
  const uint64_t mu_hi = (uint64_t)(mu >> 64), mu_lo = (uint64_t)mu;
  size_t         grid  = (n + 255) / 256;
  const size_t   cap   = (size_t)sm_count(device) * 16;
  if(grid > cap) grid = cap;
  k_pointwise<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_c, d_a, d_b, n, b_mask, p->q, mu_hi, mu_lo);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int ntt_cuda_pointwise(int device, const ntt_cuda_params_t *p, uint64_t *d_c, const uint64_t *d_a,
                                  const uint64_t *d_b, size_t n, void *stream)
{
  return pointwise_launch(device, p, d_c, d_a, d_b, n, ~(size_t)0, stream);
}

/* c[i] = a[i] * b[i mod N] mod q: one polynomial b (N = 2^logn words) against every polynomial of the batch */
extern "C" int ntt_cuda_pointwise_bcast(int device, const ntt_cuda_params_t *p, uint64_t *d_c, const uint64_t *d_a,
                                        const uint64_t *d_b, size_t n, void *stream)
{
  return pointwise_launch(device, p, d_c, d_a, d_b, n, ((size_t)1 << p->logn) - 1, stream);
}
