/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin export layer compiled TOGETHER WITH the reference's own, unmodified sources (taken where they lie
 * under /root/reference by oracle/Makefile; nothing is copied into this repo).  It gives Python/ctypes
 * access to: the 19 fixture parameter sets (tests/test_cases.h:145-208), the reference table builders
 * (include/internal/pre_compute.h), the reference transforms, and a pthread "one polynomial per thread"
 * timing loop that follows the reference's MEASURE methodology (tests/measurements.h:38-75).
 * The result lands in oracle/_ref/ (git-ignored, shipped to the GPU box as a prebuilt .so).
 */
#define _GNU_SOURCE
#include <float.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ntt_radix4.h"
#include "ntt_radix4x4.h"
#include "ntt_reference.h"
#include "ntt_seal.h"
#include "pre_compute.h"
#include "test_cases.h"

#ifdef AVX512_IFMA_SUPPORT
#  include "ntt_avx512_ifma.h"
#  include "ntt_hexl.h"
#endif

#define API __attribute__((visibility("default")))

/* ---- fixtures ------------------------------------------------------------------------------- */

API int ref_num_cases(void) { return (int)NUM_OF_TEST_CASES; }

/* out = {m, q, w(psi), w_inv, n_inv.op} exactly as written in tests/test_cases.h */
API void ref_case_params(int idx, uint64_t out[5])
{
  out[0] = tests[idx].m;
  out[1] = tests[idx].q;
  out[2] = tests[idx].w;
  out[3] = tests[idx].w_inv;
  out[4] = (uint64_t)tests[idx].n_inv.op;
}

API int ref_has_ifma(void)
{
#ifdef AVX512_IFMA_SUPPORT
  return 1;
#else
  return 0;
#endif
}

/* ---- table builders (pre_compute.h) ----------------------------------------------------------- */

typedef struct calc_w_args_s {
  uint64_t *out;
  uint64_t  w, N, q, width;
} calc_w_args_t;

static void *calc_w_thread(void *p)
{
  calc_w_args_t *a = (calc_w_args_t *)p;
  calc_w(a->out, a->w, a->N, a->q, a->width);
  return NULL;
}

API void ref_calc_w(uint64_t *out, uint64_t w, uint64_t N, uint64_t q, uint64_t width)
{
  /* calc_w keeps an N-word VLA on the stack (pre_compute.h:44): 32 MiB at N=2^22, so it runs on a
   * thread whose stack is sized for it */
  calc_w_args_t  args = {out, w, N, q, width};
  pthread_attr_t attr;
  pthread_t      th;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, (size_t)N * 8 + (16UL << 20));
  pthread_create(&th, &attr, calc_w_thread, &args);
  pthread_join(th, NULL);
  pthread_attr_destroy(&attr);
}
API void ref_calc_w_con(uint64_t *out, const uint64_t *w, uint64_t N, uint64_t q, uint64_t word_size)
{
  calc_w_con(out, w, N, q, word_size);
}
API uint64_t ref_calc_ninv_con(uint64_t ninv, uint64_t q, uint64_t word_size)
{
  return calc_ninv_con(ninv, q, word_size);
}
API uint64_t ref_bit_rev_idx(uint64_t idx, uint64_t width) { return bit_rev_idx(idx, width); }
API void ref_expand_w(uint64_t *out, const uint64_t *w, uint64_t N, uint64_t q) { expand_w(out, w, N, q); }

/* ---- transforms --------------------------------------------------------------------------------- */

API void ref_fwd_lazy(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *wc)
{
  fwd_ntt_ref_harvey_lazy(a, N, q, w, wc);
}
API void ref_fwd(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *wc)
{
  fwd_ntt_ref_harvey(a, N, q, w, wc);
}
API void ref_fwd_dbl(uint64_t *a, uint64_t *b, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *wc)
{
  fwd_ntt_ref_harvey_dbl(a, b, N, q, w, wc);
}
API void ref_inv(uint64_t *a, uint64_t N, uint64_t q, uint64_t ninv, uint64_t ninv_con, const uint64_t *w,
                 const uint64_t *wc)
{
  const mul_op_t n = {ninv, ninv_con};
  inv_ntt_ref_harvey(a, N, q, n, WORD_SIZE, w, wc);
}
API void ref_fwd_seal(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *wc)
{
  fwd_ntt_seal(a, N, q, w, wc);
}
API void ref_inv_seal(uint64_t *a, uint64_t N, uint64_t q, uint64_t ninv, uint64_t ninv_con, const uint64_t *w,
                      const uint64_t *wc)
{
  inv_ntt_seal(a, N, q, ninv, ninv_con, w, wc);
}
/* radix-4 variants take the expand_w tables (2N words) */
API void ref_fwd_radix4(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w4, const uint64_t *wc4)
{
  fwd_ntt_radix4(a, N, q, w4, wc4);
}
API void ref_inv_radix4(uint64_t *a, uint64_t N, uint64_t q, uint64_t ninv, uint64_t ninv_con, const uint64_t *w4,
                        const uint64_t *wc4)
{
  const mul_op_t n = {ninv, ninv_con};
  inv_ntt_radix4(a, N, q, n, w4, wc4);
}
API void ref_fwd_radix4x4(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w4, const uint64_t *wc4)
{
  fwd_ntt_radix4x4(a, N, q, w4, wc4);
}

/* ---- CPU baseline: one polynomial per thread -------------------------------------------------------- */

enum {
  V_FWD_REF = 0,
  V_FWD_SEAL,
  V_FWD_RADIX4,
  V_FWD_RADIX4X4,
  V_INV_REF,
  V_INV_SEAL,
  V_INV_RADIX4,
  V_FWD_R4_IFMA,
  V_FWD_R4_IFMA_UNORDERED,
  V_FWD_R4R2_IFMA,
  V_FWD_R2_16_IFMA,
  V_FWD_HEXL,
  V_COUNT
};

static const char *const variant_names[V_COUNT] = {
  "fwd_ntt_ref_harvey",
  "fwd_ntt_seal",
  "fwd_ntt_radix4",
  "fwd_ntt_radix4x4",
  "inv_ntt_ref_harvey",
  "inv_ntt_seal",
  "inv_ntt_radix4",
  "fwd_ntt_radix4_avx512_ifma",
  "fwd_ntt_radix4_avx512_ifma_unordered",
  "fwd_ntt_r4r2_avx512_ifma",
  "fwd_ntt_r2_16_avx512_ifma",
  "fwd_ntt_radix2_hexl",
};

API int         ref_num_variants(void) { return V_COUNT; }
API const char *ref_variant_name(int v) { return (v >= 0 && v < V_COUNT) ? variant_names[v] : ""; }

typedef struct bench_ctx_s {
  int          variant;
  test_case_t *t;
  uint64_t *   a;
  size_t       calls;
  double       seconds;
  pthread_barrier_t *bar;
} bench_ctx_t;

static int call_variant(int v, test_case_t *t, uint64_t *a)
{
  const uint64_t n = t->n, q = t->q;
  switch(v) {
    case V_FWD_REF: fwd_ntt_ref_harvey(a, n, q, t->w_powers.ptr, t->w_powers_con.ptr); return 0;
    case V_FWD_SEAL: fwd_ntt_seal(a, n, q, t->w_powers.ptr, t->w_powers_con.ptr); return 0;
    case V_FWD_RADIX4: fwd_ntt_radix4(a, n, q, t->w_powers_r4.ptr, t->w_powers_con_r4.ptr); return 0;
    case V_FWD_RADIX4X4: fwd_ntt_radix4x4(a, n, q, t->w_powers_r4.ptr, t->w_powers_con_r4.ptr); return 0;
    case V_INV_REF:
      inv_ntt_ref_harvey(a, n, q, t->n_inv, WORD_SIZE, t->w_inv_powers.ptr, t->w_inv_powers_con.ptr);
      return 0;
    case V_INV_SEAL:
      inv_ntt_seal(a, n, q, t->n_inv.op, t->n_inv.con, t->w_inv_powers.ptr, t->w_inv_powers_con.ptr);
      return 0;
    case V_INV_RADIX4:
      inv_ntt_radix4(a, n, q, t->n_inv, t->w_inv_powers_r4.ptr, t->w_inv_powers_con_r4.ptr);
      return 0;
#ifdef AVX512_IFMA_SUPPORT
    case V_FWD_R4_IFMA:
      fwd_ntt_radix4_avx512_ifma(a, n, q, t->w_powers_r4_avx512_ifma.ptr, t->w_powers_con_r4_avx512_ifma.ptr);
      return 0;
    case V_FWD_R4_IFMA_UNORDERED:
      fwd_ntt_radix4_avx512_ifma_unordered(a, n, q, t->w_powers_r4_avx512_ifma_unordered.ptr,
                                           t->w_powers_con_r4_avx512_ifma_unordered.ptr);
      return 0;
    case V_FWD_R4R2_IFMA:
      fwd_ntt_r4r2_avx512_ifma(a, n, q, t->w_powers_r4r2_avx512_ifma.ptr, t->w_powers_con_r4r2_avx512_ifma.ptr);
      return 0;
    case V_FWD_R2_16_IFMA:
      fwd_ntt_r2_16_avx512_ifma(a, n, q, t->w_powers_r2_16_avx512_ifma.ptr,
                                t->w_powers_con_r2_16_avx512_ifma.ptr);
      return 0;
    case V_FWD_HEXL:
      fwd_ntt_radix2_hexl(a, n, q, t->w_powers_hexl.ptr, t->w_powers_con_hexl.ptr);
      return 0;
#endif
    default: return -1;
  }
}

static double now_sec(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *bench_thread(void *p)
{
  bench_ctx_t *c = (bench_ctx_t *)p;
  /* warm-up as tests/measurements.h:38 (WARMUP 10) */
  for(int i = 0; i < 10; i++) call_variant(c->variant, c->t, c->a);
  pthread_barrier_wait(c->bar);
  const double t0 = now_sec();
  /* like the reference's timed loop, the previous output is fed back as the next input */
  for(size_t i = 0; i < c->calls; i++) call_variant(c->variant, c->t, c->a);
  c->seconds = now_sec() - t0;
  return NULL;
}

/* Runs `threads` pthreads, each transforming its own polynomial `calls` times with reference variant
 * `variant` on a case built by the reference's own _init_test (tests/test_cases.h:212-311) for
 * (m, q, psi, psi_inv, n_inv).  Returns aggregate transforms per second (sum over threads of
 * calls / thread_seconds), or a negative value if the variant is unavailable. */
API double ref_bench_variant(int variant, uint64_t m, uint64_t q, uint64_t psi, uint64_t psi_inv, uint64_t n_inv,
                             int threads, uint64_t calls, uint64_t seed)
{
  if(variant < 0 || variant >= V_COUNT || threads < 1) return -1.0;
#ifndef AVX512_IFMA_SUPPORT
  if(variant >= V_FWD_R4_IFMA) return -1.0;
#endif
  test_case_t t;
  memset(&t, 0, sizeof(t));
  t.m        = m;
  t.q        = q;
  t.w        = psi;
  t.w_inv    = psi_inv;
  t.n_inv.op = n_inv;
  _init_test(&t);

  pthread_t *       th  = calloc((size_t)threads, sizeof(*th));
  bench_ctx_t *     ctx = calloc((size_t)threads, sizeof(*ctx));
  aligned64_ptr_t * buf = calloc((size_t)threads, sizeof(*buf));
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, (unsigned)threads);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 64UL << 20);

  uint64_t s = seed;
  for(int i = 0; i < threads; i++) {
    allocate_aligned_array(&buf[i], t.n);
    for(uint64_t k = 0; k < t.n; k++) {
      uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
      z          = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
      z          = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
      buf[i].ptr[k] = (z ^ (z >> 31)) % q;
    }
    ctx[i].variant = variant;
    ctx[i].t       = &t;
    ctx[i].a       = buf[i].ptr;
    ctx[i].calls   = calls;
    ctx[i].bar     = &bar;
    pthread_create(&th[i], &attr, bench_thread, &ctx[i]);
  }
  double rate = 0.0;
  for(int i = 0; i < threads; i++) {
    pthread_join(th[i], NULL);
    rate += (double)calls / ctx[i].seconds;
    free_aligned_array(&buf[i]);
  }
  pthread_barrier_destroy(&bar);
  pthread_attr_destroy(&attr);
  _destroy_test(&t);
  free(th);
  free(ctx);
  free(buf);
  return rate;
}

/* Run one variant once on caller data with reference-built tables (used to cross-check variants) */
API int ref_run_variant(int variant, uint64_t m, uint64_t q, uint64_t psi, uint64_t psi_inv, uint64_t n_inv,
                        uint64_t *a)
{
  test_case_t t;
  memset(&t, 0, sizeof(t));
  t.m        = m;
  t.q        = q;
  t.w        = psi;
  t.w_inv    = psi_inv;
  t.n_inv.op = n_inv;
  _init_test(&t);
  const int rc = call_variant(variant, &t, a);
  _destroy_test(&t);
  return rc;
}
