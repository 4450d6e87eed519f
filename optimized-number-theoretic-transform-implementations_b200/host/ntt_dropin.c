/*
 * host/ntt_dropin.c -- reference-shaped single-polynomial entry points (host pointers).
 *
 * ntt_b200_fwd_ntt_ref_harvey[_lazy|_dbl] / ntt_b200_inv_ntt_ref_harvey take exactly the arguments of the
 * reference functions they replace (include/ntt_reference.h:13-65, src/ntt_reference.c:11-91) and run the
 * transform on the GPU.  Plans are cached per (direction, N, q, w[1], w[N/2]) -- w[N/2] is the root the
 * table was generated from (psi for forward tables, psi^-1 for inverse ones) -- so
 * the table upload and validation happen once per parameter set, as the reference's fixtures build their
 * tables once (tests/test_cases.h:313-321).
 */
#include <pthread.h>
#include <stdlib.h>

#include "../../include/ntt_b200.h"

typedef struct cache_entry {
  int                 inverse;
  uint64_t            N, q, w1, wroot, n_inv;
  ntt_b200_plan_t *   plan;
  struct cache_entry *next;
} cache_entry_t;

static cache_entry_t * g_cache = NULL;
static pthread_mutex_t g_lock  = PTHREAD_MUTEX_INITIALIZER;

static ntt_b200_plan_t *lookup(int inverse, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con,
                               uint64_t n_inv, uint64_t n_inv_con)
{
  if(!w || N < 2 || (N & (N - 1))) return NULL;
  const uint64_t   w1 = w[1], wroot = w[N / 2];
  ntt_b200_plan_t *plan = NULL;
  pthread_mutex_lock(&g_lock);
  for(cache_entry_t *e = g_cache; e; e = e->next) {
    if(e->inverse == inverse && e->N == N && e->q == q && e->w1 == w1 && e->wroot == wroot &&
       (!inverse || e->n_inv == n_inv)) {
      plan = e->plan;
      break;
    }
  }
  if(!plan) {
    const int rc = inverse ? ntt_b200_plan_create(&plan, 0, N, q, NULL, NULL, w, w_con, n_inv, n_inv_con)
                           : ntt_b200_plan_create(&plan, 0, N, q, w, w_con, NULL, NULL, 0, 0);
    if(rc == NTT_B200_SUCCESS) {
      cache_entry_t *e = calloc(1, sizeof(*e));
      if(e) {
        e->inverse = inverse;
        e->N       = N;
        e->q       = q;
        e->w1      = w1;
        e->wroot   = wroot;
        e->n_inv   = n_inv;
        e->plan    = plan;
        e->next    = g_cache;
        g_cache    = e;
      }
    } else {
      plan = NULL;
    }
  }
  pthread_mutex_unlock(&g_lock);
  return plan;
}

void ntt_b200_dropin_reset(void)
{
  pthread_mutex_lock(&g_lock);
  while(g_cache) {
    cache_entry_t *e = g_cache;
    g_cache          = e->next;
    ntt_b200_plan_destroy(e->plan);
    free(e);
  }
  pthread_mutex_unlock(&g_lock);
}

int ntt_b200_fwd_ntt_ref_harvey(uint64_t a[], uint64_t N, uint64_t q, const uint64_t w[], const uint64_t w_con[])
{
  ntt_b200_plan_t *plan = lookup(0, N, q, w, w_con, 0, 0);
  if(!plan) return NTT_B200_ERROR;
  return ntt_b200_fwd_batch_host(plan, a, 1);
}

int ntt_b200_fwd_ntt_ref_harvey_lazy(uint64_t a[], uint64_t N, uint64_t q, const uint64_t w[],
                                     const uint64_t w_con[])
{
  /* the fully reduced output is a member of the lazy range [0,4q) */
  return ntt_b200_fwd_ntt_ref_harvey(a, N, q, w, w_con);
}

int ntt_b200_fwd_ntt_ref_harvey_dbl(uint64_t a1[], uint64_t a2[], uint64_t N, uint64_t q, const uint64_t w[],
                                    const uint64_t w_con[])
{
  ntt_b200_plan_t *plan = lookup(0, N, q, w, w_con, 0, 0);
  if(!plan) return NTT_B200_ERROR;
  if(ntt_b200_fwd_batch_host(plan, a1, 1)) return NTT_B200_ERROR;
  return ntt_b200_fwd_batch_host(plan, a2, 1);
}

int ntt_b200_inv_ntt_ref_harvey(uint64_t a[], uint64_t N, uint64_t q, uint64_t n_inv, uint64_t n_inv_con,
                                uint64_t word_size, const uint64_t w[], const uint64_t w_con[])
{
  if(word_size != 64) return NTT_B200_ERROR; /* the reference path only ever passes WORD_SIZE = 64 */
  ntt_b200_plan_t *plan = lookup(1, N, q, w, w_con, n_inv, n_inv_con);
  if(!plan) return NTT_B200_ERROR;
  return ntt_b200_inv_batch_host(plan, a, 1);
}
