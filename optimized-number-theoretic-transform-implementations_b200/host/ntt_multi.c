/*
 * host/ntt_multi.c -- the multi-GPU form of the batch API: a device list in, the
 * batch (or the RNS limb set) sharded over it, no data-path collective.
 *
 * Polynomials and RNS limbs are independent transforms (the reference has no cross-polynomial operation at all:
 * src/ntt_reference.c:11-66 works on one array), so device i simply owns a contiguous range of the batch and its
 * own copy of the twiddle tables.  Device-resident calls are asynchronous launches issued from the calling thread,
 * one device after the other; host-buffer calls run one host thread per device, each driving that device's
 * copy/compute pipeline (host/ntt_plan.c) on its contiguous shard of the caller's array.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ntt_b200.h"

#define MULTI_MAX_DEVICES 64

struct ntt_b200_multi {
  int              n;
  int              device[MULTI_MAX_DEVICES];
  ntt_b200_plan_t *plan[MULTI_MAX_DEVICES];
};

static __thread char g_merr[512];
const char *ntt_b200_multi_last_error(void) { return g_merr; }

static int merr(const char *what, const char *detail)
{
  snprintf(g_merr, sizeof(g_merr), "%s%s%s", what, detail && detail[0] ? ": " : "", detail ? detail : "");
  return NTT_B200_ERROR;
}

int ntt_b200_multi_create(ntt_b200_multi_t **out, const int *devices, int n_devices, uint64_t N, uint64_t q, uint64_t psi)
{
  if(!out) return merr("multi pointer is NULL", NULL);
  *out = NULL;
  if(n_devices < 1 || n_devices > MULTI_MAX_DEVICES) return merr("device count out of range", NULL);
  ntt_b200_multi_t *m = calloc(1, sizeof(*m));
  if(!m) return merr("out of host memory", NULL);
  m->n = n_devices;
  for(int i = 0; i < n_devices; i++) {
    m->device[i] = devices ? devices[i] : i;
    if(ntt_b200_plan_create_psi(&m->plan[i], m->device[i], N, q, psi)) {
      merr("plan creation failed", ntt_b200_last_error());
      ntt_b200_multi_destroy(m);
      return NTT_B200_ERROR;
    }
  }
  *out = m;
  return NTT_B200_SUCCESS;
}

int ntt_b200_multi_destroy(ntt_b200_multi_t *m)
{
  if(!m) return NTT_B200_SUCCESS;
  for(int i = 0; i < m->n; i++) {
    if(m->plan[i]) ntt_b200_plan_destroy(m->plan[i]);
  }
  free(m);
  return NTT_B200_SUCCESS;
}

int ntt_b200_multi_devices(const ntt_b200_multi_t *m) { return m ? m->n : 0; }

const ntt_b200_plan_t *ntt_b200_multi_plan(const ntt_b200_multi_t *m, int index)
{
  return (m && index >= 0 && index < m->n) ? m->plan[index] : NULL;
}

/* contiguous shard of `batch` units owned by part `index` of `parts`: sizes differ by at most one */
void ntt_b200_shard_range(size_t batch, int parts, int index, size_t *first, size_t *count)
{
  size_t f = 0, c = 0;
  if(parts > 0 && index >= 0 && index < parts) {
    const size_t base = batch / (size_t)parts, extra = batch % (size_t)parts;
    f = (size_t)index * base + ((size_t)index < extra ? (size_t)index : extra);
    c = base + ((size_t)index < extra ? 1 : 0);
  }
  if(first) *first = f;
  if(count) *count = c;
}

/* ---- device-resident -------------------------------------------------------------------------------- */

static int multi_device_apply(const ntt_b200_multi_t *m, uint64_t *const *d_a, const size_t *batch, void *const *streams,
                              int inverse)
{
  if(!m || !d_a || !batch) return merr("NULL argument", NULL);
  for(int i = 0; i < m->n; i++) {
    if(batch[i] == 0) continue;
    void *    st = streams ? streams[i] : NULL;
    const int rc = inverse ? ntt_b200_inv_batch(m->plan[i], d_a[i], batch[i], st)
                           : ntt_b200_fwd_batch(m->plan[i], d_a[i], batch[i], st);
    if(rc) return merr(inverse ? "inverse NTT" : "forward NTT", ntt_b200_last_error());
  }
  return NTT_B200_SUCCESS;
}

int ntt_b200_multi_fwd_batch(const ntt_b200_multi_t *m, uint64_t *const *d_a, const size_t *batch, void *const *streams)
{
  return multi_device_apply(m, d_a, batch, streams, 0);
}
int ntt_b200_multi_inv_batch(const ntt_b200_multi_t *m, uint64_t *const *d_a, const size_t *batch, void *const *streams)
{
  return multi_device_apply(m, d_a, batch, streams, 1);
}

int ntt_b200_multi_sync(const ntt_b200_multi_t *m)
{
  if(!m) return merr("NULL argument", NULL);
  for(int i = 0; i < m->n; i++) {
    if(ntt_b200_device_sync(m->device[i])) return merr("device sync", ntt_b200_last_error());
  }
  return NTT_B200_SUCCESS;
}

/* ---- host-resident: one host thread per device ------------------------------------------------------------ */

typedef struct {
  const ntt_b200_plan_t *plan;
  uint64_t *             h;
  const uint64_t *       d_m;
  size_t                 batch;
  int                    mode; /* 0 forward, 1 inverse, 2 forward-multiply-inverse */
  int                    rc;
  char                   err[256];
} host_job_t;

static void *host_worker(void *arg)
{
  host_job_t *j = (host_job_t *)arg;
  if(j->batch == 0) {
    j->rc = 0;
    return NULL;
  }
  if(j->mode == 0) j->rc = ntt_b200_fwd_batch_host(j->plan, j->h, j->batch);
  else if(j->mode == 1) j->rc = ntt_b200_inv_batch_host(j->plan, j->h, j->batch);
  else j->rc = ntt_b200_fwd_mul_inv_batch_host(j->plan, j->h, j->d_m, j->batch);
  if(j->rc) snprintf(j->err, sizeof(j->err), "%s", ntt_b200_last_error()); /* the error text is thread-local */
  return NULL;
}

static int multi_host_apply(const ntt_b200_multi_t *m, uint64_t *h_a, size_t batch, int mode, const uint64_t *const *d_m)
{
  if(!m || !h_a) return merr("NULL argument", NULL);
  const uint64_t N = ntt_b200_plan_n(m->plan[0]);
  host_job_t     job[MULTI_MAX_DEVICES];
  pthread_t      tid[MULTI_MAX_DEVICES];
  int            started[MULTI_MAX_DEVICES];
  memset(job, 0, sizeof(job));
  for(int i = 0; i < m->n; i++) {
    size_t first, count;
    ntt_b200_shard_range(batch, m->n, i, &first, &count);
    job[i].plan  = m->plan[i];
    job[i].h     = h_a + first * N;
    job[i].batch = count;
    job[i].mode  = mode;
    job[i].d_m   = d_m ? d_m[i] : NULL;
    started[i]   = (m->n > 1) && pthread_create(&tid[i], NULL, host_worker, &job[i]) == 0;
    if(!started[i]) host_worker(&job[i]); /* single device, or no thread to be had: run it here */
  }
  int rc = NTT_B200_SUCCESS;
  for(int i = 0; i < m->n; i++) {
    if(started[i]) pthread_join(tid[i], NULL);
    if(job[i].rc && !rc) rc = merr("device shard failed", job[i].err);
  }
  return rc;
}

int ntt_b200_multi_fwd_batch_host(const ntt_b200_multi_t *m, uint64_t *h_a, size_t batch)
{
  return multi_host_apply(m, h_a, batch, 0, NULL);
}
int ntt_b200_multi_inv_batch_host(const ntt_b200_multi_t *m, uint64_t *h_a, size_t batch)
{
  return multi_host_apply(m, h_a, batch, 1, NULL);
}
int ntt_b200_multi_fwd_mul_inv_batch_host(const ntt_b200_multi_t *m, uint64_t *h_a, const uint64_t *const *d_m, size_t batch)
{
  return multi_host_apply(m, h_a, batch, 2, d_m);
}

/* ---- RNS limbs sharded over devices ----------------------------------------------------------------------- */

/* plans[l] may live on any device; d_limb[l] points to limb l's batch_per_limb polynomials ON THAT DEVICE.
 * Consecutive limbs that share a device and are contiguous in memory go down as one ntt_b200_fwd_rns call (which
 * spreads them over that device's internal streams); everything is asynchronous on each device's default stream. */
static int rns_multi(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *const *d_limb, size_t batch_per_limb,
                     int inverse)
{
  if(!plans || !d_limb || limbs == 0) return merr("NULL argument", NULL);
  size_t l = 0;
  while(l < limbs) {
    if(!plans[l] || !d_limb[l]) return merr("NULL plan or limb pointer", NULL);
    const int      dev   = ntt_b200_plan_device(plans[l]);
    const uint64_t words = ntt_b200_plan_n(plans[l]) * batch_per_limb;
    size_t         run   = 1;
    while(l + run < limbs && plans[l + run] && ntt_b200_plan_device(plans[l + run]) == dev &&
          d_limb[l + run] == d_limb[l] + run * words)
      run++;
    const int rc = inverse ? ntt_b200_inv_rns(plans + l, run, d_limb[l], batch_per_limb, NULL)
                           : ntt_b200_fwd_rns(plans + l, run, d_limb[l], batch_per_limb, NULL);
    if(rc) return merr("RNS transform", ntt_b200_last_error());
    l += run;
  }
  return NTT_B200_SUCCESS;
}

int ntt_b200_fwd_rns_multi(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *const *d_limb, size_t batch_per_limb)
{
  return rns_multi(plans, limbs, d_limb, batch_per_limb, 0);
}
int ntt_b200_inv_rns_multi(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *const *d_limb, size_t batch_per_limb)
{
  return rns_multi(plans, limbs, d_limb, batch_per_limb, 1);
}
