/*
 * host/ntt_plan.c -- the C host side of libntt_b200: plans, batch entry points, host-buffer pipeline.
 *
 * Mirrors the reference's operator interface for the NTT hot path (include/ntt_reference.h:13-65): same
 * argument meaning, in-place transforms, caller-owned buffers, SUCCESS 0 / ERROR -1
 * (include/internal/defs.h:20-21).  All arithmetic on coefficients happens in the CUDA layer
 * (csrc/ntt_kernels.cu); there is no CPU implementation of the transform in this library.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ntt_b200.h"
#include "../csrc/ntt_cuda.h"
#include "ntt_math.h"

typedef unsigned __int128 u128;

#define LAZY_MAX_QBITS 56 /* lazy path needs (4 + 10*24) * q < 2^64 */
#define HOST_PIPE_DEPTH_MAX 8
static int    g_pipe_depth = 4;                  /* chunks in flight, each on its own stream */
static size_t g_pipe_bytes = (size_t)32 << 20;   /* bytes per chunk */

struct ntt_b200_plan {
  int      device;
  uint64_t N, q;
  unsigned logn;
  int      has_fwd, has_inv;
  uint64_t n_inv, n_inv_con;
  uint64_t w_inv_1; /* w_inv[1] = psi^-(N/2), kept to rebuild the stage-0 constants */
  ntt_cuda_params_t params;
  /* device memory: kernel tables and the reference-format copies kept for export */
  void *    d_fwd_wu, *d_fwd_qq, *d_inv_wu, *d_inv_qq;
  void *    d_fwd_ct_wu, *d_fwd_ct_qq, *d_inv_ct_wu, *d_inv_ct_qq; /* pass-C layout of the last four stages */
  void *    d_fwd_fd, *d_inv_fd, *d_fwd_ct_fd, *d_inv_ct_fd;         /* FP64 path twiddles (q < 2^49) */
  uint64_t *d_w, *d_w_con, *d_w_inv, *d_w_inv_con;
  /* host-buffer pipeline (created on first use) */
  pthread_mutex_t pipe_lock;
  void *          pipe_stream[HOST_PIPE_DEPTH_MAX];
  uint64_t *      pipe_buf[HOST_PIPE_DEPTH_MAX];
  size_t          pipe_polys;
  int             pipe_depth;
};

static __thread char g_error[512];

static int set_error(const char *fmt, const char *detail)
{
  snprintf(g_error, sizeof(g_error), fmt, detail ? detail : "");
  return NTT_B200_ERROR;
}
static int cuda_error(const char *where)
{
  snprintf(g_error, sizeof(g_error), "%s: %s", where, ntt_cuda_error());
  return NTT_B200_ERROR;
}

const char *ntt_b200_last_error(void) { return g_error; }
int         ntt_b200_device_count(void) { return ntt_cuda_device_count(); }
const char *ntt_b200_version(void) { return "ntt_b200 0.2 sm_100a"; }

int ntt_b200_configure(const char *key, int value)
{
  if(ntt_cuda_configure(key, value)) return cuda_error("configure");
  return NTT_B200_SUCCESS;
}

/* ---- multiplier constants ------------------------------------------------------------------------ */

static ntt_cuda_mulc_t make_mulc(uint64_t w, uint64_t q, int lazy)
{
  ntt_cuda_mulc_t m;
  memset(&m, 0, sizeof(m));
  w %= q;
  m.w0 = (uint32_t)w;
  m.w1 = (uint32_t)(w >> 32);
  if(lazy) {
    const uint64_t u = (uint64_t)((((u128)w) << 32) % q);
    m.u0             = (uint32_t)u;
    m.u1             = (uint32_t)(u >> 32);
    m.wq             = (uint32_t)nttm_shoup(w, q, 30);
    m.uq             = (uint32_t)nttm_shoup(u, q, 30);
  } else {
    const uint64_t c = nttm_shoup(w, q, 64);
    m.u0             = (uint32_t)c;
    m.u1             = (uint32_t)(c >> 32);
  }
  return m;
}

static int check_shape(uint64_t N, uint64_t q, int device)
{
  if(N < 2 || (N & (N - 1)) != 0 || nttm_log2(N) > NTT_B200_MAX_LOGN)
    return set_error("N must be a power of two in [2, 2^24]%s", NULL);
  if(q < 3 || (q & 1) == 0 || (q >> 62) != 0) return set_error("q must be odd and 3 <= q < 2^62%s", NULL);
  const int ndev = ntt_cuda_device_count();
  if(ndev <= 0) return set_error("no CUDA device available (this library has no CPU fallback)%s", NULL);
  if(device < 0 || device >= ndev) return set_error("device index out of range%s", NULL);
  return NTT_B200_SUCCESS;
}

static void fill_params(ntt_b200_plan_t *pl)
{
  ntt_cuda_params_t *p = &pl->params;
  memset(p, 0, sizeof(*p));
  p->q         = pl->q;
  p->neg2q     = (uint64_t)0 - 2 * pl->q;
  p->negq      = (uint64_t)0 - pl->q;
  p->c10q      = 10 * pl->q;
  p->logn      = pl->logn;
  p->lazy      = nttm_bitlen(pl->q) <= LAZY_MAX_QBITS ? 1u : 0u;
  p->red_shift = nttm_bitlen(pl->q) > 9 ? nttm_bitlen(pl->q) - 9 : 0;
  p->red_mu    = (uint32_t)((((u128)1) << (32 + p->red_shift)) / pl->q);
  /* FP64 ring kernel: q <= 2^50 - 2048 and a chunk size the ring kernel serves (N >= 2^10) */
  p->fp64 = 0;
  if(pl->logn >= 10) {
    if(pl->q <= ((uint64_t)1 << 49) - 1024) p->fp64 = 1;
    else if(pl->q <= ((uint64_t)1 << 50) - 2048) p->fp64 = 2;
  }
  p->q_fd    = (double)pl->q;
  p->qinv_fd = 1.0 / (double)pl->q;
}

static void plan_free(ntt_b200_plan_t *pl)
{
  if(!pl) return;
  void *dev_ptrs[] = {pl->d_fwd_wu,    pl->d_fwd_qq,    pl->d_inv_wu,    pl->d_inv_qq, pl->d_w,    pl->d_w_con,
                      pl->d_w_inv,     pl->d_w_inv_con, pl->d_fwd_ct_wu, pl->d_fwd_ct_qq, pl->d_inv_ct_wu,
                      pl->d_inv_ct_qq, pl->d_fwd_fd,    pl->d_inv_fd,    pl->d_fwd_ct_fd, pl->d_inv_ct_fd};
  for(size_t i = 0; i < sizeof(dev_ptrs) / sizeof(dev_ptrs[0]); i++) {
    if(dev_ptrs[i]) ntt_cuda_free(pl->device, dev_ptrs[i]);
  }
  for(int i = 0; i < HOST_PIPE_DEPTH_MAX; i++) {
    if(pl->pipe_buf[i]) ntt_cuda_free(pl->device, pl->pipe_buf[i]);
    if(pl->pipe_stream[i]) ntt_cuda_stream_destroy(pl->device, pl->pipe_stream[i]);
  }
  pthread_mutex_destroy(&pl->pipe_lock);
  free(pl);
}

static ntt_b200_plan_t *plan_alloc(int device, uint64_t N, uint64_t q)
{
  ntt_b200_plan_t *pl = calloc(1, sizeof(*pl));
  if(!pl) return NULL;
  pl->device = device;
  pl->N      = N;
  pl->q      = q;
  pl->logn   = nttm_log2(N);
  pthread_mutex_init(&pl->pipe_lock, NULL);
  fill_params(pl);
  return pl;
}

/* d_w (reference format, on device) -> kernel tables + companion table */
static int build_direction(ntt_b200_plan_t *pl, const uint64_t *d_w, void **wu, void **qq, uint64_t **con)
{
  const size_t n = (size_t)pl->N;
  if(ntt_cuda_malloc(pl->device, wu, n * 16)) return cuda_error("table alloc");
  if(ntt_cuda_malloc(pl->device, qq, n * 8)) return cuda_error("table alloc");
  if(ntt_cuda_malloc(pl->device, (void **)con, n * 8)) return cuda_error("table alloc");
  if(ntt_cuda_build_tables(pl->device, &pl->params, d_w, *wu, *qq, *con, pl->N, NULL))
    return cuda_error("table build");
  return NTT_B200_SUCCESS;
}

/* pass-C copies of the last four stages (lazy path, N >= 2^10: the sizes the ring kernels serve) */
static int build_ctables(ntt_b200_plan_t *pl, const void *wu, const void *qq, void **ct_wu, void **ct_qq)
{
  if(!pl->params.lazy || pl->logn < 10) return NTT_B200_SUCCESS;
  const size_t entries = (size_t)15 << (pl->logn - 4);
  if(ntt_cuda_malloc(pl->device, ct_wu, entries * 16) || ntt_cuda_malloc(pl->device, ct_qq, entries * 8))
    return cuda_error("table alloc");
  if(ntt_cuda_build_ctables(pl->device, &pl->params, wu, qq, *ct_wu, *ct_qq, NULL)) return cuda_error("table build");
  return NTT_B200_SUCCESS;
}

/* FP64 twiddles of one direction (only for plans the FP64 kernel can serve) */
static int build_fd(ntt_b200_plan_t *pl, const uint64_t *d_w, void **fd, void **ct_fd)
{
  if(!pl->params.fp64) return NTT_B200_SUCCESS;
  const size_t n = (size_t)pl->N;
  if(ntt_cuda_malloc(pl->device, fd, n * 16) || ntt_cuda_malloc(pl->device, ct_fd, (size_t)15 * (n / 16) * 16))
    return cuda_error("table alloc");
  if(ntt_cuda_build_fd_tables(pl->device, &pl->params, d_w, *fd, *ct_fd, NULL)) return cuda_error("table build");
  return NTT_B200_SUCCESS;
}

static void publish_tables(ntt_b200_plan_t *pl)
{
  pl->params.fwd_fd    = pl->d_fwd_fd;
  pl->params.inv_fd    = pl->d_inv_fd;
  pl->params.fwd_ct_fd = pl->d_fwd_ct_fd;
  pl->params.inv_ct_fd = pl->d_inv_ct_fd;
  pl->params.fwd_wu    = pl->d_fwd_wu;
  pl->params.fwd_qq    = pl->d_fwd_qq;
  pl->params.inv_wu    = pl->d_inv_wu;
  pl->params.inv_qq    = pl->d_inv_qq;
  pl->params.fwd_ct_wu = pl->d_fwd_ct_wu;
  pl->params.fwd_ct_qq = pl->d_fwd_ct_qq;
  pl->params.inv_ct_wu = pl->d_inv_ct_wu;
  pl->params.inv_ct_qq = pl->d_inv_ct_qq;
}

static int finish_inverse_constants(ntt_b200_plan_t *pl, uint64_t w_inv_1)
{
  pl->w_inv_1 = w_inv_1 % pl->q;
  const int lazy      = (int)pl->params.lazy;
  pl->params.ninv     = make_mulc(pl->n_inv, pl->q, lazy);
  pl->params.ninv_w1  = make_mulc(nttm_mulmod(pl->n_inv % pl->q, w_inv_1 % pl->q, pl->q), pl->q, lazy);
  {
    /* FP64 path: multipliers centred to (-q/2, q/2) like the twiddle tables (k_build_fd) */
    const double   qd = (double)pl->q;
    const uint64_t ua = pl->n_inv % pl->q, ub = nttm_mulmod(pl->n_inv % pl->q, w_inv_1 % pl->q, pl->q);
    const double   a = ua > (pl->q >> 1) ? -(double)(pl->q - ua) : (double)ua;
    const double   b = ub > (pl->q >> 1) ? -(double)(pl->q - ub) : (double)ub;
    pl->params.ninv_fd[0]    = a;
    pl->params.ninv_fd[1]    = a / qd;
    pl->params.ninv_w1_fd[0] = b;
    pl->params.ninv_w1_fd[1] = b / qd;
  }
  if(lazy && ntt_cuda_plan_inverse_bounds(&pl->params)) return cuda_error("inverse bounds");
  return NTT_B200_SUCCESS;
}

/* upload one reference table, build the kernel tables, verify the caller's companion table */
static int adopt_table(ntt_b200_plan_t *pl, const uint64_t *w, const uint64_t *w_con, uint64_t **d_w, void **wu,
                       void **qq, uint64_t **d_con, const char *name)
{
  const size_t bytes = (size_t)pl->N * 8;
  if(ntt_cuda_malloc(pl->device, (void **)d_w, bytes)) return cuda_error("table alloc");
  if(ntt_cuda_h2d(pl->device, *d_w, w, bytes, NULL)) return cuda_error("table upload");
  int rc = build_direction(pl, *d_w, wu, qq, d_con);
  if(!rc) rc = (wu == &pl->d_fwd_wu) ? build_ctables(pl, *wu, *qq, &pl->d_fwd_ct_wu, &pl->d_fwd_ct_qq)
                                     : build_ctables(pl, *wu, *qq, &pl->d_inv_ct_wu, &pl->d_inv_ct_qq);
  if(!rc) rc = (wu == &pl->d_fwd_wu) ? build_fd(pl, *d_w, &pl->d_fwd_fd, &pl->d_fwd_ct_fd)
                                     : build_fd(pl, *d_w, &pl->d_inv_fd, &pl->d_inv_ct_fd);
  if(rc) return rc;
  if(w_con) {
    uint64_t *chk = malloc(bytes);
    if(!chk) return set_error("out of host memory%s", NULL);
    rc = ntt_cuda_d2h(pl->device, chk, *d_con, bytes, NULL) || ntt_cuda_sync(pl->device, NULL);
    if(rc) {
      free(chk);
      return cuda_error("table download");
    }
    /* entry 0 (w = 1) is never used by the transform (src/ntt_reference.c:19-22 starts at m = 1) */
    const int same = memcmp(chk + 1, w_con + 1, bytes - 8) == 0;
    free(chk);
    if(!same) return set_error("%s does not equal floor(w * 2^64 / q): tables inconsistent with q", name);
  } else if(ntt_cuda_sync(pl->device, NULL)) {
    return cuda_error("table build");
  }
  return NTT_B200_SUCCESS;
}

int ntt_b200_plan_create(ntt_b200_plan_t **plan, int device, uint64_t N, uint64_t q, const uint64_t *w,
                         const uint64_t *w_con, const uint64_t *w_inv, const uint64_t *w_inv_con, uint64_t n_inv,
                         uint64_t n_inv_con)
{
  if(!plan) return set_error("plan pointer is NULL%s", NULL);
  *plan = NULL;
  if(check_shape(N, q, device)) return NTT_B200_ERROR;
  if(!w && !w_inv) return set_error("at least one of w / w_inv must be given%s", NULL);
  ntt_b200_plan_t *pl = plan_alloc(device, N, q);
  if(!pl) return set_error("out of host memory%s", NULL);
  int rc = NTT_B200_SUCCESS;
  if(w) {
    rc = adopt_table(pl, w, w_con, &pl->d_w, &pl->d_fwd_wu, &pl->d_fwd_qq, &pl->d_w_con, "w_con");
    pl->has_fwd = (rc == NTT_B200_SUCCESS);
  }
  if(!rc && w_inv) {
    if(nttm_mulmod(n_inv % q, N % q, q) != 1 % q) {
      rc = set_error("n_inv is not N^-1 mod q%s", NULL);
    } else if(n_inv_con != nttm_shoup(n_inv, q, 64)) {
      rc = set_error("n_inv_con does not equal floor(n_inv * 2^64 / q)%s", NULL);
    } else {
      rc = adopt_table(pl, w_inv, w_inv_con, &pl->d_w_inv, &pl->d_inv_wu, &pl->d_inv_qq, &pl->d_w_inv_con,
                       "w_inv_con");
    }
    if(!rc) {
      pl->n_inv     = n_inv;
      pl->n_inv_con = n_inv_con;
      rc            = finish_inverse_constants(pl, w_inv[1]);
      pl->has_inv   = (rc == NTT_B200_SUCCESS);
    }
  }
  if(rc) {
    plan_free(pl);
    return rc;
  }
  publish_tables(pl);
  *plan = pl;
  return NTT_B200_SUCCESS;
}

int ntt_b200_plan_create_psi(ntt_b200_plan_t **plan, int device, uint64_t N, uint64_t q, uint64_t psi)
{
  if(!plan) return set_error("plan pointer is NULL%s", NULL);
  *plan = NULL;
  if(check_shape(N, q, device)) return NTT_B200_ERROR;
  psi %= q;
  if(nttm_powmod(psi, N, q) != q - 1) return set_error("psi is not a primitive 2N-th root of unity mod q%s", NULL);
  ntt_b200_plan_t *pl = plan_alloc(device, N, q);
  if(!pl) return set_error("out of host memory%s", NULL);
  const uint64_t psi_inv = nttm_powmod(psi, 2 * N - 1, q);
  /* N^-1 = ((q+1)/2)^log2(N): valid for any odd q */
  pl->n_inv     = nttm_powmod((q + 1) / 2, pl->logn, q);
  pl->n_inv_con = nttm_shoup(pl->n_inv, q, 64);
  const size_t bytes = (size_t)N * 8;
  int          rc    = NTT_B200_SUCCESS;
  if(ntt_cuda_malloc(device, (void **)&pl->d_w, bytes) || ntt_cuda_malloc(device, (void **)&pl->d_w_inv, bytes)) {
    rc = cuda_error("table alloc");
  } else if(ntt_cuda_gen_root_table(device, pl->d_w, psi, N, q, NULL) ||
            ntt_cuda_gen_root_table(device, pl->d_w_inv, psi_inv, N, q, NULL)) {
    rc = cuda_error("table generation");
  }
  if(!rc) rc = build_direction(pl, pl->d_w, &pl->d_fwd_wu, &pl->d_fwd_qq, &pl->d_w_con);
  if(!rc) rc = build_direction(pl, pl->d_w_inv, &pl->d_inv_wu, &pl->d_inv_qq, &pl->d_w_inv_con);
  if(!rc) rc = build_ctables(pl, pl->d_fwd_wu, pl->d_fwd_qq, &pl->d_fwd_ct_wu, &pl->d_fwd_ct_qq);
  if(!rc) rc = build_ctables(pl, pl->d_inv_wu, pl->d_inv_qq, &pl->d_inv_ct_wu, &pl->d_inv_ct_qq);
  if(!rc) rc = build_fd(pl, pl->d_w, &pl->d_fwd_fd, &pl->d_fwd_ct_fd);
  if(!rc) rc = build_fd(pl, pl->d_w_inv, &pl->d_inv_fd, &pl->d_inv_ct_fd);
  if(!rc && ntt_cuda_sync(device, NULL)) rc = cuda_error("table generation");
  /* w_inv[1] = psi_inv^(N/2) */
  if(!rc) rc = finish_inverse_constants(pl, nttm_powmod(psi_inv, N / 2, q));
  if(rc) {
    plan_free(pl);
    return rc;
  }
  pl->has_fwd = pl->has_inv = 1;
  publish_tables(pl);
  *plan = pl;
  return NTT_B200_SUCCESS;
}

int ntt_b200_plan_destroy(ntt_b200_plan_t *plan)
{
  plan_free(plan);
  return NTT_B200_SUCCESS;
}

uint64_t ntt_b200_plan_n(const ntt_b200_plan_t *plan) { return plan ? plan->N : 0; }
uint64_t ntt_b200_plan_q(const ntt_b200_plan_t *plan) { return plan ? plan->q : 0; }
int      ntt_b200_plan_device(const ntt_b200_plan_t *plan) { return plan ? plan->device : -1; }
int      ntt_b200_plan_is_lazy(const ntt_b200_plan_t *plan) { return plan ? (int)plan->params.lazy : 0; }

int ntt_b200_plan_describe(const ntt_b200_plan_t *plan, int inverse, char *buf, size_t n, int *launches)
{
  if(!plan) return set_error("plan is NULL%s", NULL);
  if(ntt_cuda_describe(&plan->params, inverse, buf, n, launches)) return cuda_error("describe");
  return NTT_B200_SUCCESS;
}

int ntt_b200_plan_export_tables(const ntt_b200_plan_t *plan, uint64_t *w, uint64_t *w_con, uint64_t *w_inv,
                                uint64_t *w_inv_con, uint64_t *n_inv, uint64_t *n_inv_con)
{
  if(!plan) return set_error("plan is NULL%s", NULL);
  const size_t bytes = (size_t)plan->N * 8;
  struct {
    uint64_t *      dst;
    const uint64_t *src;
  } jobs[4] = {{w, plan->d_w}, {w_con, plan->d_w_con}, {w_inv, plan->d_w_inv}, {w_inv_con, plan->d_w_inv_con}};
  for(int i = 0; i < 4; i++) {
    if(!jobs[i].dst) continue;
    if(!jobs[i].src) return set_error("plan does not hold the requested table%s", NULL);
    if(ntt_cuda_d2h(plan->device, jobs[i].dst, jobs[i].src, bytes, NULL)) return cuda_error("table download");
  }
  if(ntt_cuda_sync(plan->device, NULL)) return cuda_error("table download");
  if(n_inv) *n_inv = plan->n_inv;
  if(n_inv_con) *n_inv_con = plan->n_inv_con;
  return NTT_B200_SUCCESS;
}

/* ---- device-resident batches --------------------------------------------------------------------- */

static int check_batch(const ntt_b200_plan_t *plan, const void *d_a, int need_inv)
{
  if(!plan) return set_error("plan is NULL%s", NULL);
  if(!d_a) return set_error("data pointer is NULL%s", NULL);
  if(((uintptr_t)d_a & 15) != 0) return set_error("device data must be 16-byte aligned%s", NULL);
  if(need_inv ? !plan->has_inv : !plan->has_fwd)
    return set_error("plan was created without the %s tables", need_inv ? "inverse" : "forward");
  return NTT_B200_SUCCESS;
}

int ntt_b200_plan_set_inverse_scale(ntt_b200_plan_t *plan, uint64_t scale)
{
  if(!plan || !plan->has_inv) return set_error("plan has no inverse tables%s", NULL);
  plan->n_inv     = scale % plan->q;
  plan->n_inv_con = nttm_shoup(plan->n_inv, plan->q, 64);
  return finish_inverse_constants(plan, plan->w_inv_1);
}

int ntt_b200_fwd_tail_block(const ntt_b200_plan_t *plan, uint64_t *d_block, uint32_t log2_parts, uint32_t block,
                            void *stream)
{
  if(check_batch(plan, d_block, 0)) return NTT_B200_ERROR;
  if(ntt_cuda_tail(plan->device, &plan->params, d_block, log2_parts, block, 0, stream)) return cuda_error("forward tail");
  return NTT_B200_SUCCESS;
}

int ntt_b200_inv_tail_block(const ntt_b200_plan_t *plan, uint64_t *d_block, uint32_t log2_parts, uint32_t block,
                            void *stream)
{
  if(check_batch(plan, d_block, 1)) return NTT_B200_ERROR;
  if(ntt_cuda_tail(plan->device, &plan->params, d_block, log2_parts, block, 1, stream)) return cuda_error("inverse tail");
  return NTT_B200_SUCCESS;
}

int ntt_b200_fwd_tail_gather_batch(const ntt_b200_plan_t *plan, uint64_t *const *peer_slices, uint64_t *d_block,
                                   uint32_t log2_parts, uint32_t rank, size_t batch, void *stream)
{
  if(check_batch(plan, d_block, 0)) return NTT_B200_ERROR;
  if(!peer_slices) return set_error("peer slice table is NULL%s", NULL);
  if(ntt_cuda_tail_peer(plan->device, &plan->params, peer_slices, d_block, log2_parts, rank, batch, 0, stream))
    return cuda_error("forward tail (peer gather)");
  return NTT_B200_SUCCESS;
}

int ntt_b200_inv_tail_scatter_batch(const ntt_b200_plan_t *plan, uint64_t *const *peer_slices, uint64_t *d_block,
                                    uint32_t log2_parts, uint32_t rank, size_t batch, void *stream)
{
  if(check_batch(plan, d_block, 1)) return NTT_B200_ERROR;
  if(!peer_slices) return set_error("peer slice table is NULL%s", NULL);
  if(ntt_cuda_tail_peer(plan->device, &plan->params, peer_slices, d_block, log2_parts, rank, batch, 1, stream))
    return cuda_error("inverse tail (peer scatter)");
  return NTT_B200_SUCCESS;
}

int ntt_b200_fwd_tail_gather(const ntt_b200_plan_t *plan, uint64_t *const *peer_slices, uint64_t *d_block,
                             uint32_t log2_parts, uint32_t rank, void *stream)
{
  return ntt_b200_fwd_tail_gather_batch(plan, peer_slices, d_block, log2_parts, rank, 1, stream);
}

int ntt_b200_inv_tail_scatter(const ntt_b200_plan_t *plan, uint64_t *const *peer_slices, uint64_t *d_block,
                              uint32_t log2_parts, uint32_t rank, void *stream)
{
  return ntt_b200_inv_tail_scatter_batch(plan, peer_slices, d_block, log2_parts, rank, 1, stream);
}

int ntt_b200_peer_barrier(int device, void *const *peer_flags, void *my_flags, uint32_t rank, uint32_t world,
                          uint32_t epoch, int *d_timed_out, void *stream)
{
  if(!peer_flags || !my_flags || !d_timed_out) return set_error("peer barrier: NULL argument%s", NULL);
  if(ntt_cuda_peer_barrier(device, peer_flags, my_flags, rank, world, epoch, d_timed_out, stream))
    return cuda_error("peer barrier");
  return NTT_B200_SUCCESS;
}

int ntt_b200_ipc_export(int device, void *d_ptr, void *handle64)
{
  if(!d_ptr || !handle64) return set_error("ipc export: NULL argument%s", NULL);
  if(ntt_cuda_ipc_export(device, d_ptr, handle64)) return cuda_error("cudaIpcGetMemHandle");
  return NTT_B200_SUCCESS;
}

int ntt_b200_ipc_open(int device, const void *handle64, void **d_ptr)
{
  if(!d_ptr || !handle64) return set_error("ipc open: NULL argument%s", NULL);
  if(ntt_cuda_ipc_open(device, handle64, d_ptr)) return cuda_error("cudaIpcOpenMemHandle");
  return NTT_B200_SUCCESS;
}

int ntt_b200_ipc_close(int device, void *d_ptr)
{
  if(!d_ptr) return NTT_B200_SUCCESS;
  if(ntt_cuda_ipc_close(device, d_ptr)) return cuda_error("cudaIpcCloseMemHandle");
  return NTT_B200_SUCCESS;
}

int ntt_b200_fwd_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream)
{
  if(check_batch(plan, d_a, 0)) return NTT_B200_ERROR;
  if(ntt_cuda_forward(plan->device, &plan->params, d_a, batch, stream)) return cuda_error("forward NTT");
  return NTT_B200_SUCCESS;
}

int ntt_b200_fwd_lazy_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream)
{
  /* fwd_ntt_ref_harvey_lazy (src/ntt_reference.c:11-31) leaves its output in [0,4q) and lets the caller reduce.
   * Here the FP64 ring kernel skips its final sign correction and returns values in [0,2q); the kernels that have
   * no cheaper lazy form return the canonical residue, which satisfies the same contract. */
  if(check_batch(plan, d_a, 0)) return NTT_B200_ERROR;
  ntt_cuda_fwd_opts_t o = {NULL, 0, 1};
  if(ntt_cuda_forward_ex(plan->device, &plan->params, d_a, batch, stream, &o, NULL)) return cuda_error("forward NTT");
  return NTT_B200_SUCCESS;
}

/* ---- order-agnostic ("unordered") entry points ------------------------------------------------------------
 * The reference's fwd_ntt_radix4_avx512_ifma_unordered (include/ntt_avx512_ifma.h:88) leaves its output in the lane
 * order of its last SIMD stage and so saves a final in-register transpose; consumers that only multiply pointwise
 * do not care, and tests/test_correctness.c:179-209 (fix_a_order) restores the order for the comparison.  The
 * contract here is the same: the output order of fwd_unordered is implementation-defined, inv_unordered accepts
 * exactly that order, pointwise products commute with it, and ntt_b200_unordered_index tells where position i of
 * the unordered output sits in fwd_ntt_ref_harvey's output.  On this architecture the in-place register / shared
 * memory network leaves the reference's bit-reversed order at no cost (there is no scatter to skip: every pass
 * writes back where it read), so the permutation is the identity today; callers written against this contract keep
 * working if a later kernel (e.g. a multi-GPU transform that skips its final exchange) chooses another order. */
int ntt_b200_fwd_unordered_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream)
{
  return ntt_b200_fwd_batch(plan, d_a, batch, stream);
}
int ntt_b200_inv_unordered_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream)
{
  return ntt_b200_inv_batch(plan, d_a, batch, stream);
}
uint64_t ntt_b200_unordered_index(const ntt_b200_plan_t *plan, uint64_t i)
{
  return (plan && i < plan->N) ? i : (uint64_t)-1;
}

int ntt_b200_inv_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream)
{
  if(check_batch(plan, d_a, 1)) return NTT_B200_ERROR;
  if(ntt_cuda_inverse(plan->device, &plan->params, d_a, batch, stream)) return cuda_error("inverse NTT");
  return NTT_B200_SUCCESS;
}

static int rns_apply(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *d_a, size_t batch_per_limb,
                     void *stream, int inverse)
{
  if(!plans || limbs == 0) return set_error("no plans given%s", NULL);
  for(size_t l = 0; l < limbs; l++) {
    if(!plans[l] || plans[l]->N != plans[0]->N || plans[l]->device != plans[0]->device)
      return set_error("RNS plans must share N and device%s", NULL);
  }
  const ntt_cuda_params_t **plist = malloc(limbs * sizeof(*plist));
  if(!plist) return set_error("out of host memory%s", NULL);
  int rc = NTT_B200_SUCCESS;
  for(size_t l = 0; l < limbs && !rc; l++) {
    if(inverse ? !plans[l]->has_inv : !plans[l]->has_fwd)
      rc = set_error("an RNS plan was created without the %s tables", inverse ? "inverse" : "forward");
    plist[l] = &plans[l]->params;
  }
  if(!rc && (((uintptr_t)d_a & 15) != 0)) rc = set_error("device data must be 16-byte aligned%s", NULL);
  if(!rc && ntt_cuda_rns(plans[0]->device, plist, limbs, d_a, batch_per_limb, inverse, stream))
    rc = cuda_error("RNS transform");
  free(plist);
  return rc;
}

int ntt_b200_fwd_rns(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *d_a, size_t batch_per_limb,
                     void *stream)
{
  return rns_apply(plans, limbs, d_a, batch_per_limb, stream, 0);
}
int ntt_b200_inv_rns(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *d_a, size_t batch_per_limb,
                     void *stream)
{
  return rns_apply(plans, limbs, d_a, batch_per_limb, stream, 1);
}

int ntt_b200_pointwise_mul_batch(const ntt_b200_plan_t *plan, uint64_t *d_c, const uint64_t *d_a,
                                 const uint64_t *d_b, size_t batch, void *stream)
{
  if(!plan || !d_c || !d_a || !d_b) return set_error("NULL argument%s", NULL);
  if(ntt_cuda_pointwise(plan->device, &plan->params, d_c, d_a, d_b, batch * (size_t)plan->N, stream))
    return cuda_error("pointwise product");
  return NTT_B200_SUCCESS;
}

int ntt_b200_negacyclic_mul_batch(const ntt_b200_plan_t *plan, uint64_t *d_c, uint64_t *d_a, uint64_t *d_b,
                                  size_t batch, void *stream)
{
  if(check_batch(plan, d_a, 1) || check_batch(plan, d_b, 0) || check_batch(plan, d_c, 0)) return NTT_B200_ERROR;
  {
    /* N = 2^13 in the FP64 range: one kernel, both operands in shared memory (csrc/ntt_polymul_fp.cuh) */
    int done = 0;
    if(ntt_cuda_polymul(plan->device, &plan->params, d_c, d_a, d_b, batch, stream, &done))
      return cuda_error("negacyclic multiply");
    if(done) return NTT_B200_SUCCESS;
  }
  if(d_a == d_b) {
    /* squaring: one forward transform, the NTT-domain square, the inverse (the fused second forward would read
     * its own half-transformed buffer as the other operand) */
    if(ntt_b200_fwd_batch(plan, d_a, batch, stream)) return NTT_B200_ERROR;
    if(ntt_b200_pointwise_mul_batch(plan, d_a, d_a, d_a, batch, stream)) return NTT_B200_ERROR;
    if(ntt_b200_inv_batch(plan, d_a, batch, stream)) return NTT_B200_ERROR;
    if(d_c != d_a && ntt_cuda_d2d(plan->device, d_c, d_a, batch * (size_t)plan->N * 8, stream))
      return cuda_error("result copy");
    return NTT_B200_SUCCESS;
  }
  /* the product lands in `prod` (the operand transformed second): prefer the one that aliases d_c */
  uint64_t *first = d_a, *prod = d_b;
  if(d_c == d_a) {
    first = d_b;
    prod  = d_a;
  }
  if(ntt_b200_fwd_batch(plan, first, batch, stream)) return NTT_B200_ERROR;
  int fused = 0;
  if(ntt_cuda_forward_mul(plan->device, &plan->params, prod, first, batch, stream, &fused))
    return cuda_error("forward NTT with fused product");
  if(!fused && ntt_b200_pointwise_mul_batch(plan, prod, prod, first, batch, stream)) return NTT_B200_ERROR;
  if(ntt_b200_inv_batch(plan, prod, batch, stream)) return NTT_B200_ERROR;
  if(prod != d_c && ntt_cuda_d2d(plan->device, d_c, prod, batch * (size_t)plan->N * 8, stream))
    return cuda_error("result copy");
  return NTT_B200_SUCCESS;
}

/* ---- host-resident batches ------------------------------------------------------------------------- */

static int pipe_prepare(ntt_b200_plan_t *pl)
{
  if(pl->pipe_polys) return NTT_B200_SUCCESS;
  const size_t poly_bytes = (size_t)pl->N * 8;
  const char *ed = getenv("NTT_B200_PIPE_DEPTH"), *eb = getenv("NTT_B200_PIPE_MIB");
  if(ed && atoi(ed) >= 1 && atoi(ed) <= HOST_PIPE_DEPTH_MAX) g_pipe_depth = atoi(ed);
  if(eb && atoi(eb) >= 1 && atoi(eb) <= 1024) g_pipe_bytes = (size_t)atoi(eb) << 20;
  pl->pipe_depth     = g_pipe_depth;
  size_t       polys = g_pipe_bytes / poly_bytes;
  if(polys < 1) polys = 1;
  for(int i = 0; i < pl->pipe_depth; i++) {
    if(ntt_cuda_stream_create(pl->device, &pl->pipe_stream[i])) return cuda_error("stream create");
    if(ntt_cuda_malloc(pl->device, (void **)&pl->pipe_buf[i], polys * poly_bytes)) return cuda_error("staging alloc");
  }
  pl->pipe_polys = polys;
  return NTT_B200_SUCCESS;
}

/* The three things the host-buffer pipeline can do to a chunk between its two copies */
enum { HOST_FWD = 0, HOST_INV = 1, HOST_FWD_MUL_INV = 2 };

static int host_chunk_work(const ntt_b200_plan_t *pl, uint64_t *d, size_t take, int mode, const uint64_t *d_m, void *st)
{
  if(mode == HOST_FWD) return ntt_cuda_forward(pl->device, &pl->params, d, take, st);
  if(mode == HOST_INV) return ntt_cuda_inverse(pl->device, &pl->params, d, take, st);
  /* forward, NTT-domain product with ONE resident polynomial (fused into the transform where the kernel can), inverse */
  if(d_m) {
    ntt_cuda_fwd_opts_t o = {d_m, 1, 0};
    int                 fused = 0;
    if(ntt_cuda_forward_ex(pl->device, &pl->params, d, take, st, &o, &fused)) return -1;
    if(!fused && ntt_cuda_pointwise_bcast(pl->device, &pl->params, d, d, d_m, take * (size_t)pl->N, st)) return -1;
  } else if(ntt_cuda_forward(pl->device, &pl->params, d, take, st)) {
    return -1;
  }
  return ntt_cuda_inverse(pl->device, &pl->params, d, take, st);
}

/* H2D -> transform(s) -> D2H in chunks; chunk i runs on stream i mod depth with its own staging buffer.  Nothing
 * waits on the host inside the loop: a slot's buffer is reused by the NEXT chunk on the SAME stream, so stream
 * order already puts that chunk's H2D behind the previous D2H, while the copies and kernels of the other slots
 * overlap with it (full-duplex PCIe: H2D of chunk i+1, kernels of chunk i, D2H of chunk i-1). */
static int host_apply(const ntt_b200_plan_t *plan, uint64_t *h_a, size_t batch, int mode, const uint64_t *d_m)
{
  if(!plan) return set_error("plan is NULL%s", NULL);
  if(!h_a) return set_error("data pointer is NULL%s", NULL);
  if((mode != HOST_INV && !plan->has_fwd) || (mode != HOST_FWD && !plan->has_inv))
    return set_error("plan was created without the %s tables", mode == HOST_INV || plan->has_fwd ? "inverse" : "forward");
  if(batch == 0) return NTT_B200_SUCCESS;
  ntt_b200_plan_t *pl = (ntt_b200_plan_t *)plan; /* pipeline state is internal and lock-protected */
  pthread_mutex_lock(&pl->pipe_lock);
  int rc = pipe_prepare(pl);
  const size_t poly_words = (size_t)pl->N;
  size_t       done = 0;
  int          slot = 0;
  size_t       ramp = pl->pipe_polys >> 3; /* first chunk: an eighth of a full one */
  if(ramp < 1) ramp = 1;
  while(!rc && done < batch) {
    /* Chunk sizes ramp up 1/8, 1/4, 1/2, 1 at the head and halve again over the tail: the first H2D copy has no D2H to
     * overlap with and the last D2H copy no H2D, so both should be short (with equal chunks of 32 MiB the two
     * unpaired copies cost 1/16 of a 512 MiB batch each). */
    const size_t left = batch - done;
    size_t       take = ramp < pl->pipe_polys ? ramp : pl->pipe_polys;
    if(left <= 2 * take && left > 2 * (pl->pipe_polys >> 3) + 1) take = (left + 1) / 2;
    if(take > left) take = left;
    if(ramp < pl->pipe_polys) ramp <<= 1; /* (stops doubling at the full size: it must never wrap to 0) */
    const size_t bytes = take * poly_words * 8;
    void *       st    = pl->pipe_stream[slot];
    uint64_t *   d     = pl->pipe_buf[slot];
    if(ntt_cuda_h2d(pl->device, d, h_a + done * poly_words, bytes, st)) rc = cuda_error("H2D copy");
    if(!rc && host_chunk_work(pl, d, take, mode, d_m, st)) rc = cuda_error("transform");
    if(!rc && ntt_cuda_d2h(pl->device, h_a + done * poly_words, d, bytes, st)) rc = cuda_error("D2H copy");
    done += take;
    slot = (slot + 1) % pl->pipe_depth;
  }
  for(int i = 0; i < pl->pipe_depth; i++) {
    if(pl->pipe_stream[i] && ntt_cuda_sync(pl->device, pl->pipe_stream[i]) && !rc) rc = cuda_error("pipeline sync");
  }
  pthread_mutex_unlock(&pl->pipe_lock);
  return rc;
}

int ntt_b200_fwd_batch_host(const ntt_b200_plan_t *plan, uint64_t *h_a, size_t batch)
{
  return host_apply(plan, h_a, batch, HOST_FWD, NULL);
}
int ntt_b200_inv_batch_host(const ntt_b200_plan_t *plan, uint64_t *h_a, size_t batch)
{
  return host_apply(plan, h_a, batch, HOST_INV, NULL);
}
int ntt_b200_fwd_mul_inv_batch_host(const ntt_b200_plan_t *plan, uint64_t *h_a, const uint64_t *d_m, size_t batch)
{
  if(d_m && ((uintptr_t)d_m & 15) != 0) return set_error("device data must be 16-byte aligned%s", NULL);
  return host_apply(plan, h_a, batch, HOST_FWD_MUL_INV, d_m);
}

int ntt_b200_fwd_mul_inv_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, const uint64_t *d_m, size_t batch,
                               void *stream)
{
  if(check_batch(plan, d_a, 1) || check_batch(plan, d_a, 0)) return NTT_B200_ERROR;
  if(d_m && ((uintptr_t)d_m & 15) != 0) return set_error("device data must be 16-byte aligned%s", NULL);
  if(host_chunk_work(plan, d_a, batch, HOST_FWD_MUL_INV, d_m, stream)) return cuda_error("forward-multiply-inverse");
  return NTT_B200_SUCCESS;
}

int ntt_b200_host_alloc(void **ptr, size_t bytes)
{
  if(!ptr) return set_error("NULL argument%s", NULL);
  if(ntt_cuda_host_alloc(ptr, bytes)) return cuda_error("pinned alloc");
  return NTT_B200_SUCCESS;
}
int ntt_b200_host_free(void *ptr)
{
  if(ptr && ntt_cuda_host_free(ptr)) return cuda_error("pinned free");
  return NTT_B200_SUCCESS;
}
int ntt_b200_device_alloc(int device, void **d_ptr, size_t bytes)
{
  if(!d_ptr) return set_error("NULL argument%s", NULL);
  if(ntt_cuda_malloc(device, d_ptr, bytes)) return cuda_error("device alloc");
  return NTT_B200_SUCCESS;
}
int ntt_b200_device_free(int device, void *d_ptr)
{
  if(d_ptr && ntt_cuda_free(device, d_ptr)) return cuda_error("device free");
  return NTT_B200_SUCCESS;
}
int ntt_b200_memcpy_h2d(int device, void *d_dst, const void *h_src, size_t bytes)
{
  if(ntt_cuda_h2d(device, d_dst, h_src, bytes, NULL) || ntt_cuda_sync(device, NULL)) return cuda_error("H2D copy");
  return NTT_B200_SUCCESS;
}
int ntt_b200_memcpy_d2h(int device, void *h_dst, const void *d_src, size_t bytes)
{
  if(ntt_cuda_d2h(device, h_dst, d_src, bytes, NULL) || ntt_cuda_sync(device, NULL)) return cuda_error("D2H copy");
  return NTT_B200_SUCCESS;
}
int ntt_b200_device_sync(int device)
{
  if(ntt_cuda_sync(device, NULL)) return cuda_error("device sync");
  return NTT_B200_SUCCESS;
}

/* ---- host-side table builders (pre_compute.h:16-83 replacements) ------------------------------------- */

uint64_t ntt_b200_bit_rev_idx(uint64_t idx, uint64_t width) { return nttm_bitrev(idx, (unsigned)width); }

int ntt_b200_calc_w(uint64_t *out, uint64_t root, uint64_t N, uint64_t q)
{
  if(!out || N < 1 || (N & (N - 1)) != 0 || q < 2) return set_error("bad arguments to calc_w%s", NULL);
  const unsigned m   = nttm_log2(N);
  uint64_t       pwr = 1 % q;
  for(uint64_t i = 0; i < N; i++) {
    out[nttm_bitrev(i, m)] = pwr;
    pwr                    = nttm_mulmod(pwr, root % q, q);
  }
  return NTT_B200_SUCCESS;
}

int ntt_b200_calc_w_con(uint64_t *out, const uint64_t *w, uint64_t N, uint64_t q, uint64_t word_size)
{
  if(!out || !w || q < 2 || word_size > 64) return set_error("bad arguments to calc_w_con%s", NULL);
  for(uint64_t i = 0; i < N; i++) out[i] = nttm_shoup(w[i], q, (unsigned)word_size);
  return NTT_B200_SUCCESS;
}

uint64_t ntt_b200_calc_ninv_con(uint64_t n_inv, uint64_t q, uint64_t word_size)
{
  return nttm_shoup(n_inv, q, (unsigned)word_size);
}

uint64_t ntt_b200_pow_mod(uint64_t a, uint64_t e, uint64_t q) { return nttm_powmod(a, e, q); }
uint64_t ntt_b200_inv_mod(uint64_t a, uint64_t q) { return nttm_invmod_prime(a, q); }
int      ntt_b200_is_prime(uint64_t n) { return nttm_is_prime(n); }
uint64_t ntt_b200_min_primitive_root(uint64_t N, uint64_t q) { return nttm_min_primitive_root_2n(N, q); }
