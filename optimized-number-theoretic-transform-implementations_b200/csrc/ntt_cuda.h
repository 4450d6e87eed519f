/*
 * csrc/ntt_cuda.h -- the thin CUDA layer under the C host code.
 *
 * Everything here is extern "C" with POD arguments only, so the host side (the .c files under host/, compiled by gcc as
 * C) never sees CUDA C++ headers.  Device pointers travel as void* / uint64_t*, streams as void*.
 * All functions return 0 on success, -1 on failure with the reason retrievable via ntt_cuda_error().
 */
#ifndef NTT_CUDA_H
#define NTT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTT_MAX_STAGES 24

/* One modular multiplier in the device's "split" form (lazy path), see csrc/ntt_device.cuh:
 *   w  : the multiplier, u = w * 2^32 mod q, wq = floor(w * 2^30 / q), uq = floor(u * 2^30 / q)
 * or, on the exact (Harvey) path: w and c = floor(w * 2^64 / q) in (u0,u1). */
typedef struct ntt_cuda_mulc {
  uint32_t w0, w1, u0, u1;
  uint32_t wq, uq;
} ntt_cuda_mulc_t;

/* Everything a kernel needs to know about one (N, q); passed by value as a kernel parameter. */
typedef struct ntt_cuda_params {
  uint64_t q;
  uint64_t neg2q;   /* 2^64 - 2q (lazy path) */
  uint64_t negq;    /* 2^64 - q */
  uint64_t c10q;    /* 10q: bound of a lazy product, added to keep differences non-negative */
  uint32_t logn;
  uint32_t lazy;    /* 1: lazy split-multiplier path, 0: exact Harvey path */
  uint32_t red_shift; /* reduction: vt = v >> red_shift, Q = hi32(vt * red_mu); shift = max(0, bitlen(q)-9) */
  uint32_t red_mu;
  /* forward twiddles: wu[N] (uint4: w0,w1,u0,u1) and qq[N] (uint2: wq,uq), reference index order */
  const void *fwd_wu;
  const void *fwd_qq;
  const void *inv_wu;
  const void *inv_qq;
  /* the last four stages again, laid out [t][group] (t = 2^u-1+sub, group = 16-coefficient run, N/16 per
   * polynomial) so that the 32 threads of a warp read 32 consecutive entries (ring kernel, pass C) */
  const void *fwd_ct_wu;
  const void *fwd_ct_qq;
  const void *inv_ct_wu;
  const void *inv_ct_qq;
  /* FP64 path (q < 2^49, ntt_ring_fp.cuh): twiddles as (w, RN(w/q)) pairs of doubles, same index order as
   * wu/qq, plus their pass-C re-layout; q and RN(1/q); N^-1 and N^-1*w_inv[1] in the same form */
  uint32_t    fp64; /* 0: not eligible, 1: q <= 2^49-1024, 2: q <= 2^50-2048 (one more fold per pass) */
  uint32_t    pad0;
  double      q_fd, qinv_fd;
  double      ninv_fd[2], ninv_w1_fd[2];
  const void *fwd_fd;
  const void *inv_fd;
  const void *fwd_ct_fd;
  const void *inv_ct_fd;
  /* inverse last stage: N^-1 and N^-1 * w_inv[1] as multipliers */
  ntt_cuda_mulc_t ninv;
  ntt_cuda_mulc_t ninv_w1;
  /* inverse lazy bookkeeping, indexed by global stage s (processed m-1 .. 0):
   * inv_c[s] = B_s * q with B_s the value bound (in units of q) before stage s */
  uint64_t inv_c[NTT_MAX_STAGES];
  uint32_t inv_renorm_mask; /* bit s set: bring values below 2q before running stage s */
} ntt_cuda_params_t;

const char *ntt_cuda_error(void);
/* kernel selection for benchmarks / A-B tests: key "ring" or "fp64", value 0/1 */
int         ntt_cuda_configure(const char *key, int value);
int         ntt_cuda_device_count(void);

int ntt_cuda_malloc(int device, void **d_ptr, size_t bytes);
int ntt_cuda_free(int device, void *d_ptr);
int ntt_cuda_host_alloc(void **h_ptr, size_t bytes);
int ntt_cuda_host_free(void *h_ptr);
int ntt_cuda_h2d(int device, void *d_dst, const void *h_src, size_t bytes, void *stream);
int ntt_cuda_d2h(int device, void *h_dst, const void *d_src, size_t bytes, void *stream);
int ntt_cuda_d2d(int device, void *d_dst, const void *d_src, size_t bytes, void *stream);
int ntt_cuda_sync(int device, void *stream);
int ntt_cuda_stream_create(int device, void **stream);
int ntt_cuda_stream_destroy(int device, void *stream);

/* Fill p->inv_c[] and p->inv_renorm_mask from p->q and p->logn (lazy path bookkeeping). */
int ntt_cuda_plan_inverse_bounds(ntt_cuda_params_t *p);

/* Build device twiddle tables (wu, qq) of N entries from a reference-format table w[] that is already
 * on the device (d_w).  d_con_out (may be NULL) receives floor(w*2^64/q) for validation/export. */
int ntt_cuda_build_tables(int device, const ntt_cuda_params_t *p, const uint64_t *d_w, void *d_wu, void *d_qq,
                          uint64_t *d_con_out, uint64_t N, void *stream);
/* Re-lay the last four stages of (wu, qq) into the [15][N/16] pass-C tables (N >= 16). */
int ntt_cuda_build_ctables(int device, const ntt_cuda_params_t *p, const void *d_wu, const void *d_qq, void *d_ct_wu,
                           void *d_ct_qq, void *stream);
/* FP64 twiddles: d_fd[i] = (w[i], RN(w[i]/q)) from the reference-format table d_w; d_ct_fd (may be NULL) =
 * the last four stages in the [15][N/16] pass-C layout. */
int ntt_cuda_build_fd_tables(int device, const ntt_cuda_params_t *p, const uint64_t *d_w, void *d_fd, void *d_ct_fd,
                             void *stream);
/* Generate the reference-format table d_w[bitrev(i)] = root^i mod q on the device. */
int ntt_cuda_gen_root_table(int device, uint64_t *d_w, uint64_t root, uint64_t N, uint64_t q, void *stream);

/* Transforms over `batch` contiguous polynomials of N = 2^logn words at d_a (in place). */
int ntt_cuda_forward(int device, const ntt_cuda_params_t *p, uint64_t *d_a, size_t batch, void *stream);
int ntt_cuda_inverse(int device, const ntt_cuda_params_t *p, uint64_t *d_a, size_t batch, void *stream);
/* kernel names and launch count of one transform of this plan (for benchmark reports) */
int ntt_cuda_describe(const ntt_cuda_params_t *p, int inverse, char *buf, size_t n, int *launches);
/* RNS batch: limb l uses *plist[l] on d_a + l*batch_per_limb*N; limbs run concurrently on internal streams */
int ntt_cuda_rns(int device, const ntt_cuda_params_t *const *plist, size_t limbs, uint64_t *d_a, size_t batch_per_limb,
                 int inverse, void *stream);
/* Options of the forward transform (FP64 ring kernel; other kernels ignore them and report *fused_out = 0):
 *   d_other         multiply the transform pointwise by this array of canonical residues before it is stored
 *   other_broadcast 0: d_other holds one operand per polynomial (batch*N words); 1: ONE polynomial (N words)
 *   lazy_out        1: the output may stay in [0,2q) (skips the final sign correction); ignored with d_other */
typedef struct ntt_cuda_fwd_opts {
  const uint64_t *d_other;
  int             other_broadcast;
  int             lazy_out;
} ntt_cuda_fwd_opts_t;
int ntt_cuda_forward_ex(int device, const ntt_cuda_params_t *p, uint64_t *d_a, size_t batch, void *stream,
                        const ntt_cuda_fwd_opts_t *opts, int *fused_out);
/* forward transform of d_a followed by d_a .*= d_other; *fused_out says whether the kernel did the product */
int ntt_cuda_forward_mul(int device, const ntt_cuda_params_t *p, uint64_t *d_a, const uint64_t *d_other, size_t batch,
                         void *stream, int *fused_out);
/* whole negacyclic products in one kernel where served (N = 2^13, FP64 range); *done = 0: compose from transforms */
int ntt_cuda_polymul(int device, const ntt_cuda_params_t *p, uint64_t *d_c, uint64_t *d_a, uint64_t *d_b, size_t batch,
                     void *stream, int *done);
/* last `glog` stages on contiguous block `block` of a transform spread over 2^glog devices (see ntt_kernels.cu) */
int ntt_cuda_tail(int device, const ntt_cuda_params_t *p, uint64_t *d_block, uint32_t glog, uint32_t block, int inverse,
                  void *stream);
/* the same tail stages fused with the exchange: forward gathers group members from the peers' slices
 * (peer_slices[p], p < 2^glog, device pointers valid on this device), inverse scatters them back */
int ntt_cuda_tail_peer(int device, const ntt_cuda_params_t *p, uint64_t *const *peer_slices, uint64_t *d_block,
                       uint32_t glog, uint32_t rank, size_t batch, int inverse, void *stream);
/* GPU-timeline barrier over peer memory; peer_flags[k] = rank k's array of `world` uint32 flags */
int ntt_cuda_peer_barrier(int device, void *const *peer_flags, void *my_flags, uint32_t rank, uint32_t world,
                          uint32_t epoch, int *d_timed_out, void *stream);
int ntt_cuda_ipc_export(int device, void *d_ptr, void *handle64);
int ntt_cuda_ipc_open(int device, const void *handle64, void **d_ptr);
int ntt_cuda_ipc_close(int device, void *d_ptr);
/* c = a .* b mod q over n words; inputs < q (any q < 2^62). */
int ntt_cuda_pointwise(int device, const ntt_cuda_params_t *p, uint64_t *d_c, const uint64_t *d_a,
                       const uint64_t *d_b, size_t n, void *stream);

/* c[i] = a[i] * b[i mod N] mod q: ONE polynomial b against every polynomial of a (n = batch*N words) */
int ntt_cuda_pointwise_bcast(int device, const ntt_cuda_params_t *p, uint64_t *d_c, const uint64_t *d_a,
                             const uint64_t *d_b, size_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif
