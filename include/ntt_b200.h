/*
 * ntt_b200.h -- C-ABI of the B200-native negacyclic NTT (Z_q[X]/(X^N+1)).
 *
 * This is the drop-in boundary for the hot path of IBM/optimized-number-theoretic-transform-implementations:
 * plain pointers and sizes, `int` status codes, no C++/CUDA/torch types.  Every entry point names the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions kept from the reference:
 *   - coefficients are uint64_t, transforms are IN PLACE, the caller owns every buffer
 *     (include/ntt_reference.h:13-39);
 *   - w[] is the bit-reversed table of powers of psi, w[bitrev(i)] = psi^i, and w_con[] its Shoup
 *     companion floor(w*2^64/q) (include/internal/pre_compute.h:38-77);
 *   - forward: natural-order input in [0,4q) -> bit-reversed-order output; inverse: bit-reversed input
 *     in [0,2q) -> natural output in [0,q), scaled by N^-1 (src/ntt_reference.c:11-66);
 *   - status codes NTT_B200_SUCCESS 0 / NTT_B200_ERROR -1 (include/internal/defs.h:20-21).
 *
 * There is no CPU fallback: every call below runs hand-written sm_100a kernels and fails with
 * NTT_B200_ERROR (see ntt_b200_last_error) when no CUDA device is usable.
 */
#ifndef NTT_B200_H
#define NTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTT_B200_SUCCESS 0
#define NTT_B200_ERROR   (-1)

#define NTT_B200_MIN_LOGN 1
#define NTT_B200_MAX_LOGN 24

typedef struct ntt_b200_plan  ntt_b200_plan_t;  /* opaque; one per (device, N, q, psi) */
typedef struct ntt_b200_multi ntt_b200_multi_t; /* opaque; one plan per device of a device list */

/* ---- library ------------------------------------------------------------------------------------ */

/* Human-readable reason of the last NTT_B200_ERROR on the calling thread ("" if none). */
const char *ntt_b200_last_error(void);
/* Number of usable CUDA devices (0 if none / driver missing). */
int ntt_b200_device_count(void);
/* "ntt_b200 <version> sm_100a" */
const char *ntt_b200_version(void);
/*
 * Kernel selection, for benchmarks and A/B parity tests only (every choice is a CUDA path):
 *   "ring" 0/1  persistent TMA ring kernel for chunks of 2^10..2^14 (2^10, 2^11: FP64 range only; default 1; 0 = generic smem kernel)
 *   "fp64" 0/1  run the ring kernel's butterflies on the FP64 pipe when q <= 2^50-2048 (default 1)
 *   "polymul" 0/1  one-kernel negacyclic multiply at N = 2^13 (default 1; 0 = compose it from transforms)
 * The same switches are read from NTT_B200_NO_RING=1 / NTT_B200_NO_FP64=1 / NTT_B200_NO_FUSED_POLYMUL=1 at first use.
 * Further environment switches, read once: NTT_B200_NO_FP_STRIDED=1 (N >= 2^15 in the FP64 range: integer instead of FP64
 * strided passes), NTT_B200_NO_RNS_MULTI=1 (RNS batches: one launch per limb instead of one per kernel),
 * NTT_B200_PIPE_MIB / NTT_B200_PIPE_DEPTH (host-buffer pipeline: chunk size in MiB, default 32, and chunks in flight,
 * default 4).
 */
int ntt_b200_configure(const char *key, int value);

/* ---- plans ---------------------------------------------------------------------------------------- */

/*
 * Build a plan from reference-format tables (host pointers; copied, never retained).
 * Replaces the per-call table arguments of fwd_ntt_ref_harvey / inv_ntt_ref_harvey
 * (include/ntt_reference.h:13-39) and the fixture set-up of tests/test_cases.h:212-238:
 *   w, w_con         forward tables (N words each) -- calc_w + calc_w_con(…, 64)
 *   w_inv, w_inv_con inverse tables (powers of psi^-1); may both be NULL for a forward-only plan
 *   n_inv, n_inv_con N^-1 mod q and its Shoup companion (mul_op_t .op/.con,
 *                    include/internal/fast_mul_operators.h:10-13); ignored when w_inv is NULL
 * Preconditions as in the reference: N = 2^m (1 <= m <= 24), q odd, q < 2^62, tables consistent with q.
 * The *_con tables are validated against w (they must equal floor(w*2^64/q)) and otherwise unused:
 * the device keeps its own twiddle layout.
 */
int ntt_b200_plan_create(ntt_b200_plan_t **plan, int device, uint64_t N, uint64_t q, const uint64_t *w,
                         const uint64_t *w_con, const uint64_t *w_inv, const uint64_t *w_inv_con, uint64_t n_inv,
                         uint64_t n_inv_con);

/*
 * Build a plan from the primitive 2N-th root psi alone; all tables (both directions) and N^-1 are
 * generated ON THE DEVICE.  Replaces calc_w / calc_w_inv / calc_w_con / calc_ninv_con
 * (include/internal/pre_compute.h:38-83).  Fails if psi^N != -1 (mod q).
 */
int ntt_b200_plan_create_psi(ntt_b200_plan_t **plan, int device, uint64_t N, uint64_t q, uint64_t psi);

int ntt_b200_plan_destroy(ntt_b200_plan_t *plan);

/* Plan attributes. */
uint64_t ntt_b200_plan_n(const ntt_b200_plan_t *plan);
uint64_t ntt_b200_plan_q(const ntt_b200_plan_t *plan);
int      ntt_b200_plan_device(const ntt_b200_plan_t *plan);
/* 1 if the plan uses the lazy fast path (q small enough that no per-stage correction is needed),
 * 0 if it uses the general Harvey path (any q < 2^62). */
int ntt_b200_plan_is_lazy(const ntt_b200_plan_t *plan);

/* Names of the kernels one forward (inverse != 0: inverse) transform of this plan launches, e.g.
 * "k_strided<2> + k_ring_fp<14,fwd>", and their number -- so reports never hard-code them. */
int ntt_b200_plan_describe(const ntt_b200_plan_t *plan, int inverse, char *buf, size_t n, int *launches);

/*
 * Copy the plan's tables back to the host in REFERENCE format (each pointer may be NULL to skip):
 * what calc_w / calc_w_con / calc_w_inv would have produced (include/internal/pre_compute.h:38-77).
 * For plans built by ntt_b200_plan_create_psi this returns the device-generated tables.
 */
int ntt_b200_plan_export_tables(const ntt_b200_plan_t *plan, uint64_t *w, uint64_t *w_con, uint64_t *w_inv,
                                uint64_t *w_inv_con, uint64_t *n_inv, uint64_t *n_inv_con);

/* ---- batched transforms, DEVICE-resident data (the measured path) ------------------------------------- */

/*
 * d_a: device pointer on the plan's device, `batch` polynomials of N words each, contiguous,
 * 16-byte aligned.  stream: a cudaStream_t passed as void* (NULL = default stream).  Asynchronous.
 *
 * fwd_batch      = fwd_ntt_ref_harvey      on each polynomial (include/ntt_reference.h:19-31): out in [0,q)
 * fwd_lazy_batch = fwd_ntt_ref_harvey_lazy (src/ntt_reference.c:11-31): contract is "out in [0,4q)", equal to
 *                  fwd_batch after reduce_4q_to_q (tests/test_correctness.c:267-269).  The FP64 ring kernel
 *                  (q <= 2^50-2048, N >= 2^12) skips its final sign correction and returns values in [0,2q);
 *                  the other kernels have no cheaper lazy form and return the canonical residue.
 * inv_batch      = inv_ntt_ref_harvey      (src/ntt_reference.c:33-66): out in [0,q)
 */
int ntt_b200_fwd_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream);
int ntt_b200_fwd_lazy_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream);
int ntt_b200_inv_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream);

/*
 * Order-agnostic variants for consumers that only combine transforms pointwise (the reference's analogue is
 * fwd_ntt_radix4_avx512_ifma_unordered, include/ntt_avx512_ifma.h:88, whose order tests/test_correctness.c:179-209
 * repairs with fix_a_order).  fwd_unordered's output order is implementation-defined; inv_unordered consumes exactly
 * that order; ntt_b200_unordered_index(plan, i) is the index in fwd_ntt_ref_harvey's output of the value at
 * position i of the unordered output ((uint64_t)-1 if i >= N).  Today the permutation is the identity: the
 * in-place network of the sm_100a kernels leaves the reference order for free.
 */
int      ntt_b200_fwd_unordered_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream);
int      ntt_b200_inv_unordered_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, size_t batch, void *stream);
uint64_t ntt_b200_unordered_index(const ntt_b200_plan_t *plan, uint64_t i);

/*
 * RNS form: limb l of every polynomial uses plans[l] (its own q).  d_a holds `limbs` consecutive
 * blocks of `batch_per_limb` polynomials.  All plans must share N and device.  The limbs run concurrently on
 * internal streams that fork from and join back into `stream`.
 */
int ntt_b200_fwd_rns(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *d_a, size_t batch_per_limb,
                     void *stream);
int ntt_b200_inv_rns(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *d_a, size_t batch_per_limb,
                     void *stream);

/*
 * Negacyclic product c = a * b in Z_q[X]/(X^N+1) for `batch` pairs (next row of the scope table:
 * fwd x2, pointwise multiply, inverse).  d_c may alias d_a or d_b; d_a == d_b squares.  d_a and d_b may be
 * overwritten (work space).  At N = 2^13 with q <= 2^50-2048 the whole product is ONE kernel: both operands sit in
 * shared memory, both forward transforms, the product and the inverse run on chip, a and b are read once and c is
 * written once (csrc/ntt_polymul_fp.cuh).  Other sizes compose it from transforms (on the FP64 ring kernel the
 * pointwise product is fused into the second forward transform).
 */
int ntt_b200_negacyclic_mul_batch(const ntt_b200_plan_t *plan, uint64_t *d_c, uint64_t *d_a, uint64_t *d_b,
                                  size_t batch, void *stream);
/* c[i] = a[i]*b[i] mod q over batch*N words (NTT-domain product); inputs in [0,q). */
int ntt_b200_pointwise_mul_batch(const ntt_b200_plan_t *plan, uint64_t *d_c, const uint64_t *d_a,
                                 const uint64_t *d_b, size_t batch, void *stream);
/*
 * a[b] <- INTT( NTT(a[b]) .* m ) for every polynomial b of the batch: the product of a batch with ONE fixed
 * polynomial given in the NTT domain (d_m: N canonical residues on the plan's device; NULL = no product, a plain
 * forward+inverse round trip).  The forward transform, the product and the inverse run back to back on `stream`
 * (the product is fused into the forward kernel where the FP64 ring kernel serves the plan).  Composition of
 * fwd_ntt_ref_harvey (include/ntt_reference.h:19-31), a pointwise product and inv_ntt_ref_harvey
 * (src/ntt_reference.c:33-66).
 */
int ntt_b200_fwd_mul_inv_batch(const ntt_b200_plan_t *plan, uint64_t *d_a, const uint64_t *d_m, size_t batch,
                               void *stream);

/* ---- one transform spread over several GPUs (SURVEY.md section 8e, N = 2^22 config) ----------------------- */

/*
 * Building blocks of the distributed transform over G = 2^log2_parts devices.  With the input held CYCLICALLY
 * (device p owns a[p + G*k]) the first log2(N/G) stages of src/ntt_reference.c:19-30 are exactly a complete
 * forward NTT of size N/G with root psi^G on each slice (use an ordinary plan for (N/G, q, psi^G) and
 * ntt_b200_fwd_batch); after one all-to-all to CONTIGUOUS blocks (device r owns a[r*N/G .. (r+1)*N/G)), the
 * last log2_parts stages are local again:
 *   ntt_b200_fwd_tail_block  runs them on block `block` (N/G words at d_block) with the size-N plan's tables and
 *                            fully reduces: the blocks then hold fwd_ntt_ref_harvey's output, in order.
 *   ntt_b200_inv_tail_block  the mirror image: the first log2_parts stages of inv_ntt_ref_harvey
 *                            (src/ntt_reference.c:43-53) on a contiguous block, output below 2q; exchange back to
 *                            cyclic slices and finish with the size-N/G inverse whose scale was set to N^-1.
 * The exchange itself is the caller's (NCCL all-to-all; see optimized-..._b200/fourstep.py).
 */
int ntt_b200_fwd_tail_block(const ntt_b200_plan_t *plan, uint64_t *d_block, uint32_t log2_parts, uint32_t block,
                            void *stream);
int ntt_b200_inv_tail_block(const ntt_b200_plan_t *plan, uint64_t *d_block, uint32_t log2_parts, uint32_t block,
                            void *stream);
/*
 * The same tail stages FUSED with the exchange over peer memory (NVLink/NVSwitch), no collective call:
 * peer_slices[p] (p < G, host array of device pointers valid on this plan's device -- the local pointer for
 * p == rank, CUDA-IPC mappings for the others) is rank p's slice buffer of N/G words.
 *   ntt_b200_fwd_tail_gather   loads the G members of every butterfly group straight from the peers' slices
 *                              (after their size-N/G forward transforms), runs the last log2_parts stages and
 *                              writes this rank's block, fully reduced -- equals all-to-all + fwd_tail_block.
 *   ntt_b200_inv_tail_scatter  reads this rank's block, runs the first log2_parts inverse stages and stores
 *                              member p of every group into peer p's slice -- equals inv_tail_block + all-to-all.
 *   ntt_b200_peer_barrier      orders the ranks on the GPU timeline (stream-ordered kernel, no host sync):
 *                              peer_flags[k] = rank k's array of 128 uint32 (zeroed once; world <= 32 flags, the
 *                              rest is bookkeeping); epoch grows by one per call, or pass 0 to let the device count
 *                              the calls (the form a captured CUDA graph can replay); *d_timed_out (device int,
 *                              zeroed) is set if a peer never arrived.
 *                              Call it between the local transforms and the gather (forward) and between the
 *                              scatter and the local transforms (inverse).
 *   ntt_b200_ipc_export/open/close  CUDA IPC plumbing for buffers from ntt_b200_device_alloc (64-byte handles).
 */
int ntt_b200_fwd_tail_gather(const ntt_b200_plan_t *plan, uint64_t *const *peer_slices, uint64_t *d_block,
                             uint32_t log2_parts, uint32_t rank, void *stream);
int ntt_b200_inv_tail_scatter(const ntt_b200_plan_t *plan, uint64_t *const *peer_slices, uint64_t *d_block,
                              uint32_t log2_parts, uint32_t rank, void *stream);
/* The same for `batch` transforms at once: every slice buffer holds `batch` slices of N/G words one after the
 * other, d_block `batch` blocks; one launch and one barrier serve the whole batch (a single N = 2^22 exchange is
 * latency-bound). */
int ntt_b200_fwd_tail_gather_batch(const ntt_b200_plan_t *plan, uint64_t *const *peer_slices, uint64_t *d_block,
                                   uint32_t log2_parts, uint32_t rank, size_t batch, void *stream);
int ntt_b200_inv_tail_scatter_batch(const ntt_b200_plan_t *plan, uint64_t *const *peer_slices, uint64_t *d_block,
                                    uint32_t log2_parts, uint32_t rank, size_t batch, void *stream);
int ntt_b200_peer_barrier(int device, void *const *peer_flags, void *my_flags, uint32_t rank, uint32_t world,
                          uint32_t epoch, int *d_timed_out, void *stream);
int ntt_b200_ipc_export(int device, void *d_ptr, void *handle64);
int ntt_b200_ipc_open(int device, const void *handle64, void **d_ptr);
int ntt_b200_ipc_close(int device, void *d_ptr);
/* Replace the inverse transform's scaling constant N^-1 by `scale` (mod q): inv_batch then returns
 * scale * N * (true inverse).  Used by the distributed inverse, whose local size-N/G transform must scale by
 * the global N^-1. */
int ntt_b200_plan_set_inverse_scale(ntt_b200_plan_t *plan, uint64_t scale);

/* ---- batched transforms, HOST-resident data (end-to-end path) ---------------------------------------- */

/*
 * Same transforms on host memory: H2D copy, kernels, D2H copy, synchronous.  h_a may be pageable or
 * pinned (pinned is faster); large batches are pipelined in chunks over two streams.
 */
int ntt_b200_fwd_batch_host(const ntt_b200_plan_t *plan, uint64_t *h_a, size_t batch);
int ntt_b200_inv_batch_host(const ntt_b200_plan_t *plan, uint64_t *h_a, size_t batch);
/* ntt_b200_fwd_mul_inv_batch on host memory: every chunk crosses PCIe once in each direction for a forward AND
 * an inverse transform (d_m stays on the device; NULL = plain round trip). */
int ntt_b200_fwd_mul_inv_batch_host(const ntt_b200_plan_t *plan, uint64_t *h_a, const uint64_t *d_m, size_t batch);

/* ---- several GPUs of one node: batch / RNS-limb sharding, no collective ----- */

/*
 * One plan per device of `devices[0..n_devices)` (NULL = devices 0..n_devices-1), all for the same (N, q, psi);
 * tables are generated on each device.  Polynomials are independent transforms (src/ntt_reference.c:11-66 works on
 * one array), so device i owns the contiguous shard ntt_b200_shard_range(batch, n_devices, i) of a batch.
 */
int ntt_b200_multi_create(ntt_b200_multi_t **multi, const int *devices, int n_devices, uint64_t N, uint64_t q,
                          uint64_t psi);
int ntt_b200_multi_destroy(ntt_b200_multi_t *multi);
int ntt_b200_multi_devices(const ntt_b200_multi_t *multi);
const ntt_b200_plan_t *ntt_b200_multi_plan(const ntt_b200_multi_t *multi, int index);
const char *           ntt_b200_multi_last_error(void);
/* first unit and unit count of part `index` when `batch` units are split into `parts` contiguous shards */
void ntt_b200_shard_range(size_t batch, int parts, int index, size_t *first, size_t *count);
/* device-resident: d_a[i] = batch[i] polynomials on device i; streams[i] may be NULL (or streams NULL).
 * Asynchronous: the launches are issued device after device from the calling thread. */
int ntt_b200_multi_fwd_batch(const ntt_b200_multi_t *multi, uint64_t *const *d_a, const size_t *batch,
                             void *const *streams);
int ntt_b200_multi_inv_batch(const ntt_b200_multi_t *multi, uint64_t *const *d_a, const size_t *batch,
                             void *const *streams);
int ntt_b200_multi_sync(const ntt_b200_multi_t *multi);
/* host-resident: h_a = `batch` polynomials; one host thread per device runs that device's copy/compute pipeline
 * on its contiguous shard.  Synchronous.  d_m[i] (or d_m NULL): the multiplier's copy on device i. */
int ntt_b200_multi_fwd_batch_host(const ntt_b200_multi_t *multi, uint64_t *h_a, size_t batch);
int ntt_b200_multi_inv_batch_host(const ntt_b200_multi_t *multi, uint64_t *h_a, size_t batch);
int ntt_b200_multi_fwd_mul_inv_batch_host(const ntt_b200_multi_t *multi, uint64_t *h_a, const uint64_t *const *d_m,
                                          size_t batch);
/* RNS limbs sharded over devices: plans[l] may live on any device and d_limb[l] points to limb l's
 * batch_per_limb polynomials on THAT device (BASELINE config 3: 48 limbs over 2/4/8 GPUs).  Asynchronous on each
 * device's default stream; follow with ntt_b200_device_sync per device. */
int ntt_b200_fwd_rns_multi(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *const *d_limb, size_t batch_per_limb);
int ntt_b200_inv_rns_multi(ntt_b200_plan_t *const *plans, size_t limbs, uint64_t *const *d_limb, size_t batch_per_limb);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) so C callers need no CUDA headers. */
int ntt_b200_host_alloc(void **ptr, size_t bytes);
int ntt_b200_host_free(void *ptr);
/* Device memory helpers for C callers of the device-resident API. */
int ntt_b200_device_alloc(int device, void **d_ptr, size_t bytes);
int ntt_b200_device_free(int device, void *d_ptr);
int ntt_b200_memcpy_h2d(int device, void *d_dst, const void *h_src, size_t bytes);
int ntt_b200_memcpy_d2h(int device, void *h_dst, const void *d_src, size_t bytes);
int ntt_b200_device_sync(int device);

/* ---- reference-shaped single-polynomial entry points (host pointers) --------------------------------- */

/*
 * Same argument lists as the reference functions they replace; each call looks up (or builds and
 * caches) a plan keyed on (N, q, w[N/2]), copies the polynomial to the device, transforms, copies back.
 *
 *   ntt_b200_fwd_ntt_ref_harvey_lazy  <- fwd_ntt_ref_harvey_lazy  src/ntt_reference.c:11
 *   ntt_b200_fwd_ntt_ref_harvey       <- fwd_ntt_ref_harvey       include/ntt_reference.h:19
 *   ntt_b200_inv_ntt_ref_harvey       <- inv_ntt_ref_harvey       src/ntt_reference.c:33
 *                                        (n_inv/n_inv_con = mul_op_t .op/.con; word_size must be 64)
 *   ntt_b200_fwd_ntt_ref_harvey_dbl   <- fwd_ntt_ref_harvey_dbl   include/ntt_reference.h:51
 *
 * The library libntt_b200_dropin.so additionally exports these under the reference's own symbol names
 * and with its mul_op_t-by-value signature, so the reference's test and bench
 * drivers link against it unchanged.  Return: NTT_B200_SUCCESS / NTT_B200_ERROR.
 */
int ntt_b200_fwd_ntt_ref_harvey_lazy(uint64_t a[], uint64_t N, uint64_t q, const uint64_t w[],
                                     const uint64_t w_con[]);
int ntt_b200_fwd_ntt_ref_harvey(uint64_t a[], uint64_t N, uint64_t q, const uint64_t w[], const uint64_t w_con[]);
int ntt_b200_inv_ntt_ref_harvey(uint64_t a[], uint64_t N, uint64_t q, uint64_t n_inv, uint64_t n_inv_con,
                                uint64_t word_size, const uint64_t w[], const uint64_t w_con[]);
int ntt_b200_fwd_ntt_ref_harvey_dbl(uint64_t a1[], uint64_t a2[], uint64_t N, uint64_t q, const uint64_t w[],
                                    const uint64_t w_con[]);
/* Drop every cached plan created by the reference-shaped entry points. */
void ntt_b200_dropin_reset(void);

/* ---- host-side table builders (C replacements of include/internal/pre_compute.h:16-83) ---------------- */

uint64_t ntt_b200_bit_rev_idx(uint64_t idx, uint64_t width);
/* out[bitrev_m(i)] = root^i mod q, i < N (calc_w / calc_w_inv) */
int ntt_b200_calc_w(uint64_t *out, uint64_t root, uint64_t N, uint64_t q);
/* out[i] = floor(w[i] * 2^word_size / q) (calc_w_con) */
int      ntt_b200_calc_w_con(uint64_t *out, const uint64_t *w, uint64_t N, uint64_t q, uint64_t word_size);
uint64_t ntt_b200_calc_ninv_con(uint64_t n_inv, uint64_t q, uint64_t word_size);
/* number theory helpers for callers that only know (N, q) */
uint64_t ntt_b200_pow_mod(uint64_t a, uint64_t e, uint64_t q);
uint64_t ntt_b200_inv_mod(uint64_t a, uint64_t q); /* q prime */
int      ntt_b200_is_prime(uint64_t n);
/* smallest primitive 2N-th root of unity mod q (the rule of tests/test_cases.h:113-142); 0 if none */
uint64_t ntt_b200_min_primitive_root(uint64_t N, uint64_t q);

#ifdef __cplusplus
}
#endif
#endif /* NTT_B200_H */
