/*
 * examples/multi_gpu_ntt.c -- the multi-GPU form of the batch C-ABI from plain C: a device list in, the batch
 * sharded over it, results equal to running every polynomial through fwd_ntt_ref_harvey / inv_ntt_ref_harvey
 * (include/ntt_reference.h:19-39), timed with the host clock like the reference's MEASURE
 * (tests/measurements.h:57-75).
 *
 *   gcc -std=c11 -O2 -Iinclude examples/multi_gpu_ntt.c -Loptimized-number-theoretic-transform-implementations_b200 \
 *       -lntt_b200 -Wl,-rpath,$PWD/optimized-number-theoretic-transform-implementations_b200 -o multi_gpu_ntt
 *   ./multi_gpu_ntt [log2N] [batch per GPU] [GPUs]     exit 0 = checks passed, 2 = no usable GPU
 *
 * Two things are exercised:
 *   1. host-resident data: ntt_b200_multi_fwd_mul_inv_batch_host -- one host thread per device, every polynomial
 *      crosses PCIe once per direction for a forward transform, an NTT-domain product and an inverse transform;
 *   2. device-resident data: ntt_b200_multi_fwd_batch / _inv_batch on per-device buffers (the throughput path).
 * The same work on 1 device and on all devices must give identical bytes.
 */
#define _POSIX_C_SOURCE 200809L
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ntt_b200.h"

#define CHECK(call)                                                                                              \
  do {                                                                                                           \
    if((call) != NTT_B200_SUCCESS) {                                                                             \
      fprintf(stderr, "%s failed: %s / %s\n", #call, ntt_b200_multi_last_error(), ntt_b200_last_error());        \
      return 1;                                                                                                  \
    }                                                                                                            \
  } while(0)

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
  const uint64_t logn    = argc > 1 ? strtoull(argv[1], NULL, 10) : 14;
  const size_t   per_gpu = argc > 2 ? strtoull(argv[2], NULL, 10) : 256;
  int            ndev    = ntt_b200_device_count();
  if(argc > 3 && atoi(argv[3]) > 0 && atoi(argv[3]) < ndev) ndev = atoi(argv[3]);
  const uint64_t N = 1ull << logn, q = 0x1fffffc800001ull;
  if(ndev < 1) {
    fprintf(stderr, "no CUDA device: %s (there is no CPU fallback)\n", ntt_b200_last_error());
    return 2;
  }
  if(logn < 1 || logn > 22) return 1;
  const uint64_t psi   = ntt_b200_min_primitive_root(N, q);
  const size_t   batch = per_gpu * (size_t)ndev;

  ntt_b200_multi_t *all = NULL, *one = NULL;
  CHECK(ntt_b200_multi_create(&all, NULL, ndev, N, q, psi));
  CHECK(ntt_b200_multi_create(&one, NULL, 1, N, q, psi));

  /* inputs in pinned host memory + a multiplier polynomial (NTT domain), one copy per device */
  uint64_t *a = NULL, *b = NULL, *m = malloc(N * 8);
  CHECK(ntt_b200_host_alloc((void **)&a, batch * N * 8));
  CHECK(ntt_b200_host_alloc((void **)&b, batch * N * 8));
  uint64_t s = 88172645463325252ull;
  for(size_t i = 0; i < batch * N; i++) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    a[i] = b[i] = s % q;
  }
  for(size_t i = 0; i < N; i++) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    m[i] = s % q;
  }
  uint64_t *d_m[64] = {0};
  for(int i = 0; i < ndev; i++) {
    CHECK(ntt_b200_device_alloc(i, (void **)&d_m[i], N * 8));
    CHECK(ntt_b200_memcpy_h2d(i, d_m[i], m, N * 8));
  }

  /* 1. host-resident: all devices vs. one device, identical bytes */
  CHECK(ntt_b200_multi_fwd_mul_inv_batch_host(all, a, (const uint64_t *const *)d_m, batch)); /* warm-up + result */
  CHECK(ntt_b200_multi_fwd_mul_inv_batch_host(one, b, (const uint64_t *const *)d_m, batch));
  if(memcmp(a, b, batch * N * 8) != 0) {
    fprintf(stderr, "sharded result differs from the single-device result\n");
    return 1;
  }
  double t0 = now_s();
  const int reps = 3;
  for(int r = 0; r < reps; r++) CHECK(ntt_b200_multi_fwd_mul_inv_batch_host(all, a, (const uint64_t *const *)d_m, batch));
  const double host_s = (now_s() - t0) / reps;

  /* 2. device-resident: per-device shards, forward + inverse must reproduce the input */
  uint64_t *d_a[64] = {0};
  size_t    cnt[64] = {0};
  for(int i = 0; i < ndev; i++) {
    size_t first;
    ntt_b200_shard_range(batch, ndev, i, &first, &cnt[i]);
    CHECK(ntt_b200_device_alloc(i, (void **)&d_a[i], cnt[i] * N * 8));
    CHECK(ntt_b200_memcpy_h2d(i, d_a[i], b + first * N, cnt[i] * N * 8));
  }
  for(int r = 0; r < 3; r++) { /* warm-up */
    CHECK(ntt_b200_multi_fwd_batch(all, d_a, cnt, NULL));
    CHECK(ntt_b200_multi_inv_batch(all, d_a, cnt, NULL));
  }
  CHECK(ntt_b200_multi_sync(all));
  t0 = now_s();
  const int dreps = 20;
  for(int r = 0; r < dreps; r++) {
    CHECK(ntt_b200_multi_fwd_batch(all, d_a, cnt, NULL));
    CHECK(ntt_b200_multi_inv_batch(all, d_a, cnt, NULL));
  }
  CHECK(ntt_b200_multi_sync(all));
  const double dev_s = (now_s() - t0) / dreps;
  for(int i = 0; i < ndev; i++) {
    size_t first;
    ntt_b200_shard_range(batch, ndev, i, &first, NULL);
    CHECK(ntt_b200_memcpy_d2h(i, a + first * N, d_a[i], cnt[i] * N * 8));
  }
  if(memcmp(a, b, batch * N * 8) != 0) {
    fprintf(stderr, "device-resident round trip is not the identity\n");
    return 1;
  }
  printf("N = 2^%" PRIu64 ", %d GPU(s), batch %zu: sharded == single device, round trip exact; "
         "host buffers %.3f M NTT/s (fwd+mul+inv per call), device-resident %.2f M NTT/s\n",
         logn, ndev, batch, 2.0 * (double)batch / host_s * 1e-6, 2.0 * (double)batch / dev_s * 1e-6);

  for(int i = 0; i < ndev; i++) {
    ntt_b200_device_free(i, d_a[i]);
    ntt_b200_device_free(i, d_m[i]);
  }
  ntt_b200_host_free(a);
  ntt_b200_host_free(b);
  free(m);
  ntt_b200_multi_destroy(all);
  ntt_b200_multi_destroy(one);
  return 0;
}
