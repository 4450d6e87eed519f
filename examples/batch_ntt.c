/*
 * examples/batch_ntt.c -- the batch C-ABI from plain C, as a caller of the reference would use it.
 *
 * The reference's own loop (tests/test_correctness.c:24-34: build the tables, fwd_ntt_ref_harvey on one
 * polynomial, inv_ntt_ref_harvey, memcmp) becomes: build a plan once, transform `batch` polynomials per call.
 *
 *   gcc -std=c11 -O2 -Iinclude examples/batch_ntt.c -Loptimized-number-theoretic-transform-implementations_b200 \
 *       -lntt_b200 -Wl,-rpath,$PWD/optimized-number-theoretic-transform-implementations_b200 -o batch_ntt
 *   ./batch_ntt [log2N] [batch]        exit code 0 = round trip and table check passed, 2 = no usable GPU
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ntt_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    if((call) != NTT_B200_SUCCESS) {                                             \
      fprintf(stderr, "%s failed: %s\n", #call, ntt_b200_last_error());          \
      return 1;                                                                  \
    }                                                                            \
  } while(0)

int main(int argc, char **argv)
{
  const uint64_t logn  = argc > 1 ? strtoull(argv[1], NULL, 10) : 14;
  const size_t   batch = argc > 2 ? strtoull(argv[2], NULL, 10) : 64;
  const uint64_t N = 1ull << logn, q = 0x1fffffc800001ull; /* BASELINE config 2: 49-bit prime, q = 1 (mod 2^23) */
  if(logn < 1 || logn > 22) {
    fprintf(stderr, "log2N must be 1..22 for this modulus\n");
    return 1;
  }
  printf("%s\n", ntt_b200_version());
  if(ntt_b200_device_count() < 1) {
    fprintf(stderr, "no CUDA device: %s (there is no CPU fallback)\n", ntt_b200_last_error());
    return 2;
  }

  /* psi: the smallest primitive 2N-th root of unity, as tests/test_cases.h fixes one per case */
  const uint64_t psi = ntt_b200_min_primitive_root(N, q);
  if(psi == 0 || ntt_b200_pow_mod(psi, N, q) != q - 1) {
    fprintf(stderr, "no primitive 2N-th root found\n");
    return 1;
  }

  ntt_b200_plan_t *plan = NULL;
  CHECK(ntt_b200_plan_create_psi(&plan, 0, N, q, psi));

  /* the plan's device-generated tables are the reference's: w[bitrev(i)] = psi^i (pre_compute.h:38-58) */
  uint64_t *w = malloc(N * 8), *w_con = malloc(N * 8), *w_ref = malloc(N * 8);
  uint64_t  n_inv = 0, n_inv_con = 0;
  CHECK(ntt_b200_plan_export_tables(plan, w, w_con, NULL, NULL, &n_inv, &n_inv_con));
  CHECK(ntt_b200_calc_w(w_ref, psi, N, q));
  if(memcmp(w, w_ref, N * 8) != 0) {
    fprintf(stderr, "device table differs from calc_w\n");
    return 1;
  }

  /* batch x N coefficients in pinned host memory; the host entry points pipeline H2D, kernels and D2H */
  uint64_t *a = NULL, *orig = malloc(batch * N * 8);
  CHECK(ntt_b200_host_alloc((void **)&a, batch * N * 8));
  uint64_t s = 88172645463325252ull;
  for(size_t i = 0; i < batch * N; i++) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17; /* xorshift64 */
    orig[i] = a[i] = s % q;
  }
  CHECK(ntt_b200_fwd_batch_host(plan, a, batch)); /* = fwd_ntt_ref_harvey on every polynomial */
  size_t changed = 0;
  for(size_t i = 0; i < batch * N; i++) changed += a[i] != orig[i];
  CHECK(ntt_b200_inv_batch_host(plan, a, batch)); /* = inv_ntt_ref_harvey */
  if(memcmp(a, orig, batch * N * 8) != 0) {
    fprintf(stderr, "inverse(forward(a)) != a\n");
    return 1;
  }
  printf("N = 2^%" PRIu64 ", q = 0x%" PRIx64 ", psi = %" PRIu64 ", batch %zu: tables match calc_w, forward changed %zu words, "
         "round trip exact\n", logn, q, psi, batch, changed);

  ntt_b200_host_free(a);
  ntt_b200_plan_destroy(plan);
  free(orig), free(w), free(w_con), free(w_ref);
  return 0;
}
