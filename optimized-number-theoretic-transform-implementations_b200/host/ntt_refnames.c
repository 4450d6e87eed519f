/*
 * host/ntt_refnames.c -- libntt_b200_dropin.so: the reference's OWN symbol names and signatures.
 *
 * Linking the reference's tests/test_correctness.c or tests/bench.c against this library instead of
 * src/ntt_reference.c routes fwd_ntt_ref_harvey_lazy / inv_ntt_ref_harvey / fwd_ntt_ref_harvey_lazy_dbl
 * (include/ntt_reference.h:13-49) to the GPU.  mul_op_t is re-declared here with the reference's layout
 * (include/internal/fast_mul_operators.h:10-13: two __uint128_t, 32 bytes, passed by value in memory on
 * SysV x86-64) so no reference header is needed to build.
 *
 * The reference functions return void and have no error channel; a failure (no GPU, inconsistent tables)
 * is reported on stderr and aborts, it is never papered over with a CPU computation.
 */
#include <stdio.h>
#include <stdlib.h>

#include "../../include/ntt_b200.h"

typedef struct mul_op_s {
  __uint128_t op;
  __uint128_t con;
} mul_op_t;

#define EXPORT __attribute__((visibility("default")))

static void must(int rc, const char *fn)
{
  if(rc != NTT_B200_SUCCESS) {
    fprintf(stderr, "ntt_b200 drop-in: %s failed: %s\n", fn, ntt_b200_last_error());
    abort();
  }
}

EXPORT void fwd_ntt_ref_harvey_lazy(uint64_t a[], uint64_t N, uint64_t q, const uint64_t w[], const uint64_t w_con[])
{
  must(ntt_b200_fwd_ntt_ref_harvey_lazy(a, N, q, w, w_con), "fwd_ntt_ref_harvey_lazy");
}

EXPORT void inv_ntt_ref_harvey(uint64_t a[], uint64_t N, uint64_t q, mul_op_t n_inv, uint64_t word_size,
                               const uint64_t w[], const uint64_t w_con[])
{
  must(ntt_b200_inv_ntt_ref_harvey(a, N, q, (uint64_t)n_inv.op, (uint64_t)n_inv.con, word_size, w, w_con),
       "inv_ntt_ref_harvey");
}

EXPORT void fwd_ntt_ref_harvey_lazy_dbl(uint64_t a1[], uint64_t a2[], uint64_t N, uint64_t q, const uint64_t w[],
                                        const uint64_t w_con[])
{
  must(ntt_b200_fwd_ntt_ref_harvey_dbl(a1, a2, N, q, w, w_con), "fwd_ntt_ref_harvey_lazy_dbl");
}
