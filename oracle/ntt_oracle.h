/*
 * oracle/ntt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference's negacyclic NTT hot path
 * (IBM/optimized-number-theoretic-transform-implementations).  It exists so that the CUDA product
 * path can be checked bit-for-bit.  Nothing under optimized-number-theoretic-transform-implementations_b200/
 * may include, link or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against the reference's own
 * compiled sources (oracle/_ref/libntt_ref.so, built from /root/reference by oracle/Makefile) on all
 * 19 fixture cases of tests/test_cases.h:145-208, and against the committed fixtures in tests/golden/
 * that were generated from that same reference build (oracle/gen_golden.py).
 *
 * Every function cites the reference lines it restates (paths relative to the reference root).
 */
#ifndef NTT_ORACLE_H
#define NTT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/internal/pre_compute.h:16-26 (bit_rev_idx) */
uint64_t oracle_bitrev(uint64_t idx, unsigned width);

/* include/internal/pre_compute.h:38-66 (calc_w / calc_w_inv + bit_rev): tbl[bitrev_m(i)] = root^i mod q */
void oracle_root_table(uint64_t *tbl, uint64_t root, uint64_t N, uint64_t q);

/* include/internal/pre_compute.h:68-83 (calc_w_con, calc_ninv_con): floor(v * 2^word_bits / q) */
uint64_t oracle_shoup_companion(uint64_t v, uint64_t q, unsigned word_bits);
void     oracle_shoup_table(uint64_t *con, const uint64_t *tbl, uint64_t N, uint64_t q, unsigned word_bits);

/* src/ntt_reference.c:11-31 (fwd_ntt_ref_harvey_lazy): output lazily reduced, in [0,4q) */
void oracle_fwd_lazy(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con);

/* include/ntt_reference.h:19-31 (fwd_ntt_ref_harvey): lazy transform + full reduction to [0,q) */
void oracle_fwd(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con);

/* src/ntt_reference.c:33-66 (inv_ntt_ref_harvey); n_inv / n_inv_con are the .op/.con of the reference's
 * mul_op_t argument (include/internal/fast_mul_operators.h:10-13). w / w_con are tables of psi^-1. */
void oracle_inv(uint64_t *a, uint64_t N, uint64_t q, uint64_t n_inv, uint64_t n_inv_con, unsigned word_bits,
                const uint64_t *w, const uint64_t *w_con);

/* src/ntt_reference.c:71-91 + include/ntt_reference.h:51-65 (fwd_ntt_ref_harvey_dbl) */
void oracle_fwd_dbl(uint64_t *a1, uint64_t *a2, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con);

/* oracle_fwd / oracle_inv over `rows` contiguous polynomials, rows dealt round-robin to `threads` pthreads
 * (the arithmetic per polynomial is the single-threaded code above) */
void oracle_fwd_batch(uint64_t *a, size_t rows, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con,
                      unsigned threads);
void oracle_inv_batch(uint64_t *a, size_t rows, uint64_t N, uint64_t q, uint64_t n_inv, uint64_t n_inv_con,
                      const uint64_t *w, const uint64_t *w_con, unsigned threads);

/* Mathematical definition: out[i] = sum_j a[j] * psi^((2*bitrev(i)+1)*j) mod q.
 * O(N^2); an independent cross-check of the butterfly network for small N. */
void oracle_fwd_definition(uint64_t *out, const uint64_t *a, uint64_t N, uint64_t q, uint64_t psi);

/* c = a * b in Z_q[X]/(X^N+1), schoolbook O(N^2) (cross-check for the fused polynomial multiply). */
void oracle_negacyclic_mul(uint64_t *c, const uint64_t *a, const uint64_t *b, uint64_t N, uint64_t q);

/* c[i] = a[i] * b[i] mod q (pointwise product in the NTT domain). */
void oracle_pointwise_mul(uint64_t *c, const uint64_t *a, const uint64_t *b, uint64_t N, uint64_t q);

/* small number theory used to derive fixtures: a^e mod q, a^-1 mod q (q prime), Miller-Rabin,
 * smallest primitive 2N-th root of unity (the rule of the sage script in tests/test_cases.h:113-142) */
uint64_t oracle_powmod(uint64_t a, uint64_t e, uint64_t q);
uint64_t oracle_invmod(uint64_t a, uint64_t q);
int      oracle_is_prime(uint64_t n);
uint64_t oracle_min_primitive_root_2n(uint64_t N, uint64_t q);

/* splitmix64 stream reduced mod q: the generator SURVEY.md Appendix C uses for full-range inputs */
void     oracle_fill_uniform(uint64_t *a, size_t n, uint64_t q, uint64_t seed);
/* FNV-1a-64 over the little-endian bytes of a[0..n) */
uint64_t oracle_fnv1a64(const uint64_t *a, size_t n);

#ifdef __cplusplus
}
#endif
#endif
