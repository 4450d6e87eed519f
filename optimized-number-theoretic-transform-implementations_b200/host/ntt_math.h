/*
 * host/ntt_math.h -- host-side modular arithmetic and table builders (plain C, 128-bit intermediates).
 * Product code: used by the plan layer and exported through include/ntt_b200.h as the C replacements of
 * the reference's include/internal/pre_compute.h:16-83.
 */
#ifndef NTT_HOST_MATH_H
#define NTT_HOST_MATH_H
#include <stdint.h>

uint64_t nttm_mulmod(uint64_t a, uint64_t b, uint64_t q);
uint64_t nttm_powmod(uint64_t a, uint64_t e, uint64_t q);
uint64_t nttm_invmod_prime(uint64_t a, uint64_t q);
int      nttm_is_prime(uint64_t n);
uint64_t nttm_bitrev(uint64_t idx, unsigned width);
unsigned nttm_log2(uint64_t n); /* floor(log2 n), n > 0 */
unsigned nttm_bitlen(uint64_t n);
uint64_t nttm_shoup(uint64_t w, uint64_t q, unsigned word_bits); /* floor(w * 2^word_bits / q) */
uint64_t nttm_min_primitive_root_2n(uint64_t N, uint64_t q);
#endif
