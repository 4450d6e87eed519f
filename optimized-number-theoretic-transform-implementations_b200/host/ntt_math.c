/* host/ntt_math.c -- see ntt_math.h */
#include "ntt_math.h"

typedef unsigned __int128 u128;

uint64_t nttm_mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }

uint64_t nttm_powmod(uint64_t a, uint64_t e, uint64_t q)
{
  uint64_t acc = 1 % q;
  for(a %= q; e; e >>= 1) {
    if(e & 1) acc = nttm_mulmod(acc, a, q);
    a = nttm_mulmod(a, a, q);
  }
  return acc;
}

uint64_t nttm_invmod_prime(uint64_t a, uint64_t q) { return nttm_powmod(a, q - 2, q); }

unsigned nttm_log2(uint64_t n) { return 63u - (unsigned)__builtin_clzll(n); }
unsigned nttm_bitlen(uint64_t n) { return n ? 64u - (unsigned)__builtin_clzll(n) : 0u; }

uint64_t nttm_bitrev(uint64_t idx, unsigned width)
{
  /* reverse all 64 bits by swapping halves, then keep the top `width` bits */
  uint64_t v = idx;
  v          = ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
  v          = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
  v          = ((v >> 4) & 0x0f0f0f0f0f0f0f0fULL) | ((v & 0x0f0f0f0f0f0f0f0fULL) << 4);
  v          = __builtin_bswap64(v);
  return width ? v >> (64 - width) : 0;
}

uint64_t nttm_shoup(uint64_t w, uint64_t q, unsigned word_bits) { return (uint64_t)((((u128)w) << word_bits) / q); }

int nttm_is_prime(uint64_t n)
{
  /* deterministic Miller-Rabin for 64-bit n with the first twelve primes as witnesses */
  static const uint64_t wit[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if(n < 2) return 0;
  for(int i = 0; i < 12; i++) {
    if(n == wit[i]) return 1;
    if(n % wit[i] == 0) return 0;
  }
  uint64_t odd = n - 1;
  unsigned twos = 0;
  while((odd & 1) == 0) {
    odd >>= 1;
    twos++;
  }
  for(int i = 0; i < 12; i++) {
    uint64_t x = nttm_powmod(wit[i], odd, n);
    if(x == 1 || x == n - 1) continue;
    unsigned r = 1;
    for(; r < twos; r++) {
      x = nttm_mulmod(x, x, n);
      if(x == n - 1) break;
    }
    if(r == twos) return 0;
  }
  return 1;
}

uint64_t nttm_min_primitive_root_2n(uint64_t N, uint64_t q)
{
  const uint64_t order = 2 * N;
  if(q < 3 || (q - 1) % order != 0) return 0;
  /* any x with x^((q-1)/2) == -1 is a non-residue; its (q-1)/2N-th power then has exact order 2N */
  uint64_t gen = 0;
  for(uint64_t x = 2; x < q && !gen; x++) {
    const uint64_t cand = nttm_powmod(x, (q - 1) / order, q);
    if(nttm_powmod(cand, N, q) == q - 1) gen = cand;
  }
  if(!gen) return 0;
  /* all primitive 2N-th roots are the odd powers of gen; return the numerically smallest */
  const uint64_t step = nttm_mulmod(gen, gen, q);
  uint64_t       cur = gen, best = gen;
  for(uint64_t k = 1; k < N; k++) {
    cur = nttm_mulmod(cur, step, q);
    if(cur < best) best = cur;
  }
  return best;
}
