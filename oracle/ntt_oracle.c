/*
 * oracle/ntt_oracle.c -- TEST INFRASTRUCTURE ONLY (see ntt_oracle.h for the rules and the parity status).
 *
 * CPU restatement of the reference hot path.  The arithmetic is stated with explicit 128-bit products
 * so that every intermediate (including the lazy, not-fully-reduced values) is bit-identical to what
 * the reference's inline primitives produce:
 *
 *   shoup_lazy()        <- fast_mul_mod_q2            include/internal/fast_mul_operators.h:49-54
 *   fold_2q()/fold_q()  <- reduce_4q_to_2q / 2q_to_q  include/internal/fast_mul_operators.h:15-28
 *   forward butterfly   <- harvey_fwd_butterfly       include/internal/fast_mul_operators.h:72-81
 *   inverse butterfly   <- harvey_bkw_butterfly       include/internal/fast_mul_operators.h:83-92
 *   last inverse stage  <- harvey_bkw_butterfly_final include/internal/fast_mul_operators.h:94-106
 */
#include "ntt_oracle.h"

typedef unsigned __int128 u128;

/* ---- modular primitives -------------------------------------------------------------------- */

/* r = w*t - floor(w_con*t / 2^64)*q  (mod 2^64); lies in [0,2q) for every t < 2^64 */
static inline uint64_t shoup_lazy(uint64_t w, uint64_t w_con, uint64_t t, uint64_t q)
{
  const uint64_t quot = (uint64_t)(((u128)w_con * t) >> 64);
  return w * t - quot * q;
}

static inline uint64_t fold_2q(uint64_t v, uint64_t q) { return v >= 2 * q ? v - 2 * q : v; }
static inline uint64_t fold_q(uint64_t v, uint64_t q) { return v >= q ? v - q : v; }

static inline unsigned log2_u64(uint64_t n)
{
  unsigned l = 0;
  while((n >> l) > 1) l++;
  return l;
}

/* ---- tables -------------------------------------------------------------------------------- */

uint64_t oracle_bitrev(uint64_t idx, unsigned width)
{
  uint64_t r = 0;
  for(unsigned b = 0; b < width; b++) {
    r = (r << 1) | ((idx >> b) & 1);
  }
  return r;
}

void oracle_root_table(uint64_t *tbl, uint64_t root, uint64_t N, uint64_t q)
{
  const unsigned m   = log2_u64(N);
  uint64_t       pwr = 1;
  for(uint64_t i = 0; i < N; i++) {
    tbl[oracle_bitrev(i, m)] = pwr;
    pwr                      = (uint64_t)(((u128)pwr * root) % q);
  }
}

uint64_t oracle_shoup_companion(uint64_t v, uint64_t q, unsigned word_bits)
{
  return (uint64_t)((((u128)v) << word_bits) / q);
}

void oracle_shoup_table(uint64_t *con, const uint64_t *tbl, uint64_t N, uint64_t q, unsigned word_bits)
{
  for(uint64_t i = 0; i < N; i++) con[i] = oracle_shoup_companion(tbl[i], q, word_bits);
}

/* ---- forward transform ---------------------------------------------------------------------- */

/* Stage s (s = 0 .. log2N-1) has 2^s blocks of 2t points, t = N / 2^(s+1); block i pairs positions
 * (base+j, base+j+t) with base = 2*t*i and multiplies the upper one by w[2^s + i].
 * Same schedule as the triple loop at src/ntt_reference.c:19-30, walked by butterfly number. */
void oracle_fwd_lazy(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con)
{
  const unsigned m  = log2_u64(N);
  const uint64_t q2 = 2 * q;
  for(unsigned s = 0; s < m; s++) {
    const unsigned lt = m - 1 - s; /* log2 t */
    for(uint64_t b = 0; b < N / 2; b++) {
      const uint64_t i  = b >> lt;
      const uint64_t j  = b & ((1ULL << lt) - 1);
      const uint64_t lo = (i << (lt + 1)) + j;
      const uint64_t hi = lo + (1ULL << lt);
      const uint64_t x  = fold_2q(a[lo], q);
      const uint64_t p  = shoup_lazy(w[(1ULL << s) + i], w_con[(1ULL << s) + i], a[hi], q);
      a[lo]             = x + p;
      a[hi]             = x - p + q2;
    }
  }
}

void oracle_fwd(uint64_t *a, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con)
{
  oracle_fwd_lazy(a, N, q, w, w_con);
  for(uint64_t i = 0; i < N; i++) a[i] = fold_q(fold_2q(a[i], q), q);
}

void oracle_fwd_dbl(uint64_t *a1, uint64_t *a2, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con)
{
  /* the reference interleaves the two polynomials butterfly by butterfly (ntt_reference.c:83-86);
   * the two transforms are independent, so running them back to back gives the same bytes */
  oracle_fwd(a1, N, q, w, w_con);
  oracle_fwd(a2, N, q, w, w_con);
}

/* ---- inverse transform ---------------------------------------------------------------------- */

void oracle_inv(uint64_t *a, uint64_t N, uint64_t q, uint64_t n_inv, uint64_t n_inv_con, unsigned word_bits,
                const uint64_t *w, const uint64_t *w_con)
{
  const unsigned m  = log2_u64(N);
  const uint64_t q2 = 2 * q;

  /* Gentleman-Sande stages with 2^s blocks for s = m-1 .. 1, t = 1 .. N/4 (ntt_reference.c:43-53) */
  for(unsigned s = m - 1; s >= 1; s--) {
    const unsigned lt = m - 1 - s;
    for(uint64_t b = 0; b < N / 2; b++) {
      const uint64_t i  = b >> lt;
      const uint64_t j  = b & ((1ULL << lt) - 1);
      const uint64_t lo = (i << (lt + 1)) + j;
      const uint64_t hi = lo + (1ULL << lt);
      const uint64_t x = a[lo], y = a[hi];
      a[lo] = fold_2q(x + y, q);
      a[hi] = shoup_lazy(w[(1ULL << s) + i], w_con[(1ULL << s) + i], x - y + q2, q);
    }
  }

  /* last stage (one block, t = N/2) with N^-1 folded in: the twiddle becomes n_inv * w[1], kept in
   * [0,2q), and its companion is recomputed by a 128-bit divide (ntt_reference.c:57-61) */
  const uint64_t sw     = shoup_lazy(n_inv, n_inv_con, w[1], q);
  const uint64_t sw_con = (uint64_t)((((u128)sw) << word_bits) / q);
  const uint64_t half   = N / 2;
  for(uint64_t j = 0; j < half; j++) {
    const uint64_t x = a[j], y = a[j + half];
    a[j]        = fold_q(shoup_lazy(n_inv, n_inv_con, x + y, q), q);
    a[j + half] = fold_q(shoup_lazy(sw, sw_con, x - y + q2, q), q);
  }
}

/* ---- definitions used as independent cross-checks ---------------------------------------------- */

uint64_t oracle_powmod(uint64_t a, uint64_t e, uint64_t q)
{
  uint64_t r = 1 % q;
  a %= q;
  while(e) {
    if(e & 1) r = (uint64_t)(((u128)r * a) % q);
    a = (uint64_t)(((u128)a * a) % q);
    e >>= 1;
  }
  return r;
}

uint64_t oracle_invmod(uint64_t a, uint64_t q) { return oracle_powmod(a, q - 2, q); }

void oracle_fwd_definition(uint64_t *out, const uint64_t *a, uint64_t N, uint64_t q, uint64_t psi)
{
  const unsigned m = log2_u64(N);
  for(uint64_t i = 0; i < N; i++) {
    const uint64_t x   = oracle_powmod(psi, 2 * oracle_bitrev(i, m) + 1, q);
    uint64_t       acc = 0, xp = 1;
    for(uint64_t j = 0; j < N; j++) {
      acc = (uint64_t)((acc + (u128)(a[j] % q) * xp) % q);
      xp  = (uint64_t)(((u128)xp * x) % q);
    }
    out[i] = acc;
  }
}

void oracle_negacyclic_mul(uint64_t *c, const uint64_t *a, const uint64_t *b, uint64_t N, uint64_t q)
{
  for(uint64_t k = 0; k < N; k++) c[k] = 0;
  for(uint64_t i = 0; i < N; i++) {
    for(uint64_t j = 0; j < N; j++) {
      const uint64_t p = (uint64_t)(((u128)a[i] * b[j]) % q);
      const uint64_t k = i + j;
      if(k < N) {
        c[k] = (uint64_t)(((u128)c[k] + p) % q);
      } else {
        c[k - N] = (uint64_t)(((u128)c[k - N] + q - p) % q);
      }
    }
  }
}

void oracle_pointwise_mul(uint64_t *c, const uint64_t *a, const uint64_t *b, uint64_t N, uint64_t q)
{
  for(uint64_t i = 0; i < N; i++) c[i] = (uint64_t)(((u128)a[i] * b[i]) % q);
}

int oracle_is_prime(uint64_t n)
{
  static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if(n < 2) return 0;
  for(size_t k = 0; k < sizeof(bases) / sizeof(bases[0]); k++) {
    if(n % bases[k] == 0) return n == bases[k];
  }
  uint64_t d = n - 1;
  unsigned r = 0;
  while(!(d & 1)) {
    d >>= 1;
    r++;
  }
  for(size_t k = 0; k < sizeof(bases) / sizeof(bases[0]); k++) {
    uint64_t x = oracle_powmod(bases[k], d, n);
    if(x == 1 || x == n - 1) continue;
    int composite = 1;
    for(unsigned i = 1; i < r; i++) {
      x = (uint64_t)(((u128)x * x) % n);
      if(x == n - 1) {
        composite = 0;
        break;
      }
    }
    if(composite) return 0;
  }
  return 1;
}

uint64_t oracle_min_primitive_root_2n(uint64_t N, uint64_t q)
{
  /* find one primitive 2N-th root g (g^N == -1), then the minimum over its odd powers */
  const uint64_t two_n = 2 * N;
  if((q - 1) % two_n) return 0;
  uint64_t g = 0;
  for(uint64_t x = 2; x < q; x++) {
    const uint64_t c = oracle_powmod(x, (q - 1) / two_n, q);
    if(oracle_powmod(c, N, q) == q - 1) {
      g = c;
      break;
    }
  }
  if(!g) return 0;
  const uint64_t g2   = (uint64_t)(((u128)g * g) % q);
  uint64_t       best = g, cur = g;
  for(uint64_t k = 1; k < N; k++) {
    cur = (uint64_t)(((u128)cur * g2) % q);
    if(cur < best) best = cur;
  }
  return best;
}

void oracle_fill_uniform(uint64_t *a, size_t n, uint64_t q, uint64_t seed)
{
  uint64_t s = seed;
  for(size_t i = 0; i < n; i++) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z          = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z          = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    a[i]       = (z ^ (z >> 31)) % q;
  }
}

uint64_t oracle_fnv1a64(const uint64_t *a, size_t n)
{
  uint64_t h = 0xcbf29ce484222325ULL;
  for(size_t i = 0; i < n; i++) {
    for(unsigned b = 0; b < 8; b++) {
      h ^= (a[i] >> (8 * b)) & 0xff;
      h *= 0x100000001b3ULL;
    }
  }
  return h;
}

/* ---- whole batches on several host threads (the every-row soak and bench.py's in-run check) ----------------
 * rows x N contiguous polynomials, each transformed by oracle_fwd / oracle_inv above; `threads` pthreads take
 * rows round-robin.  Same arithmetic, only the loop over independent polynomials is parallel. */
#include <pthread.h>
typedef struct {
  uint64_t *a;
  size_t rows, first, step;
  uint64_t N, q, n_inv, n_inv_con;
  const uint64_t *w, *w_con;
  int inverse;
} batch_job_t;

static void *batch_worker(void *arg)
{
  const batch_job_t *j = (const batch_job_t *)arg;
  for(size_t r = j->first; r < j->rows; r += j->step) {
    if(j->inverse) oracle_inv(j->a + r * j->N, j->N, j->q, j->n_inv, j->n_inv_con, 64, j->w, j->w_con);
    else oracle_fwd(j->a + r * j->N, j->N, j->q, j->w, j->w_con);
  }
  return NULL;
}

static void batch_run(batch_job_t base, unsigned threads)
{
  if(threads < 1) threads = 1;
  if(threads > 256) threads = 256;
  pthread_t   tid[256];
  batch_job_t job[256];
  unsigned    started = 0;
  for(unsigned t = 0; t < threads; t++) {
    job[t]       = base;
    job[t].first = t;
    job[t].step  = threads;
    if(pthread_create(&tid[t], NULL, batch_worker, &job[t]) != 0) break;
    started++;
  }
  if(started < threads) { /* could not start them all: finish the missing residues here */
    for(unsigned t = started; t < threads; t++) {
      job[t]       = base;
      job[t].first = t;
      job[t].step  = threads;
      batch_worker(&job[t]);
    }
  }
  for(unsigned t = 0; t < started; t++) pthread_join(tid[t], NULL);
}

void oracle_fwd_batch(uint64_t *a, size_t rows, uint64_t N, uint64_t q, const uint64_t *w, const uint64_t *w_con,
                      unsigned threads)
{
  batch_job_t b = {a, rows, 0, 1, N, q, 0, 0, w, w_con, 0};
  batch_run(b, threads);
}

void oracle_inv_batch(uint64_t *a, size_t rows, uint64_t N, uint64_t q, uint64_t n_inv, uint64_t n_inv_con,
                      const uint64_t *w, const uint64_t *w_con, unsigned threads)
{
  batch_job_t b = {a, rows, 0, 1, N, q, n_inv, n_inv_con, w, w_con, 1};
  batch_run(b, threads);
}
