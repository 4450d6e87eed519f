"""CPU model of the FP64 butterfly arithmetic of csrc/ntt_ring_fp.cuh.

Python floats are IEEE binary64 and `float(Fraction)` rounds to nearest-even, so fused multiply-add can be
emulated exactly.  The tests below (no GPU needed) pin down the claims the kernel's range schedules rest on:

  * fp_mul / fp_mul_wide / fp_fold return exact integers congruent to the true product / value, with the
    stated magnitude bounds, for operands up to the limits the schedules allow -- and fp_mul really does break
    just beyond 2^51 (which is why fp_mul_wide exists);
  * the worst-case bound propagation of every pass shape (3, 4, 5 stages; forward, inverse, final N^-1 stage;
    q <= 2^49 - 1024 and q <= 2^50 - 2048) stays inside those limits.

The schedule model mirrors fp_network() line by line; if the kernel's schedule changes, change it here too.
"""
import random
from fractions import Fraction as F

import pytest

MAGIC = 6755399441055744.0      # 1.5 * 2^52
T52 = 4503599627370496.0        # 2^52
Q49 = (1 << 49) - 1024 - 1023   # largest odd value the first schedule accepts is 2^49 - 1025; any odd q below works
Q50 = (1 << 50) - 2049


def fma(a, b, c):
    return float(F(a) * F(b) + F(c))


def fp_fold(v, q):
    qd, qinv = float(q), 1.0 / float(q)
    k = fma(v, qinv, MAGIC) - MAGIC
    return fma(-k, qd, v)


def fp_mul(y, w, q):
    qd = float(q)
    winv = float(w) / qd
    cc = fma(y, winv, MAGIC) - MAGIC
    h = y * float(w)
    l = fma(y, float(w), -h)
    d = fma(-cc, qd, h)
    return d + l, cc


def fp_mul_wide(y, w, q):
    qd = float(q)
    winv = float(w) / qd
    arg = y * winv
    ca = (abs(arg) + T52) - T52
    qs = -qd if (arg > 0 or (arg == 0 and str(arg)[0] != "-")) else qd
    h = y * float(w)
    l = fma(y, float(w), -h)
    d = fma(ca, qs, h)
    return d + l, ca


def operands(q, limit_q, n, rng):
    """(y, w) pairs with |y| up to limit_q * q, biased to the extremes."""
    top = int(limit_q * q)
    for i in range(n):
        w = rng.choice([1, 2, q - 1, q - 2, (q + 1) // 2, rng.randrange(1, q)])
        mag = rng.choice([top, top - 1, top - rng.randrange(1, 1 << 20), rng.randrange(0, top + 1), q, q - 1, 0])
        yield float(rng.choice([1, -1]) * mag), w


@pytest.mark.parametrize("q,lim", [(Q49, 3.99), (Q50, 1.99)])
def test_fp_mul_is_exact_below_2_pow_51(q, lim):
    rng = random.Random(1)
    assert lim * q < (1 << 51)
    for y, w in operands(q, lim, 4000, rng):
        t, cc = fp_mul(y, w, q)
        assert cc == int(cc) and t == int(t)
        assert (int(t) - int(y) * w) % q == 0
        assert F(abs(int(t))) <= q * (F(1, 2) + F(abs(int(y)), 1 << 54))


def test_fp_mul_breaks_beyond_2_pow_51_and_wide_does_not():
    """A negative operand whose quotient exceeds 2^51 lands where ulp = 1/2: the magic rounding yields a
    half-integer quotient (this is the failure found on the GPU); the 2^52 form stays exact up to 2^52."""
    q, rng = Q49, random.Random(2)
    broke = 0
    for _ in range(4000):
        w = rng.randrange(q - (1 << 20), q)
        y = -float(rng.randrange(int(4.2 * q), int(7.9 * q)))
        _, cc = fp_mul(y, w, q)
        broke += cc != int(cc)
        t, ca = fp_mul_wide(y, w, q)
        assert ca == int(ca) and t == int(t) and (int(t) - int(y) * w) % q == 0
        assert abs(t) < q
    assert broke > 0


@pytest.mark.parametrize("q,lim", [(Q49, 7.99), (Q50, 3.99)])
def test_fp_mul_wide_is_exact_below_2_pow_52(q, lim):
    rng = random.Random(3)
    assert lim * q < (1 << 52)
    for y, w in operands(q, lim, 4000, rng):
        t, ca = fp_mul_wide(y, w, q)
        assert ca == int(ca) and t == int(t)
        assert (int(t) - int(y) * w) % q == 0
        assert F(abs(int(t))) <= q * (F(1, 2) + min(F(abs(int(y)), 1 << 53), F(1, 4)) + F(abs(int(y)), 1 << 54))


@pytest.mark.parametrize("q", [Q49, Q50, 7681, 0x10001])
def test_fp_fold(q):
    rng = random.Random(4)
    for _ in range(4000):
        v = float(rng.choice([1, -1]) * rng.randrange(0, min(16 * q, (1 << 53) - 1)))
        r = fp_fold(v, q)
        assert r == int(r) and (int(r) - int(v)) % q == 0
        assert abs(r) <= q / 2 + 6


# ---- worst-case bounds of the pass schedules ----------------------------------------------------------------

class Schedule:
    """Mirror of fp_network(): tracks the largest magnitude any coefficient can have (exact rationals, absolute
    units) and checks every multiplied operand against the limit of the rounding used for it.

    |fold(v)| <= q/2 + 6.  With winv = RN(w/q) (absolute error <= 2^-54 because w < q):
      fp_mul       one fused rounding straight to an integer:   |t| <= q * (1/2 + |y| * 2^-54)
      fp_mul_wide  RN(y*winv), then rint of that (|y| < 2^52):   |t| <= q * (1/2 + min(|y| * 2^-53, 1/4) + |y| * 2^-54)
    """

    def __init__(self, q, q50):
        self.q, self.q50 = q, q50
        self.fold = F(q, 2) + 6

    def mul(self, operand, wide):
        assert operand < (1 << (52 if wide else 51)), (float(operand / self.q), wide)
        err = operand / (1 << 54)
        if wide:
            err += min(operand / (1 << 53), F(1, 4))
        return self.q * (F(1, 2) + err)

    def forward(self, R, b_in, role):
        """role: 1 first pass (input centred, |v| <= 2q), 0 middle, 2 last (the caller folds afterwards)"""
        assert b_in < (1 << 53)
        lean = not self.q50 and role != 0
        b = b_in if lean else self.fold
        for u in range(R):
            if self.q50 and R == 5 and u == 3:
                b = self.fold
            if self.q50:
                wide = R == 4 and u == 3
            else:
                wide = u >= 4 if role == 1 else (u >= 2 if role == 2 else False)
            b = b + self.mul(b, wide)
            assert b < (1 << 53)
        return b

    def inverse(self, R, b_in, final):
        assert b_in < (1 << 53)
        q50 = self.q50
        b = self.fold
        for u in range(R - 1, -1, -1):
            if (R == 5 and u == 1) or (q50 and R == 4 and u == 0):
                b = self.fold
            since = (5 - u if u >= 2 else 2 - u) if R == 5 else (1 if (q50 and R == 4 and u == 0) else R - u)
            wide = since >= 2 if q50 else since >= 4
            if final and u == 0:
                b = self.mul(2 * b, True)      # both outputs are products (of the sum and of the difference)
            else:
                b = max(2 * b, self.mul(2 * b, wide))
            assert b < (1 << 53)
        return b


@pytest.mark.parametrize("q,q50", [((1 << 49) - 1025, False), (Q49, False), ((1 << 49) - 1023, True), (Q50, True),
                                   (7681, False), (0x1fffffc800001, False), (0x3ffffffef4001, True)])
def test_range_schedules_stay_inside_their_limits(q, q50):
    s = Schedule(q, q50)
    for RA in (3, 4, 5):                       # L = 12, 13, 14
        # forward: input contract [0,4q); passes A (RA stages), B (5), C (4), then a fold
        b = s.forward(RA, 4 * q if q50 else 2 * q, 1)
        b = s.forward(5, b, 0)
        b = s.forward(4, b, 2)
        assert b < (1 << 53)                   # what the final fold accepts
        # inverse: input contract [0,2q); passes C (4), B (5), A (RA, with or without the N^-1 stage)
        b = s.inverse(4, 2 * q, False)
        b = s.inverse(5, b, False)
        s.inverse(RA, b, False)
        # the N^-1 products are converted without another fold: they must be below q in magnitude
        assert s.inverse(RA, b, True) < q, (RA, float(s.inverse(RA, b, True) / q))


def primes_below(top, step, count):
    def is_prime(n):
        if n % 2 == 0:
            return n == 2
        d, r = n - 1, 0
        while d % 2 == 0:
            d, r = d // 2, r + 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            x = pow(a, d, n)
            if x in (1, n - 1):
                continue
            for _ in range(r - 1):
                x = x * x % n
                if x == n - 1:
                    break
            else:
                return False
        return True

    q = top - ((top - 1) % step)
    while count:
        if is_prime(q):
            count -= 1
            yield q
        q -= step


@pytest.mark.parametrize("logn,top,n_values", [(13, (1 << 49) - 1024, 16), (12, (1 << 50) - 2048, 8)])
def test_final_stage_products_stay_below_q(logn, top, n_values):
    """The two geometries whose last inverse stage multiplies sums of almost 2^52 (49-bit q at N = 2^13, 50-bit q
    at N = 2^12) convert the N^-1 products without another fold, which needs |t| < q.  The bound above gives
    that (1/2 + 1/4 + <1/4); here the worst multipliers -- N^-1 * w_inv[1] close to q -- are tried on the
    largest moduli with operands at the top of the range."""
    N, rng = 1 << logn, random.Random(6)
    worst = 0.0
    for q in primes_below(top, 2 * N, 40):
        ninv = pow(N, -1, q)
        x = 2
        while pow(x, (q - 1) // 2, q) != q - 1:
            x += 1
        i = pow(x, (q - 1) // 4, q)             # a square root of -1: w_inv[1] is +-i
        for w in (ninv, ninv * i % q, ninv * (q - i) % q):
            for _ in range(300):
                v = rng.choice([1, -1]) * ((q - 1) // 2 - rng.choice([0, 1, rng.randrange(0, 1 << 30)]))
                s = n_values * v
                assert abs(s) < (1 << 52)
                t, _ = fp_mul_wide(float(s), w, q)
                assert abs(t) < q and (int(t) - s * w) % q == 0
                worst = max(worst, abs(t) / q)
    assert worst > 0.74                          # the sample does reach the regime the bound is about
