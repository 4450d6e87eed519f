"""CPU model of the FP64 butterfly arithmetic of csrc/ntt_ring_fp.cuh.

Python floats are IEEE binary64 and `float(Fraction)` rounds to nearest-even, so fused multiply-add can be
emulated exactly.  The tests below (no GPU needed) pin down the claims the kernel's range schedules rest on:

  * fp_mul (plain and coarse rounding) and fp_fold return exact integers congruent to the true product / value,
    with the stated magnitude bounds, for operands up to the limits the schedules allow -- and the plain rounding
    really does break just beyond 2^51 (which is why the coarse one exists);
  * the generated schedule tables (csrc/ntt_fp_schedule.h, tools/gen_fp_schedule.py) are the ones the generator
    produces today, and an INDEPENDENT bound propagation over those tables -- per position, exact rationals --
    stays inside every limit for the largest moduli each schedule serves and for small ones;
  * whole register networks, emulated instruction by instruction with the tables' fold / rounding decisions on
    extreme inputs, produce exact integers congruent to the true butterflies.
"""
import importlib.util
import os
import random
from fractions import Fraction as F

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("gen_fp_schedule", os.path.join(ROOT, "tools", "gen_fp_schedule.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

MAGIC = 6755399441055744.0      # 1.5 * 2^52
MAGIC2 = 13510798882111488.0    # 3 * 2^52
Q49 = (1 << 49) - 1024 - 1023   # largest odd value the first schedule accepts is 2^49 - 1025; any odd q below works
Q50 = (1 << 50) - 2049


def fma(a, b, c):
    return float(F(a) * F(b) + F(c))


def fp_fold(v, q):
    qd, qinv = float(q), 1.0 / float(q)
    k = fma(v, qinv, MAGIC) - MAGIC
    return fma(-k, qd, v)


def centred(w, q):
    """The table representative of a multiplier: w in (-q/2, q/2) (k_build_fd, finish_inverse_constants)."""
    w %= q
    return w - q if w > q // 2 else w


def fp_mul(y, w, q, coarse=False):
    """fp_mul of ntt_ring_fp.cuh with a TABLE multiplier: w is centred, winv = RN(w/q)."""
    qd = float(q)
    wc = float(centred(w, q))
    winv = wc / qd
    assert abs(winv) <= 0.5
    m = MAGIC2 if coarse else MAGIC
    cc = fma(y, winv, m) - m
    h = y * wc
    l = fma(y, wc, -h)
    d = fma(-cc, qd, h)
    return d + l, cc


def operands(q, limit_q, n, rng):
    """(y, w) pairs with |y| up to limit_q * q, biased to the extremes."""
    top = int(limit_q * q)
    for i in range(n):
        w = rng.choice([1, 2, q - 1, q - 2, (q + 1) // 2, (q - 1) // 2, (q - 3) // 2, rng.randrange(1, q)])
        mag = rng.choice([top, top - 1, top - rng.randrange(1, 1 << 20), rng.randrange(0, top + 1), q, q - 1, 0])
        yield float(rng.choice([1, -1]) * mag), w


@pytest.mark.parametrize("q,lim", [(Q49, 7.99), (Q50, 3.99)])
def test_fp_mul_is_exact_below_2_pow_52(q, lim):
    """Plain rounding with a centred multiplier: operands up to 2^52, |t| <= q*(1/2 + |y|*2^-55)."""
    rng = random.Random(1)
    assert lim * q < (1 << 52)
    for y, w in operands(q, lim, 4000, rng):
        t, cc = fp_mul(y, w, q)
        assert cc == int(cc) and t == int(t)
        assert (int(t) - int(y) * w) % q == 0
        assert F(abs(int(t))) <= q * (F(1, 2) + F(abs(int(y)), 1 << 55))


def test_plain_rounding_breaks_beyond_2_pow_52_and_coarse_does_not():
    """A negative quotient beyond 2^51 in magnitude lands where ulp = 1/2: the 1.5*2^52 constant yields a half-integer
    quotient (the failure once seen on the GPU); the 3*2^52 constant stays exact up to 2^52.  With centred multipliers
    (|w/q| <= 1/2) that takes an operand beyond 2^52."""
    q, rng = Q49, random.Random(2)
    broke = 0
    for _ in range(4000):
        w = (q - 1) // 2 - rng.randrange(0, 1 << 20)          # |w/q| close to 1/2, positive: y*w/q < -2^51
        y = -float(rng.randrange(int(8.4 * q), int(15.9 * q)))
        _, cc = fp_mul(y, w, q)
        broke += cc != int(cc)
        t, ca = fp_mul(y, w, q, coarse=True)
        assert ca == int(ca) and int(ca) % 2 == 0 and t == int(t) and (int(t) - int(y) * w) % q == 0
        assert F(abs(int(t))) <= q * (1 + F(abs(int(y)), 1 << 55))
    assert broke > 0


@pytest.mark.parametrize("q,lim", [(Q49, 15.99), (Q50, 7.99)])
def test_coarse_rounding_is_exact_below_2_pow_53(q, lim):
    rng = random.Random(3)
    assert lim * q < (1 << 53)
    for y, w in operands(q, lim, 4000, rng):
        t, ca = fp_mul(y, w, q, coarse=True)
        assert ca == int(ca) and int(ca) % 2 == 0 and t == int(t)
        assert (int(t) - int(y) * w) % q == 0
        assert F(abs(int(t))) <= q * (1 + F(abs(int(y)), 1 << 55))
        assert abs(t) < (1 << 53)


@pytest.mark.parametrize("q", [Q49, Q50, 7681, 0x10001])
def test_fp_fold(q):
    rng = random.Random(4)
    for _ in range(4000):
        v = float(rng.choice([1, -1]) * rng.randrange(0, min(16 * q, (1 << 53) - 1)))
        r = fp_fold(v, q)
        assert r == int(r) and (int(r) - int(v)) % q == 0
        assert abs(r) <= q / 2 + 6


# ---- the generated schedules ---------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def schedules():
    return gen.all_schedules()


def test_schedule_header_is_current():
    with open(gen.HEADER) as fh:
        assert fh.read() == gen.render(), "run python tools/gen_fp_schedule.py"


P51, P52, P53 = F(1 << 51), F(1 << 52), F(1 << 53)


def check_forward(passes, q):
    """Independent re-derivation: forward bounds are uniform; every fold mask must be all-or-nothing."""
    b = F(2 * q)
    fb = F(q, 2) + 6
    for p in passes:
        n = 1 << p.R
        full = (1 << n) - 1
        for s in range(p.R):
            assert p.fold_before[s] in (0, full) and p.coarse[s] in (0, full)
            if p.fold_before[s]:
                assert b < P53
                b = fb
            if p.coarse[s]:
                assert b < P53                               # |y*winv| <= |y|/2 < 2^52
                b = b + q * (1 + b / (1 << 55))
            else:
                assert b < P52                               # |y*winv| <= |y|/2 < 2^51
                b = b + q * (F(1, 2) + b / (1 << 55))
            assert b < P53
        assert p.fold_end == 0
    return b


def check_inverse_pass(p, b_in, q, final):
    n = 1 << p.R
    fb = F(q, 2) + 6
    b = [F(b_in)] * n
    for s in range(p.R):
        d = 1 << s
        for j in range(n):
            if (p.fold_before[s] >> j) & 1:
                assert b[j] < P53
                b[j] = fb
        nb = list(b)
        last = final and s == p.R - 1
        for lo in range(n):
            if lo & d:
                assert not (p.coarse[s] >> lo) & 1
                continue
            hi = lo + d
            D = b[lo] + b[hi]
            assert D < P53                                   # X + Y and X - Y are exact
            if (p.coarse[s] >> lo) & 1:
                assert D < P53 and not last
                t = q * (1 + D / (1 << 55))
            else:
                assert D < P52
                t = q * (F(1, 2) + D / (1 << 55))
            if last:
                nb[lo] = nb[hi] = t
            else:
                nb[lo], nb[hi] = D, t
        b = nb
    for j in range(n):
        if (p.fold_end >> j) & 1:
            b[j] = fb
    return max(b)


@pytest.mark.parametrize("q50,q", [(0, (1 << 49) - 1025), (0, Q49), (0, 0x1fffffc800001), (0, 7681),
                                   (1, (1 << 50) - 2049), (1, 0x3ffffffef4001), (1, (1 << 49) - 1023)])
def test_schedules_stay_inside_their_limits(schedules, q50, q):
    for L in (10, 11, 12, 13, 14):
        b = check_forward(schedules[("fwd", q50, L)], q)
        assert b < P53                                       # what the final fold accepts
        pc, pb, pa = schedules[("inv", q50, L)]
        pc2, pb2, pn = schedules[("invnf", q50, L)]
        assert (pc.fold_before, pc.coarse, pc.fold_end) == (pc2.fold_before, pc2.coarse, pc2.fold_end)
        assert (pb.fold_before, pb.coarse, pb.fold_end) == (pb2.fold_before, pb2.coarse, pb2.fold_end)
        bc = check_inverse_pass(pc, q, q, False)             # input [0,2q) centred to [-q,q)
        bb = check_inverse_pass(pb, bc, q, False)
        assert check_inverse_pass(pa, bb, q, True) < q       # N^-1 products are converted without another fold
        assert check_inverse_pass(pn, bb, q, False) < P53    # chunk of a larger transform: folded afterwards


@pytest.mark.parametrize("q50,q", [(0, (1 << 49) - 1025), (0, Q49), (0, 0x1fffffc800001), (0, 7681),
                                   (1, (1 << 50) - 2049), (1, Q50), (1, (1 << 49) - 1023)])
def test_strided_pass_schedules_stay_inside_their_limits(schedules, q50, q):
    """k_strided_fp (ntt_strided_fp.cuh): one network of R = 1..5 stages.  Forward input centred to |v| <= 2q, every
    output folded afterwards; inverse input centred to |v| <= q, either ending with the N^-1 stage (products below q,
    converted without a fold) or folded afterwards."""
    for R in range(1, 6):
        (pf,), (pi,), (pn,) = schedules[("sfwd", q50, R)], schedules[("sinv", q50, R)], schedules[("sinvnf", q50, R)]
        assert pf.R == pi.R == pn.R == R
        assert check_forward([pf], q) < P53
        assert check_inverse_pass(pi, q, q, True) < q
        assert check_inverse_pass(pn, q, q, False) < P53


@pytest.mark.parametrize("q50,q", [(0, (1 << 49) - 1025), (0, 0x1fffffc800001), (0, 7681), (1, (1 << 50) - 2049)])
def test_polymul_inverse_schedule(schedules, q50, q):
    """Inverse passes of the one-kernel multiply: input = product of two folded values (|p| <= 0.5625 q <= q), pass B
    split over two lanes -- from its second stage on, positions 2i and 2i+1 must be treated alike."""
    pc, pb, pa = schedules[("pminv", q50, 13)]
    fb = F(q, 2) + 6
    assert q * (F(1, 2) + fb / (1 << 53)) <= F(q) * F(9, 16) + 1      # the product bound the kernel comment states
    bc = check_inverse_pass(pc, q, q, False)
    bb = check_inverse_pass(pb, bc, q, False)
    assert check_inverse_pass(pa, bb, q, True) < q
    sym = lambda m: all(((m >> (2 * i)) & 1) == ((m >> (2 * i + 1)) & 1) for i in range(16))
    for s in range(1, 5):
        assert sym(pb.fold_before[s]) and sym(pb.coarse[s]), s
    assert sym(pb.fold_end)


def emulate_inverse_pass(p, x, tw, q, final, ninv=None, ninv_w=None):
    """The kernel's instruction sequence (fp_network_inv) on doubles; tw[(u, sub)] = twiddle of network stage u."""
    R, n = p.R, 1 << p.R
    for s in range(R):
        u, d = R - 1 - s, 1 << s
        for j in range(n):
            if (p.fold_before[s] >> j) & 1:
                x[j] = fp_fold(x[j], q)
        if final and u == 0:
            for k in range(d):
                sm, df = x[k] + x[k + d], x[k] - x[k + d]
                x[k], _ = fp_mul(sm, ninv, q)
                x[k + d], _ = fp_mul(df, ninv_w, q)
        else:
            for sub in range(1 << u):
                for k in range(d):
                    lo = sub * 2 * d + k
                    df = x[lo] - x[lo + d]
                    x[lo] = x[lo] + x[lo + d]
                    x[lo + d], _ = fp_mul(df, tw[(u, sub)], q, coarse=bool((p.coarse[s] >> lo) & 1))
    for j in range(n):
        if (p.fold_end >> j) & 1:
            x[j] = fp_fold(x[j], q)
    return x


@pytest.mark.parametrize("q50,q", [(0, Q49), (1, Q50)])
def test_inverse_networks_emulated_on_extreme_inputs(schedules, q50, q):
    """Pass C -> B -> A (with the N^-1 stage) of the L = 14 inverse, emulated with exact FMA on inputs pinned to
    the contract's edges: every intermediate is an integer below 2^53 and the result is congruent to the exact
    Gentleman-Sande network."""
    rng = random.Random(7 + q50)
    pc, pb, pa = schedules[("inv", q50, 14)]
    ninv, ninv_w = rng.randrange(1, q), rng.randrange(1, q)
    for trial in range(60):
        edge = rng.choice([q, -q, q - 1, 1 - q, None])
        x = [float(edge if edge is not None else rng.randrange(-q, q)) for _ in range(16)]
        exact = [int(v) for v in x]
        tw = {(u, sub): rng.choice([1, q - 1, rng.randrange(1, q)]) for u in range(5) for sub in range(1 << u)}
        y = emulate_inverse_pass(pc, list(x), tw, q, False)
        for s in range(4):                                   # exact network on integers
            u, d = 3 - s, 1 << s
            for sub in range(1 << u):
                for k in range(d):
                    lo = sub * 2 * d + k
                    a, b = exact[lo], exact[lo + d]
                    exact[lo], exact[lo + d] = a + b, (a - b) * tw[(u, sub)]
        for got, want in zip(y, exact):
            assert got == int(got) and abs(got) < (1 << 53) and (int(got) - want) % q == 0
        # the value a next-pass thread sees is any of these outputs: feed the largest ones into pass B and A
        big = max(y, key=abs)
        for p, final in ((pb, False), (pa, True)):
            n = 1 << p.R
            xin = [big if rng.random() < 0.7 else float(rng.randrange(-q // 2, q // 2)) for _ in range(n)]
            ex = [int(v) for v in xin]
            out = emulate_inverse_pass(p, list(xin), tw, q, final, ninv, ninv_w)
            for s in range(p.R):
                u, d = p.R - 1 - s, 1 << s
                if final and u == 0:
                    for k in range(d):
                        a, b = ex[k], ex[k + d]
                        ex[k], ex[k + d] = (a + b) * ninv, (a - b) * ninv_w
                else:
                    for sub in range(1 << u):
                        for k in range(d):
                            lo = sub * 2 * d + k
                            a, b = ex[lo], ex[lo + d]
                            ex[lo], ex[lo + d] = a + b, (a - b) * tw[(u, sub)]
            for got, want in zip(out, ex):
                assert got == int(got) and (int(got) - want) % q == 0
                assert abs(got) < (q if final else (1 << 53))
            big = max(out, key=abs) if not final else big


@pytest.mark.parametrize("q50,q", [(0, Q49), (1, Q50)])
def test_forward_network_emulated_on_extreme_inputs(schedules, q50, q):
    """The L = 14 forward transform as the kernel runs it (passes A, B, C: 14 stages; with centred multipliers the first
    schedule has NO fold before the final one), emulated with exact FMA on a pool of values that starts at the edges of
    the centred input range [-2q, 2q): every stage pairs pool members with multipliers near +-q/2, and every value must
    stay an integer, congruent to the exact butterfly, and inside the bound the schedule was derived from."""
    rng = random.Random(11 + q50)
    passes = schedules[("fwd", q50, 14)]
    fb = F(q, 2) + 6
    pool = [float(v) for v in (2 * q - 1, -2 * q, 2 * q - 2, -(2 * q - 1), q, -q, 1, 0)] + \
           [float(rng.randrange(-2 * q, 2 * q)) for _ in range(24)]
    exact = [int(v) for v in pool]
    b = F(2 * q)
    for p in passes:
        n = 1 << p.R
        for s in range(p.R):
            fold, coarse = bool(p.fold_before[s]), bool(p.coarse[s])
            assert p.fold_before[s] in (0, (1 << n) - 1) and p.coarse[s] in (0, (1 << n) - 1)
            if fold:
                pool = [fp_fold(v, q) for v in pool]
                b = fb
            b = b + (q * (1 + b / (1 << 55)) if coarse else q * (F(1, 2) + b / (1 << 55)))
            order = sorted(range(len(pool)), key=lambda i: -abs(pool[i]))      # pair the largest with the largest
            new, new_exact = list(pool), list(exact)
            for a in range(0, len(order), 2):
                i, j = order[a], order[a + 1]
                w = rng.choice([(q - 1) // 2, (q + 1) // 2, (q - 1) // 2 - rng.randrange(1 << 20), rng.randrange(1, q)])
                t, cc = fp_mul(pool[j], w, q, coarse=coarse)
                assert cc == int(cc) and t == int(t)
                new[i], new[j] = pool[i] + t, pool[i] - t
                new_exact[i], new_exact[j] = exact[i] + exact[j] * w, exact[i] - exact[j] * w
            pool, exact = new, new_exact
            for v, e in zip(pool, exact):
                assert v == int(v) and abs(v) < (1 << 53) and F(abs(int(v))) <= b and (int(v) - e) % q == 0
    assert b < P53                                           # what the final fold accepts
    for v, e in zip(pool, exact):
        r = fp_fold(v, q)
        assert r == int(r) and abs(r) <= q / 2 + 6 and (int(r) - e) % q == 0


@pytest.mark.parametrize("q50,q", [(0, Q49), (1, Q50)])
def test_strided_inverse_networks_emulated_on_extreme_inputs(schedules, q50, q):
    """k_strided_fp inverse passes (R = 1..5, with and without the N^-1 stage), emulated with exact FMA on inputs pinned
    to the edges of the centred input range [-q, q): integers below 2^53 throughout, congruent to the exact
    Gentleman-Sande network, and below q in magnitude after the N^-1 stage (converted without a fold)."""
    rng = random.Random(21 + q50)
    ninv, ninv_w = rng.randrange(1, q), rng.randrange(1, q)
    for R in range(1, 6):
        n = 1 << R
        for kind, final in (("sinv", True), ("sinvnf", False)):
            (p,) = schedules[(kind, q50, R)]
            for trial in range(25):
                edge = rng.choice([q, -q, q - 1, 1 - q, None])
                x = [float(edge if edge is not None else rng.randrange(-q, q)) for _ in range(n)]
                ex = [int(v) for v in x]
                tw = {(u, sub): rng.choice([1, q - 1, (q - 1) // 2, (q + 1) // 2, rng.randrange(1, q)])
                      for u in range(R) for sub in range(1 << u)}
                out = emulate_inverse_pass(p, list(x), tw, q, final, ninv, ninv_w)
                for s_ in range(R):
                    u, d = R - 1 - s_, 1 << s_
                    if final and u == 0:
                        for k in range(d):
                            a, b = ex[k], ex[k + d]
                            ex[k], ex[k + d] = (a + b) * ninv, (a - b) * ninv_w
                    else:
                        for sub in range(1 << u):
                            for k in range(d):
                                lo = sub * 2 * d + k
                                a, b = ex[lo], ex[lo + d]
                                ex[lo], ex[lo + d] = a + b, (a - b) * tw[(u, sub)]
                for got, want in zip(out, ex):
                    assert got == int(got) and (int(got) - want) % q == 0
                    assert abs(got) < (q if final else (1 << 53))


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


@pytest.mark.parametrize("logn,top", [(13, (1 << 49) - 1024), (12, (1 << 50) - 2048), (14, (1 << 49) - 1024)])
def test_final_stage_products_stay_below_q(logn, top):
    """The N^-1 stage converts its products without another fold, which needs |t| < q: the schedule keeps the
    operands of that stage below 2^52 so the plain rounding applies (|t| <= q*(1/2 + 1/8)).  The actual multipliers
    N^-1 and N^-1 * w_inv[1] (centred) are tried on the largest moduli with operands at the top of that range."""
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
                s = rng.choice([1, -1]) * ((1 << 52) - 1 - rng.choice([0, 1, rng.randrange(0, 1 << 30)]))
                t, _ = fp_mul(float(s), w, q)
                assert abs(t) < q and (int(t) - s * w) % q == 0
                worst = max(worst, abs(t) / q)
    assert worst > 0.55                          # the sample does reach the regime the bound is about
