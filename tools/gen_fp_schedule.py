#!/usr/bin/env python
"""Range schedules of the FP64 ring kernels (csrc/ntt_ring_fp.cuh): where values are folded and which rounding a
product uses.  This file is the single source of truth: it derives the schedules with exact rational bounds,
writes csrc/ntt_fp_schedule.h, and tests/test_fp64_arith_model.py imports it to re-check every bound and to make
sure the header on disk is the one this model produces.

    python tools/gen_fp_schedule.py            # rewrite the header
    python tools/gen_fp_schedule.py --check    # exit 1 if the header is stale

Arithmetic recap (see the header of ntt_ring_fp.cuh).  Coefficients are integers held in doubles, |v| < 2^53.
Twiddles are stored CENTRED: w in (-q/2, q/2) with winv = RN(w/q), |winv| <= 1/2, absolute error <= 2^-55 (the doubles
below 1/2 are 2^-54 apart).  For an integer operand y:
    plain   c = (y*winv + 1.5*2^52) - 1.5*2^52   one rounding to the nearest integer, needs |y*winv| < 2^51, i.e.
            |y| < 2^52:  |t| <= q*(1/2 + |y|*2^-55)
    coarse  c = (y*winv + 3*2^52) - 3*2^52       one rounding to the nearest EVEN integer, needs |y*winv| < 2^52, i.e.
            |y| < 2^53:  |t| <= q*(1 + |y|*2^-55)
both followed by the same exact h/l/d/t steps (6 FP64 instructions either way).  A fold is v - rint(v/q)*q,
|result| <= q/2 + 6 for |v| < 2^53 (3 instructions).
(Until the centred tables, w lay in [0,q): the operand limits were 2^51 / 2^52 and the error term |y|*2^-54, which cost
the N = 2^14 forward transform a fold of every value after its eighth stage; centred, fourteen stages fit under 2^53
without one.)

Forward (Cooley-Tukey, X' = X + t, Y' = X - t): every value of a stage has the same bound b' = b + |t|(b), so the
schedule is per stage: plain while b < 2^52, coarse above, and a fold of everything first whenever the new bound would
reach 2^53.
Inverse (Gentleman-Sande, X' = X + Y, Y' = t(X - Y)): sums double but products come back small, so bounds depend
on the POSITION inside the register network -- only a few of the 2^R values ever get large.  The schedule tracks
one bound per position and folds exactly the positions that would break a limit (sums and product operands < 2^53,
< 2^52 where the rounding must be plain), plus the positions above a cap at the end of a pass so the next pass (whose threads regroup the
values) can start from one uniform bound.  The caps are searched for the fewest folds.
"""
import os
import sys
from fractions import Fraction as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "optimized-number-theoretic-transform-implementations_b200", "csrc", "ntt_fp_schedule.h")

P51, P52, P53 = F(1 << 51), F(1 << 52), F(1 << 53)
QMAX = {0: (1 << 49) - 1024, 1: (1 << 50) - 2048}     # largest modulus each schedule serves


def fold_bound(q):
    return F(q, 2) + 6


def t_plain(y, q):
    if isinstance(y, float):                              # cap search: floats are enough to rank candidates
        return q * (0.5 + y / 36028797018963968.0)
    return q * (F(1, 2) + y / (1 << 55))


def t_coarse(y, q):
    if isinstance(y, float):
        return q * (1.0 + y / 36028797018963968.0)
    return q * (1 + y / (1 << 55))


class Pass:
    """One register network of R stages, in processing order.  fold_before[s] / coarse[s]: bit i set = position i
    is folded before stage s / the butterfly whose LOWER position is i uses the coarse rounding.  fold_end: positions
    folded after the last stage."""

    def __init__(self, R):
        self.R = R
        self.fold_before = [0] * R
        self.coarse = [0] * R
        self.fold_end = 0
        self.b_out = None

    def folds(self):
        return sum(bin(m).count("1") for m in self.fold_before) + bin(self.fold_end).count("1")


def forward_schedule(L, q50):
    """Passes A (L-9 stages), B (5), C (4) of the forward chunk transform; input centred to |v| <= 2q."""
    q = QMAX[q50]
    shapes = [L - 9, 5, 4]
    passes = [Pass(R) for R in shapes]
    b = F(2 * q)
    for p in passes:
        n = 1 << p.R
        for s in range(p.R):
            b = forward_stage(p, s, n, b, q)
        p.b_out = b
    return passes


def forward_stage(p, s, n, b, q):
    """One forward stage on the uniform bound b: plain rounding while the operand is below 2^52, coarse above; everything
    is folded first if the operand or the new bound would reach 2^53.  Returns the new bound."""
    def step(v):
        return (v + t_plain(v, q), False) if v < P52 else (v + t_coarse(v, q), True)
    nb, coarse = step(b)
    if b >= P53 or nb >= P53:
        p.fold_before[s] = (1 << n) - 1
        nb, coarse = step(fold_bound(q))
    if coarse:
        p.coarse[s] = (1 << n) - 1
    assert nb < P53
    return nb


def inverse_pass(R, b_in, q, final, cap_out, paired=False):
    """Position-aware inverse network.  Stage s (processing order) pairs positions at distance d = 2^s.
    final: the last stage is global stage 0, BOTH outputs are products (by N^-1 and N^-1*w) and must come out
    below q in magnitude because they are converted without another fold -> plain rounding only.
    paired: positions 2i and 2i+1 live in two different lanes that run the same instruction stream from stage 1 on
    (k_polymul_fp splits its 32-value inverse pass B over both half-warps: stage 0 pairs (2i, 2i+1) across the two
    lanes, the later stages run on each lane's 16 values), so from stage 1 on every decision is taken jointly for
    positions 2i and 2i+1 and the masks come out symmetric."""
    n = 1 << R
    fb = float(fold_bound(q)) if isinstance(b_in, float) else fold_bound(q)
    b = [b_in] * n
    p = Pass(R)
    for s in range(R):
        d = 1 << s
        last = final and s == R - 1
        lim = P52 if last else P53                      # operand of a plain / coarse product
        joint = paired and s >= 1
        for lo in range(n):
            if lo & d or (joint and lo & 1):
                continue
            los = (lo, lo + 1) if joint else (lo,)
            while any(b[l] + b[l + d] >= lim or b[l] + b[l + d] >= P53 for l in los):
                side = 0 if max(b[l] for l in los) >= max(b[l + d] for l in los) else d
                if all(b[l + side] <= fb for l in los):
                    return None                      # cannot be scheduled (does not happen for the moduli served)
                for l in los:
                    b[l + side] = min(b[l + side], fb)
                    p.fold_before[s] |= 1 << (l + side)
        nb = list(b)
        for lo in range(n):
            if lo & d or (joint and lo & 1):
                continue
            los = (lo, lo + 1) if joint else (lo,)
            coarse = any(b[l] + b[l + d] >= P52 for l in los)
            for l in los:
                D = b[l] + b[l + d]
                if coarse:
                    t = t_coarse(D, q)
                    p.coarse[s] |= 1 << l
                else:
                    t = t_plain(D, q)
                if last:
                    nb[l] = nb[l + d] = t
                else:
                    nb[l], nb[l + d] = D, t
        b = nb
    if cap_out is not None:
        for j in range(0, n, 2 if paired else 1):
            js = (j, j + 1) if paired else (j,)
            if any(b[x] > cap_out for x in js):
                for x in js:
                    b[x] = min(b[x], fb)
                    p.fold_end |= 1 << x
    p.b_out = max(b)
    return p


def inverse_schedule(L, q50):
    """Passes C (4 stages, input centred to |v| <= q), B (5), A (L-9 stages) in two forms: with the N^-1 stage
    (the chunk is the whole polynomial) and without (strided passes follow; the kernel folds and converts every
    value afterwards).  Passes C and B are shared by both forms.  The caps between the passes are chosen for the
    fewest folds per thread (pass C runs twice per thread); the search runs on floats, the winner is then
    re-derived with exact rationals."""
    q = QMAX[q50]
    best = None
    steps = list(range(2, 33))
    for xc in steps:
        pc = inverse_pass(4, float(q), q, False, xc / 4 * q)
        if pc is None:
            continue
        for xb in steps:
            pb = inverse_pass(5, pc.b_out, q, False, xb / 4 * q)
            if pb is None:
                continue
            pa = inverse_pass(L - 9, pb.b_out, q, True, None)
            pn = inverse_pass(L - 9, pb.b_out, q, False, None)
            if pa is None or pn is None or pa.b_out >= 0.999 * q or pn.b_out >= 0.999 * float(P53):
                continue
            cost = 2 * pc.folds() + pb.folds() + pa.folds()
            if best is None or cost < best[0]:
                best = (cost, xc, xb)
    assert best is not None
    _, xc, xb = best
    pc = inverse_pass(4, F(q), q, False, F(xc, 4) * q)
    pb = inverse_pass(5, pc.b_out, q, False, F(xb, 4) * q)
    pa = inverse_pass(L - 9, pb.b_out, q, True, None)
    pn = inverse_pass(L - 9, pb.b_out, q, False, None)
    assert pa.b_out < q and pn.b_out < P53
    return pc, pb, pa, pn


def polymul_inverse_schedule(q50):
    """Inverse passes of k_polymul_fp (N = 2^13): C (4 stages, input = the product of two folded values,
    |p| <= 0.5625 q, bounded by q here), B split over both half-warps (paired), A (4 stages with the N^-1 stage)."""
    q = QMAX[q50]
    best = None
    steps = list(range(2, 33))
    for xc in steps:
        pc = inverse_pass(4, float(q), q, False, xc / 4 * q)
        if pc is None:
            continue
        for xb in steps:
            pb = inverse_pass(5, pc.b_out, q, False, xb / 4 * q, paired=True)
            if pb is None:
                continue
            pa = inverse_pass(4, pb.b_out, q, True, None)
            if pa is None or pa.b_out >= 0.999 * q:
                continue
            # per thread: pass C once, pass B on 16 of the 32 positions (both lanes run the same stream), pass A once
            cost = pc.folds() + pb.folds() / 2 + pa.folds()
            if best is None or cost < best[0]:
                best = (cost, xc, xb)
    assert best is not None
    _, xc, xb = best
    pc = inverse_pass(4, F(q), q, False, F(xc, 4) * q)
    pb = inverse_pass(5, pc.b_out, q, False, F(xb, 4) * q, paired=True)
    pa = inverse_pass(4, pb.b_out, q, True, None)
    assert pa.b_out < q
    return pc, pb, pa


def forward_pass_schedule(R, q50):
    """One forward network of R stages on input centred to |v| <= 2q (a strided pass of a transform larger than a
    chunk: its input is the caller's [0,4q), or the canonical output of the strided pass before it)."""
    q = QMAX[q50]
    p = Pass(R)
    n = 1 << R
    b = F(2 * q)
    for s in range(R):
        b = forward_stage(p, s, n, b, q)
    p.b_out = b
    return p


def strided_schedules(q50):
    """FP64 strided passes (k_strided_fp): R = 1..5 stages over global memory.  Forward: see forward_pass_schedule; every
    output is folded and converted to the canonical residue.  Inverse: input = canonical residues (contract [0,2q),
    centred to |v| <= q); `final`: the pass ends with global stage 0 (N^-1 products, converted without another fold),
    otherwise every output is folded and converted."""
    q = QMAX[q50]
    fwd = [forward_pass_schedule(R, q50) for R in range(1, 6)]
    inv = [inverse_pass(R, F(q), q, True, None) for R in range(1, 6)]
    invnf = [inverse_pass(R, F(q), q, False, None) for R in range(1, 6)]
    for p in inv:
        assert p is not None and p.b_out < q
    for p in invnf:
        assert p is not None and p.b_out < P53
    return fwd, inv, invnf


def all_schedules():
    out = {}
    for q50 in (0, 1):
        fwd, inv, invnf = strided_schedules(q50)
        for R in range(1, 6):
            out[("sfwd", q50, R)] = [fwd[R - 1]]
            out[("sinv", q50, R)] = [inv[R - 1]]
            out[("sinvnf", q50, R)] = [invnf[R - 1]]
    for q50 in (0, 1):
        out[("pminv", q50, 13)] = list(polymul_inverse_schedule(q50))
    for q50 in (0, 1):
        for L in (10, 11, 12, 13, 14):
            out[("fwd", q50, L)] = forward_schedule(L, q50)
            pc, pb, pa, pn = inverse_schedule(L, q50)
            out[("inv", q50, L)] = [pc, pb, pa]
            out[("invnf", q50, L)] = [pc, pb, pn]
    return out


def render():
    sch = all_schedules()
    lines = [
        "/* csrc/ntt_fp_schedule.h -- GENERATED by tools/gen_fp_schedule.py; do not edit.",
        " * Range schedules of the FP64 ring kernels: which positions of a register network are folded before each",
        " * stage (processing order), which butterflies use the coarse quotient rounding, which positions are folded",
        " * after the last stage.  Bit i = position i (for `coarse`: the butterfly whose lower position is i).",
        " * Indexed [Q50][L-10]; passes A, B, C as in ntt_ring_fp.cuh.  tests/test_fp64_arith_model.py re-derives",
        " * every bound with exact rationals and fails if this file is stale. */",
        "#pragma once",
        "#include <cstdint>",
        "namespace nttb200 {",
        "struct FpPass {",
        "  uint32_t fold_before[5];",
        "  uint32_t coarse[5];",
        "  uint32_t fold_end;",
        "};",
        "struct FpSchedule {",
        "  FpPass a, b, c;",
        "};",
    ]

    def pass_txt(p):
        fb = p.fold_before + [0] * (5 - p.R)
        co = p.coarse + [0] * (5 - p.R)
        return "{{%s}, {%s}, %s}" % (", ".join("0x%08xu" % m for m in fb), ", ".join("0x%08xu" % m for m in co),
                                     "0x%08xu" % p.fold_end)

    for kind, name in (("fwd", "FP_SCHED_FWD"), ("inv", "FP_SCHED_INV"), ("invnf", "FP_SCHED_INV_NOFINAL")):
        lines.append("constexpr FpSchedule %s[2][5] = {" % name)
        for q50 in (0, 1):
            lines.append("  {")
            for L in (10, 11, 12, 13, 14):
                ps = sch[(kind, q50, L)]
                if kind == "fwd":
                    a, b, c = ps
                else:
                    c, b, a = ps
                lines.append("    /* Q50=%d L=%d: %d folds per thread */" % (
                    q50, L, (2 * c.folds() + b.folds() + a.folds()) if kind != "fwd" else
                    sum(bin(m).count("1") for p in ps for m in p.fold_before) + 16 * 0))
                lines.append("    {%s,\n     %s,\n     %s}," % (pass_txt(a), pass_txt(b), pass_txt(c)))
            lines.append("  },")
        lines.append("};")
    lines.append("/* strided passes in FP64 (k_strided_fp): one network of R = 1..5 stages, indexed [Q50][R-1].  Forward: input centred")
    lines.append(" * to |v| <= 2q; inverse: input centred to |v| <= q, with / without the N^-1 stage at the end. */")
    for kind, name in (("sfwd", "FP_SCHED_STRIDED_FWD"), ("sinv", "FP_SCHED_STRIDED_INV"), ("sinvnf", "FP_SCHED_STRIDED_INV_NOFINAL")):
        lines.append("constexpr FpPass %s[2][5] = {" % name)
        for q50 in (0, 1):
            lines.append("  {" + ",\n   ".join(pass_txt(sch[(kind, q50, R)][0]) for R in range(1, 6)) + "},")
        lines.append("};")
    lines.append("/* inverse passes of the one-kernel multiply (N = 2^13); pass B is split over both half-warps: from its")
    lines.append(" * second stage on the masks are symmetric in positions 2i / 2i+1 */")
    lines.append("constexpr FpSchedule FP_SCHED_INV_POLYMUL[2] = {")
    for q50 in (0, 1):
        c, b, a = sch[("pminv", q50, 13)]
        lines.append("  /* Q50=%d */" % q50)
        lines.append("  {%s,\n   %s,\n   %s}," % (pass_txt(a), pass_txt(b), pass_txt(c)))
    lines.append("};")
    lines.append("}  // namespace nttb200")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    txt = render()
    if "--check" in sys.argv:
        ok = os.path.exists(HEADER) and open(HEADER).read() == txt
        print("ntt_fp_schedule.h is %s" % ("up to date" if ok else "STALE"))
        sys.exit(0 if ok else 1)
    with open(HEADER, "w") as fh:
        fh.write(txt)
    for k, ps in sorted(all_schedules().items(), key=lambda kv: str(kv[0])):
        print(k, [(p.R, p.folds(), float(p.b_out / QMAX[k[1]])) for p in ps])
