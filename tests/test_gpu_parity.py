"""GPU parity tests: the CUDA path, called through the C-ABI (include/ntt_b200.h), against the oracle.

They mirror tests/test_correctness.c of the reference (every case: forward == fwd_ntt_ref_harvey after full
reduction, inverse(forward(a)) == a, the _dbl entry point equals two single transforms) and add what
SURVEY.md section 4 asks for: full-range and lazy-range inputs, edge vectors, batches of distinct
polynomials, device-generated tables, and size-independent properties at the benchmark sizes.
Bit-exact comparison everywhere (integer arithmetic).
"""
import numpy as np
import pytest

from conftest import CaseTables, edge_inputs

pytestmark = pytest.mark.gpu

ALL_CASES = list(range(19))


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def to_dev(a):
    torch = _torch()
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_host(t):
    return t.cpu().numpy().view(np.uint64)


# ---- the 19 fixture cases through the reference-shaped entry points ------------------------------------

@pytest.mark.parametrize("idx", ALL_CASES)
def test_fixture_case_dropin(ntt, oracle, golden_cases, case_tables, idx):
    g, t = golden_cases[idx], case_tables(idx)
    a = oracle.uniform(t.N, t.q, 0x5EED0000 + idx)
    assert "%016x" % oracle.fnv(a) == g["in_fnv"]

    x = a.copy()
    ntt.fwd_ntt_ref_harvey(x, t.N, t.q, t.w, t.w_con)
    assert np.array_equal(x, oracle.fwd(a, t.q, t.w, t.w_con)), "forward differs from oracle"
    assert "%016x" % oracle.fnv(x) == g["fwd_fnv"], "forward differs from the reference's golden hash"

    # lazy entry point: contract is [0,4q) and equality after full reduction (test_correctness.c:267-269)
    y = a.copy()
    ntt.fwd_ntt_ref_harvey_lazy(y, t.N, t.q, t.w, t.w_con)
    assert (y < 4 * t.q).all()
    assert np.array_equal(y % np.uint64(t.q), x)

    ntt.inv_ntt_ref_harvey(x, t.N, t.q, t.n_inv, t.n_inv_con, 64, t.w_inv, t.w_inv_con)
    assert np.array_equal(x, a), "inverse(forward(a)) != a"


@pytest.mark.parametrize("idx", ALL_CASES)
def test_fixture_case_lazy_range_inputs(ntt, oracle, golden_cases, case_tables, idx):
    """Inputs at the edge of the reference's input contracts: [0,4q) forward, [0,2q) inverse."""
    g, t = golden_cases[idx], case_tables(idx)
    a4 = oracle.uniform(t.N, 4 * t.q, 0x4A2F0000 + idx)
    x = a4.copy()
    ntt.fwd_ntt_ref_harvey(x, t.N, t.q, t.w, t.w_con)
    assert "%016x" % oracle.fnv(x) == g["fwd4q_fnv"]
    assert np.array_equal(x, oracle.fwd(a4, t.q, t.w, t.w_con))

    a2 = oracle.uniform(t.N, 2 * t.q, 0x2A2F0000 + idx)
    y = a2.copy()
    ntt.inv_ntt_ref_harvey(y, t.N, t.q, t.n_inv, t.n_inv_con, 64, t.w_inv, t.w_inv_con)
    assert "%016x" % oracle.fnv(y) == g["inv2q_fnv"]
    assert np.array_equal(y, oracle.inv(a2, t.q, t.n_inv, t.w_inv, t.w_inv_con))

    # worst case for the lazy accumulators: every coefficient at the top of the range
    top = np.full(t.N, 4 * t.q - 1, dtype=np.uint64)
    z = top.copy()
    ntt.fwd_ntt_ref_harvey(z, t.N, t.q, t.w, t.w_con)
    assert np.array_equal(z, oracle.fwd(top, t.q, t.w, t.w_con))
    top2 = np.full(t.N, 2 * t.q - 1, dtype=np.uint64)
    z = top2.copy()
    ntt.inv_ntt_ref_harvey(z, t.N, t.q, t.n_inv, t.n_inv_con, 64, t.w_inv, t.w_inv_con)
    assert np.array_equal(z, oracle.inv(top2, t.q, t.n_inv, t.w_inv, t.w_inv_con))


@pytest.mark.parametrize("idx", [0, 5, 9, 12, 13, 15, 18])
def test_fixture_case_edges_and_dbl(ntt, oracle, golden_cases, case_tables, idx):
    g, t = golden_cases[idx], case_tables(idx)
    for name, v in edge_inputs(t.N, t.q).items():
        x = v.copy()
        ntt.fwd_ntt_ref_harvey(x, t.N, t.q, t.w, t.w_con)
        assert "%016x" % oracle.fnv(x) == g["edges"][name]["fwd"], name
        y = v.copy()
        ntt.inv_ntt_ref_harvey(y, t.N, t.q, t.n_inv, t.n_inv_con, 64, t.w_inv, t.w_inv_con)
        assert "%016x" % oracle.fnv(y) == g["edges"][name]["inv"], name
    # double-input entry point: both lanes equal the single-input result (test_correctness.c:40-59)
    a = oracle.uniform(t.N, t.q, 77 + idx)
    a1, a2 = a.copy(), a.copy()
    ntt.fwd_ntt_ref_harvey_dbl(a1, a2, t.N, t.q, t.w, t.w_con)
    want = oracle.fwd(a, t.q, t.w, t.w_con)
    assert np.array_equal(a1, want) and np.array_equal(a2, want)


def test_dropin_unaligned_host_pointers(ntt, oracle, case_tables):
    """The reference's bench passes (64-byte aligned + 8) arrays (tests/test_cases.h:37-46, bench.c:160-186);
    the reference-shaped entry points must take any 8-byte aligned host pointer."""
    t = case_tables(9)
    buf = np.zeros(t.N + 9, dtype=np.uint64)
    off = ((8 - buf.ctypes.data % 64) % 64) // 8
    a = buf[off:off + t.N]
    assert a.ctypes.data % 64 == 8 and a.flags["C_CONTIGUOUS"]
    src = oracle.uniform(t.N, t.q, 99)
    a[:] = src
    ntt.fwd_ntt_ref_harvey(a, t.N, t.q, t.w, t.w_con)
    assert np.array_equal(a, oracle.fwd(src, t.q, t.w, t.w_con))
    ntt.inv_ntt_ref_harvey(a, t.N, t.q, t.n_inv, t.n_inv_con, 64, t.w_inv, t.w_inv_con)
    assert np.array_equal(a, src)


def test_case0_full_vectors(ntt, golden_case0):
    g = golden_case0
    N, q = 1 << g["m"], g["q"]
    w, wc = np.array(g["w"], dtype=np.uint64), np.array(g["w_con"], dtype=np.uint64)
    a = np.array(g["a"], dtype=np.uint64)
    ntt.fwd_ntt_ref_harvey(a, N, q, w, wc)
    assert a.tolist() == g["fwd"]


# ---- plan / batch API on device-resident data -------------------------------------------------------

@pytest.mark.parametrize("idx", [0, 3, 6, 9, 12, 13, 14, 15, 16, 17, 18])
def test_batch_api_distinct_polynomials(ntt, oracle, case_tables, idx):
    t = case_tables(idx)
    batch = 5 if t.m <= 14 else 3
    plan = ntt.Plan.from_tables(t.N, t.q, t.w, t.w_con, t.w_inv, t.w_inv_con, t.n_inv, t.n_inv_con)
    a = oracle.uniform(batch * t.N, t.q, 1000 + idx).reshape(batch, t.N)
    d = to_dev(a)
    plan.fwd(d, batch)
    f = to_host(d)
    assert np.array_equal(f, oracle.fwd(a, t.q, t.w, t.w_con))
    plan.inv(d, batch)
    assert np.array_equal(to_host(d), a)
    plan.close()


@pytest.mark.parametrize("idx", [0, 4, 9, 13, 16, 18])
def test_device_generated_tables_match_reference_tables(ntt, oracle, golden_cases, case_tables, idx):
    """north star item 4: tables generated on the device equal calc_w / calc_w_con output bit for bit."""
    g, t = golden_cases[idx], case_tables(idx)
    plan = ntt.Plan.from_psi(t.N, t.q, t.psi)
    tb = plan.export_tables()
    assert "%016x" % oracle.fnv(tb["w"]) == g["w_fnv"]
    assert "%016x" % oracle.fnv(tb["w_con"]) == g["w_con_fnv"]
    assert "%016x" % oracle.fnv(tb["w_inv"]) == g["w_inv_fnv"]
    assert "%016x" % oracle.fnv(tb["w_inv_con"]) == g["w_inv_con_fnv"]
    assert tb["n_inv"] == g["n_inv"] and tb["n_inv_con"] == g["n_inv_con"]
    a = oracle.uniform(t.N, t.q, 0x5EED0000 + idx)
    d = to_dev(a)
    plan.fwd(d, 1)
    assert "%016x" % oracle.fnv(to_host(d)) == g["fwd_fnv"]
    plan.close()


def test_exact_path_large_modulus(ntt, oracle):
    """q above the lazy fast path's limit (2^56) takes the general Harvey kernel (valid to 2^62)."""
    for bits, m in ((60, 10), (61, 13), (58, 15)):
        N = 1 << m
        q = (1 << bits) - (1 << bits) % (2 * N) + 1
        while not oracle.is_prime(q):
            q -= 2 * N
        psi = oracle.min_root(N, q) if m <= 10 else None
        if psi is None:
            # any primitive 2N-th root will do for parity
            x = 2
            while True:
                c = oracle.powmod(x, (q - 1) // (2 * N), q)
                if oracle.powmod(c, N, q) == q - 1:
                    psi = c
                    break
                x += 1
        t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
        plan = ntt.Plan.from_tables(t.N, t.q, t.w, t.w_con, t.w_inv, t.w_inv_con, t.n_inv, t.n_inv_con)
        assert not plan.is_lazy
        a = oracle.uniform(2 * N, 4 * q, bits).reshape(2, N)
        d = to_dev(a)
        plan.fwd(d, 2)
        f = to_host(d)
        assert np.array_equal(f, oracle.fwd(a, q, t.w, t.w_con))
        plan.inv(d, 2)
        assert np.array_equal(to_host(d), a % np.uint64(q))
        plan.close()


@pytest.mark.parametrize("bits,m", [(50, 14), (50, 13), (53, 14), (56, 14), (56, 16), (49, 12)])
def test_integer_ring_kernel_wide_moduli(ntt, oracle, bits, m):
    """2^49 <= q < 2^56 runs the integer lazy ring kernel (no FP64): bounds, renormalisation schedule and the
    final reduction are exercised with the largest primes of each size and inputs at the top of the contracts."""
    N = 1 << m
    q = (1 << bits) - ((1 << bits) - 1) % (2 * N)
    while not oracle.is_prime(q) or q >= (1 << bits):
        q -= 2 * N
    x = 2
    while True:
        psi = oracle.powmod(x, (q - 1) // (2 * N), q)
        if oracle.powmod(psi, N, q) == q - 1:
            break
        x += 1
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    plan = ntt.Plan.from_psi(N, q, psi)
    assert plan.is_lazy
    batch = 160 if m <= 14 else 6
    a = oracle.uniform(batch * N, 4 * q, bits).reshape(batch, N)
    a[1, :] = 4 * q - 1
    b = oracle.uniform(batch * N, 2 * q, bits + 1).reshape(batch, N)
    b[1, :] = 2 * q - 1
    da, db = to_dev(a), to_dev(b)
    plan.fwd(da, batch)
    plan.inv(db, batch)
    fa, ib = to_host(da), to_host(db)
    for r in (0, 1, batch - 1):
        assert np.array_equal(fa[r], oracle.fwd(a[r], q, t.w, t.w_con)), "forward row %d" % r
        assert np.array_equal(ib[r], oracle.inv(b[r], q, t.n_inv, t.w_inv, t.w_inv_con)), "inverse row %d" % r
    plan.inv(da, batch)
    assert np.array_equal(to_host(da), a % np.uint64(q))
    plan.close()


def test_synthetic_configs_golden(ntt, oracle, golden_synth):
    """The throughput parameter sets (49-bit q; N = 2^13, 2^14, 2^16) against reference-generated hashes."""
    for s in golden_synth:
        N, q = 1 << s["m"], s["q"]
        plan = ntt.Plan.from_psi(N, q, s["psi"])
        a = oracle.uniform(s["batch"] * N, q, s["seed"]).reshape(s["batch"], N)
        assert "%016x" % oracle.fnv(a) == s["in_fnv"]
        d = to_dev(a)
        plan.fwd(d, s["batch"])
        assert "%016x" % oracle.fnv(to_host(d)) == s["fwd_fnv"]
        plan.inv(d, s["batch"])
        assert np.array_equal(to_host(d), a)
        plan.close()


# ---- size-independent properties at the benchmark sizes ------------------------------------------------

def _spot_check(plan_tables, oracle, a, f, rows):
    t = plan_tables
    for r in rows:
        assert np.array_equal(f[r], oracle.fwd(a[r], t.q, t.w, t.w_con)), "row %d" % r


def test_headline_config_roundtrip_and_linearity(ntt, oracle, golden_synth):
    """BASELINE config 2: N=2^14, 49-bit q, batch 4096 -- round trip, linearity, spot rows vs oracle."""
    torch = _torch()
    s = [x for x in golden_synth if x["m"] == 14][0]
    N, q, batch = 1 << 14, s["q"], 4096
    t = CaseTables(oracle, 14, q, s["psi"], s["psi_inv"], s["n_inv"])
    plan = ntt.Plan.from_psi(N, q, s["psi"])
    a = oracle.uniform(batch * N, q, 1).reshape(batch, N)
    b = oracle.uniform(batch * N, q, 11).reshape(batch, N)
    da, db = to_dev(a), to_dev(b)
    dsum = to_dev((a + b) % np.uint64(q))
    plan.fwd(da, batch)
    plan.fwd(db, batch)
    plan.fwd(dsum, batch)
    fa, fb, fs = to_host(da), to_host(db), to_host(dsum)
    assert (fa < q).all() and (fb < q).all()
    assert np.array_equal((fa + fb) % np.uint64(q), fs), "NTT(a+b) != NTT(a)+NTT(b)"
    _spot_check(t, oracle, a, fa, [0, 1, 777, 2048, 4095])
    plan.inv(da, batch)
    assert np.array_equal(to_host(da), a), "round trip failed at full batch"
    torch.cuda.synchronize()
    plan.close()


@pytest.mark.parametrize("m,bits", [(14, 49), (14, 50), (13, 49), (13, 50), (12, 49), (12, 50), (11, 49), (11, 50),
                                    (10, 49), (10, 50),                                     (16, 49), (16, 50)])
def test_kernel_paths_agree_on_full_batch(ntt, oracle, golden_synth, m, bits):
    """The three CUDA paths (FP64 ring, integer ring, generic smem kernel) must produce identical bytes on a
    large batch, forward and inverse, on inputs at the edge of the contracts; rows are spot-checked vs the oracle.
    bits = 49: the headline modulus (first FP64 range schedule); 50: the largest 50-bit prime (second schedule).
    m = 10 .. 14: the five ring-kernel geometries (10, 11: FP64 ring kernel only); 16: strided pass + 2^14 chunks."""
    N = 1 << m
    batch = (1 << 25) >> m  # 2^25 coefficients per direction: 2048 polynomials at N = 2^14
    if bits == 49:
        q = 0x1FFFFFC800001
    else:
        q = (1 << 50) - ((1 << 50) - 1) % (2 * N)
        while not oracle.is_prime(q) or q > (1 << 50) - 2048:
            q -= 2 * N
    x = 2
    while True:
        psi = oracle.powmod(x, (q - 1) // (2 * N), q)
        if oracle.powmod(psi, N, q) == q - 1:
            break
        x += 1
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    plan = ntt.Plan.from_psi(N, q, psi)
    a4 = oracle.uniform(batch * N, 4 * q, 51).reshape(batch, N)   # forward contract [0,4q)
    a2 = oracle.uniform(batch * N, 2 * q, 52).reshape(batch, N)   # inverse contract [0,2q)
    a4[7, :] = 4 * q - 1
    a2[7, :] = 2 * q - 1
    a4[8, ::2] = 0
    a2[8, 1::2] = 0
    outs = {}
    try:
        for name, ring, fp in (("fp64", 1, 1), ("int", 1, 0), ("generic", 0, 0)):
            ntt.configure("ring", ring)
            ntt.configure("fp64", fp)
            df, di = to_dev(a4), to_dev(a2)
            plan.fwd(df, batch)
            plan.inv(di, batch)
            outs[name] = (to_host(df), to_host(di))
    finally:
        ntt.configure("ring", 1)
        ntt.configure("fp64", 1)
    for name in ("int", "generic"):
        assert np.array_equal(outs["fp64"][0], outs[name][0]), "forward: fp64 vs %s" % name
        assert np.array_equal(outs["fp64"][1], outs[name][1]), "inverse: fp64 vs %s" % name
    for r in (0, 7, 8, batch // 3, batch - 1):
        assert np.array_equal(outs["fp64"][0][r], oracle.fwd(a4[r], q, t.w, t.w_con))
        assert np.array_equal(outs["fp64"][1][r], oracle.inv(a2[r], q, t.n_inv, t.w_inv, t.w_inv_con))
    plan.close()


@pytest.mark.parametrize("bits", [49, 50])
def test_fp64_path_structured_inputs(ntt, oracle, bits):
    """Structured vectors that push sums and differences to their extremes (constant, alternating, half-filled,
    single spikes, top-of-range), through the FP64 ring kernel at N = 2^14, against the oracle."""
    m, N = 14, 1 << 14
    if bits == 49:
        q = 0x1FFFFFC800001
    else:
        q = (1 << 50) - ((1 << 50) - 1) % (2 * N)
        while not oracle.is_prime(q) or q > (1 << 50) - 2048:
            q -= 2 * N
    x = 2
    while True:
        psi = oracle.powmod(x, (q - 1) // (2 * N), q)
        if oracle.powmod(psi, N, q) == q - 1:
            break
        x += 1
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    plan = ntt.Plan.from_psi(N, q, psi)
    idx = np.arange(N)
    rows = []
    for top in (q - 1, 4 * q - 1):
        rows += [np.full(N, top), np.where(idx % 2 == 0, top, 0), np.where(idx < N // 2, top, 0),
                 np.where(idx % 4 < 2, top, 1), np.where((idx >> 7) % 2 == 0, top, 0)]
    for pos in (0, 1, N // 2, N - 1):
        v = np.zeros(N); v[pos] = q - 1; rows.append(v)
    rows.append(idx % q); rows.append((q - 1 - idx) % q)
    a = np.stack(rows).astype(np.uint64)
    batch = a.shape[0]
    d = to_dev(a)
    plan.fwd(d, batch)
    f = to_host(d)
    assert np.array_equal(f, oracle.fwd(a, q, t.w, t.w_con))
    # the same vectors as NTT-domain input of the inverse (contract [0,2q)), and the inverse of their transforms
    b = np.minimum(a, np.uint64(2 * q - 1))
    d = to_dev(b)
    plan.inv(d, batch)
    assert np.array_equal(to_host(d), oracle.inv(b, q, t.n_inv, t.w_inv, t.w_inv_con))
    d = to_dev(f)
    plan.inv(d, batch)
    assert np.array_equal(to_host(d), a % np.uint64(q))
    plan.close()


def test_rns_limbs(ntt, oracle):
    """BASELINE config 3 shape (scaled down): N=2^16, several ~50-bit limbs, each with its own q."""
    N, m, limbs, per = 1 << 16, 16, 3, 2
    qs, q = [], (1 << 50) + 1
    q -= (q - 1) % (2 * N)
    while len(qs) < limbs:
        q -= 2 * N
        if oracle.is_prime(q):
            qs.append(q)
    plans, tabs = [], []
    for q in qs:
        psi = oracle.min_root(N, q)
        plans.append(ntt.Plan.from_psi(N, q, psi))
        tabs.append(CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q)))
    a = np.stack([oracle.uniform(per * N, q, 5 + i).reshape(per, N) for i, q in enumerate(qs)])
    d = to_dev(a)
    ntt.fwd_rns(plans, d, per)
    f = to_host(d)
    for i, t in enumerate(tabs):
        assert np.array_equal(f[i], oracle.fwd(a[i], t.q, t.w, t.w_con))
    ntt.inv_rns(plans, d, per)
    assert np.array_equal(to_host(d), a)
    for p in plans:
        p.close()


def test_negacyclic_polymul(ntt, oracle, golden_synth):
    """BASELINE config 4 shape: c = INTT(NTT(a) o NTT(b)); checked against schoolbook and oracle NTTs."""
    # small ring: exact schoolbook product
    N, q = 256, 7681
    plan = ntt.Plan.from_psi(N, q, 62)
    a, b = oracle.uniform(N, q, 21), oracle.uniform(N, q, 22)
    da, db = to_dev(a), to_dev(b)
    plan.negacyclic_mul(da, da, db, 1)
    assert np.array_equal(to_host(da), oracle.negacyclic_mul(a, b, q))
    plan.close()
    # N = 2^13, 49-bit q: against the oracle pipeline fwd, fwd, pointwise, inv
    s = [x for x in golden_synth if x["m"] == 13][0]
    N, q, batch = 1 << 13, s["q"], 8
    t = CaseTables(oracle, 13, q, s["psi"], s["psi_inv"], s["n_inv"])
    plan = ntt.Plan.from_psi(N, q, s["psi"])
    a = oracle.uniform(batch * N, q, 31).reshape(batch, N)
    b = oracle.uniform(batch * N, q, 32).reshape(batch, N)
    da, db = to_dev(a), to_dev(b)
    dc = to_dev(np.zeros_like(a))
    plan.negacyclic_mul(dc, da, db, batch)
    fa, fb = oracle.fwd(a, q, t.w, t.w_con), oracle.fwd(b, q, t.w, t.w_con)
    prod = oracle.pointwise_mul(fa, fb, q).reshape(batch, N)
    want = oracle.inv(prod, q, t.n_inv, t.w_inv, t.w_inv_con)
    assert np.array_equal(to_host(dc), want)
    # the fused product (FP64 kernel) against the unfused integer pipeline on a large batch, byte for byte
    big = 2048
    a = oracle.uniform(big * N, q, 33).reshape(big, N)
    b = oracle.uniform(big * N, q, 34).reshape(big, N)
    res = {}
    try:
        for name, fp in (("fused", 1), ("unfused", 0)):
            ntt.configure("fp64", fp)
            da, db = to_dev(a), to_dev(b)
            plan.negacyclic_mul(da, da, db, big)
            res[name] = to_host(da)
    finally:
        ntt.configure("fp64", 1)
    assert np.array_equal(res["fused"], res["unfused"])
    plan.close()


def test_host_batch_api_pinned_and_pageable(ntt, oracle, golden_synth):
    torch = _torch()
    s = [x for x in golden_synth if x["m"] == 14][0]
    N, q, batch = 1 << 14, s["q"], 600  # > one pipeline chunk (32 MiB = 256 polynomials)
    t = CaseTables(oracle, 14, q, s["psi"], s["psi_inv"], s["n_inv"])
    plan = ntt.Plan.from_psi(N, q, s["psi"])
    a = oracle.uniform(batch * N, q, 41).reshape(batch, N)
    pageable = a.copy()
    plan.fwd_host(pageable, batch)
    pinned = torch.from_numpy(a.view(np.int64).copy()).pin_memory()
    plan.fwd_host(pinned, batch)
    f = pinned.numpy().view(np.uint64)
    assert np.array_equal(f, pageable)
    for r in (0, 255, 256, 599):
        assert np.array_equal(f[r], oracle.fwd(a[r], q, t.w, t.w_con))
    plan.inv_host(pinned, batch)
    assert np.array_equal(pinned.numpy().view(np.uint64), a)
    plan.close()


def test_host_pipeline_many_chunks(ntt, oracle, golden_synth, monkeypatch):
    """The host-buffer pipeline with far more than 64 chunks (1 MiB chunks, 80 MiB batch; chunk sizes ramp up at the head
    and halve over the tail): every row equals the device-resident result, ragged batch sizes included."""
    import os
    s = [x for x in golden_synth if x["m"] == 13][0]
    N, q = 1 << 13, s["q"]
    monkeypatch.setenv("NTT_B200_PIPE_MIB", "1")               # 16 polynomials of 64 KiB per chunk
    try:
        plan = ntt.Plan.from_psi(N, q, s["psi"])
        for batch in (1283, 1, 17, 32, 33):
            a = oracle.uniform(batch * N, q, 43 + batch).reshape(batch, N)
            h = a.copy()
            plan.fwd_host(h, batch)
            d = to_dev(a)
            plan.fwd(d, batch)
            assert np.array_equal(h, to_host(d).reshape(batch, N)), batch
            plan.inv_host(h, batch)
            assert np.array_equal(h, a), batch
        plan.close()
    finally:
        os.environ["NTT_B200_PIPE_MIB"] = "32"                 # the setting is process-wide: put the default back
        p2 = ntt.Plan.from_psi(N, q, s["psi"])
        x = oracle.uniform(N, q, 5).reshape(1, N)
        p2.fwd_host(x, 1)
        p2.close()


def test_large_n_two_pass_split(ntt, oracle):
    """N = 2^20 (strided passes + chunk kernel): one polynomial against the oracle."""
    m, q = 20, 0x1FFFFFC800001
    N = 1 << m
    x = 2
    while True:
        psi = oracle.powmod(x, (q - 1) // (2 * N), q)
        if oracle.powmod(psi, N, q) == q - 1:
            break
        x += 1
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    plan = ntt.Plan.from_psi(N, q, psi)
    a = oracle.uniform(N, q, 4)
    d = to_dev(a)
    plan.fwd(d, 1)
    assert np.array_equal(to_host(d), oracle.fwd(a, q, t.w, t.w_con))
    plan.inv(d, 1)
    assert np.array_equal(to_host(d), a)
    plan.close()


@pytest.mark.parametrize("m,bits", [(15, 49), (16, 50), (17, 49), (17, 50), (18, 49), (18, 50), (19, 49), (19, 50),
                                    (20, 50), (21, 49)])
def test_fp64_strided_passes_all_radices(ntt, oracle, m, bits):
    """N = 2^15 .. 2^21 in the FP64 range: the stages above a 2^14 chunk run as k_strided_fp passes of radix 2^1 .. 2^5
    (one or two passes; 16-byte and 8-byte variants; both range schedules).  Inputs at the edges of the contracts
    ([0,4q) forward, [0,2q) inverse), three polynomials, every coefficient against the oracle; and the integer strided
    passes (NTT_B200_NO_FP64-style switch) must give the same bytes."""
    N = 1 << m
    top = (1 << bits) - (2048 if bits == 50 else 1024)
    q = (1 << bits) + 1
    q -= (q - 1) % (2 * N)
    while True:
        q -= 2 * N
        if q <= top and oracle.is_prime(q):
            break
    psi = oracle.min_root(N, q)
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    plan = ntt.Plan.from_psi(N, q, psi)
    assert "k_strided_fp" in plan.describe()[0] and "k_strided_fp" in plan.describe(inverse=True)[0], plan.describe()
    a = oracle.uniform(3 * N, 4 * q, 700 + m).reshape(3, N)
    a[0, :8] = [4 * q - 1, 0, 4 * q - 1, 2 * q, q, q - 1, 3 * q + 1, 1]
    a[1, :] = 4 * q - 1
    d = to_dev(a)
    plan.fwd(d, 3)
    f = to_host(d).reshape(3, N)
    assert np.array_equal(f, oracle.fwd_batch(a, q, t.w, t.w_con)), "forward"
    b = oracle.uniform(3 * N, 2 * q, 800 + m).reshape(3, N)
    b[1, :] = 2 * q - 1
    d = to_dev(b)
    plan.inv(d, 3)
    g = to_host(d).reshape(3, N)
    assert np.array_equal(g, oracle.inv_batch(b, q, t.n_inv, t.w_inv, t.w_inv_con, t.n_inv_con)), "inverse"
    try:
        ntt.configure("fp64", 0)
        d = to_dev(a)
        plan.fwd(d, 3)
        assert np.array_equal(to_host(d).reshape(3, N), f)
        d = to_dev(b)
        plan.inv(d, 3)
        assert np.array_equal(to_host(d).reshape(3, N), g)
    finally:
        ntt.configure("fp64", 1)
    plan.close()


def _random_plan_params(oracle, rng, m, bits):
    """A prime q = 1 (mod 2N) of the given width (searched downwards from a random start) and a primitive 2N-th root."""
    N = 1 << m
    lo, hi = 1 << (bits - 1), (1 << bits) - 1
    q = int(rng.integers(lo, hi, dtype=np.uint64))
    q -= (q - 1) % (2 * N)
    while q > lo and not oracle.is_prime(q):
        q -= 2 * N
    assert q > lo
    x = 2
    while True:
        psi = oracle.powmod(x, (q - 1) // (2 * N), q)
        if oracle.powmod(psi, N, q) == q - 1:
            return N, q, psi
        x += 1


@pytest.mark.parametrize("seed", range(24))
def test_random_plans_against_oracle(ntt, oracle, seed):
    """Random (N, modulus width, batch) combinations -- whatever kernel the plan picks (generic, integer ring, FP64
    ring in both range schedules, exact, strided + ring) must agree with the oracle on inputs spanning the lazy
    contracts, with a batch that is not a multiple of anything."""
    rng = np.random.default_rng(1000 + seed)
    m = int(rng.integers(1, 18))
    bits = int(rng.integers(max(m + 2, 12), 63))
    if seed % 4 == 0:
        bits = int(rng.integers(47, 51))          # around the FP64 schedules' boundaries
        m = int(rng.integers(11, 17))
    N, q, psi = _random_plan_params(oracle, rng, m, bits)
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    batch = int(rng.integers(1, 12)) if m > 12 else int(rng.integers(1, 300))
    plan = ntt.Plan.from_psi(N, q, psi)
    top_f = min(4 * q, 1 << 64) - 1
    a4 = oracle.uniform(batch * N, top_f, 77 + seed).reshape(batch, N)
    a2 = oracle.uniform(batch * N, 2 * q, 78 + seed).reshape(batch, N)
    a4[0, :] = top_f - 1
    a2[0, :] = 2 * q - 1
    df, di = to_dev(a4), to_dev(a2)
    plan.fwd(df, batch)
    plan.inv(di, batch)
    f, i = to_host(df), to_host(di)
    for r in sorted({0, batch // 2, batch - 1}):
        assert np.array_equal(f[r], oracle.fwd(a4[r], q, t.w, t.w_con)), "forward, m=%d bits=%d row %d" % (m, bits, r)
        assert np.array_equal(i[r], oracle.inv(a2[r], q, t.n_inv, t.w_inv, t.w_inv_con)), \
            "inverse, m=%d bits=%d row %d" % (m, bits, r)
    plan.inv(df, batch)
    assert np.array_equal(to_host(df), a4 % np.uint64(q)), "round trip, m=%d bits=%d" % (m, bits)
    plan.close()


def test_plans_on_two_devices_in_one_process(ntt, oracle):
    """The library keeps no per-process device state: plans on different GPUs can be used from one process, each on
    its own data and stream, without disturbing the caller's current device (one process per GPU is how bench.py
    scales, but a host application may drive several GPUs itself)."""
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    m, q = 14, 0x1FFFFFC800001
    N = 1 << m
    psi = 20456969886
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    batch = 40
    a = oracle.uniform(batch * N, q, 91).reshape(batch, N)
    plans = [ntt.Plan.from_psi(N, q, psi, device=d) for d in (0, 1)]
    bufs = [torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to("cuda:%d" % d) for d in (0, 1)]
    torch.cuda.set_device(0)
    for rnd in range(2):                       # interleave the devices
        for d in (1, 0):
            plans[d].fwd(bufs[d], batch)
        assert torch.cuda.current_device() == 0
    for d in (0, 1):
        torch.cuda.synchronize(d)
    want1 = oracle.fwd(a[:2], q, t.w, t.w_con)
    want2 = oracle.fwd(want1 % np.uint64(q), q, t.w, t.w_con)
    for d in (0, 1):
        got = bufs[d].cpu().numpy().view(np.uint64)
        assert np.array_equal(got[:2], want2), "device %d" % d
    assert torch.equal(bufs[0].cpu(), bufs[1].cpu())
    for p in plans:
        p.close()


def test_error_behaviour(ntt, oracle, case_tables):
    t = case_tables(0)
    with pytest.raises(ntt.NttError):
        ntt.Plan.from_psi(t.N, t.q, 3)  # not a primitive 2N-th root
    with pytest.raises(ntt.NttError):
        ntt.Plan.from_psi(t.N + 1, t.q, t.psi)  # N not a power of two
    bad = t.w_con.copy()
    bad[5] ^= 1
    with pytest.raises(ntt.NttError):
        ntt.Plan.from_tables(t.N, t.q, t.w, bad)  # companion table inconsistent with q
    fwd_only = ntt.Plan.from_tables(t.N, t.q, t.w, t.w_con)
    d = to_dev(np.zeros(t.N, dtype=np.uint64))
    with pytest.raises(ntt.NttError):
        fwd_only.inv(d, 1)
    fwd_only.fwd(d, 0)  # empty batch is a no-op
    fwd_only.close()


@pytest.mark.parametrize("bits", [50, 58])
def test_maximum_size(ntt, oracle, bits):
    """N = 2^24, the largest size the C-ABI accepts (NTT_B200_MAX_LOGN): two strided passes + the chunk kernel.
    bits = 50: FP64 ring kernel; 58: the exact (Harvey) path.  One polynomial against the oracle, forward over the
    full lazy contract, and the round trip."""
    m = 24
    N = 1 << m
    q = (1 << bits) - ((1 << bits) - 1) % (2 * N)
    while not oracle.is_prime(q) or (bits == 50 and q > (1 << 50) - 2048):
        q -= 2 * N
    x = 2
    while True:
        psi = oracle.powmod(x, (q - 1) // (2 * N), q)
        if oracle.powmod(psi, N, q) == q - 1:
            break
        x += 1
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    plan = ntt.Plan.from_psi(N, q, psi)
    a = oracle.uniform(N, min(4 * q, 1 << 64) - 1, 2424)
    a[:4096] = min(4 * q, 1 << 64) - 2
    d = to_dev(a)
    plan.fwd(d, 1)
    f = to_host(d)
    assert np.array_equal(f, oracle.fwd(a, q, t.w, t.w_con))
    plan.inv(d, 1)
    assert np.array_equal(to_host(d), a % np.uint64(q))
    plan.close()
