"""CPU tests: the oracle is pinned to the reference (golden fixtures generated from the reference build,
and the live reference build where oracle/_ref is present), and the product's host-side table builders
reproduce the reference tables bit for bit."""
import numpy as np
import pytest

from conftest import edge_inputs

FAST_CASES = list(range(19))


@pytest.mark.parametrize("idx", FAST_CASES)
def test_oracle_matches_golden(oracle, golden_cases, case_tables, idx):
    g, t = golden_cases[idx], case_tables(idx)
    h = lambda v: "%016x" % oracle.fnv(v)
    assert (g["m"], g["q"]) == (t.m, t.q)
    # tables: calc_w / calc_w_con restatement
    assert h(t.w) == g["w_fnv"] and h(t.w_con) == g["w_con_fnv"]
    assert h(t.w_inv) == g["w_inv_fnv"] and h(t.w_inv_con) == g["w_inv_con_fnv"]
    assert t.n_inv_con == g["n_inv_con"]
    # structural facts of the tables
    assert t.w[0] == 1 and t.w[t.N // 2] == t.psi and t.w_inv[t.N // 2] == t.psi_inv
    assert int(t.w[1]) * int(t.w[1]) % t.q == t.q - 1
    a = oracle.uniform(t.N, t.q, 0x5EED0000 + idx)
    assert h(a) == g["in_fnv"]
    f = oracle.fwd(a, t.q, t.w, t.w_con)
    assert h(f) == g["fwd_fnv"]
    assert h(oracle.fwd_lazy(a, t.q, t.w, t.w_con)) == g["fwd_lazy_fnv"]
    assert np.array_equal(oracle.inv(f, t.q, t.n_inv, t.w_inv, t.w_inv_con), a)
    a4 = oracle.uniform(t.N, 4 * t.q, 0x4A2F0000 + idx)
    assert h(a4) == g["in4q_fnv"] and h(oracle.fwd(a4, t.q, t.w, t.w_con)) == g["fwd4q_fnv"]
    a2 = oracle.uniform(t.N, 2 * t.q, 0x2A2F0000 + idx)
    assert h(a2) == g["in2q_fnv"] and h(oracle.inv(a2, t.q, t.n_inv, t.w_inv, t.w_inv_con)) == g["inv2q_fnv"]
    if idx % 3 == 0:
        for name, v in edge_inputs(t.N, t.q).items():
            assert h(oracle.fwd(v, t.q, t.w, t.w_con)) == g["edges"][name]["fwd"]
            assert h(oracle.inv(v, t.q, t.n_inv, t.w_inv, t.w_inv_con)) == g["edges"][name]["inv"]


def test_oracle_case0_full_vectors(oracle, golden_case0):
    g = golden_case0
    N, q = 1 << g["m"], g["q"]
    w, wc = oracle.tables(N, q, g["psi"])
    assert w.tolist() == g["w"] and wc.tolist() == g["w_con"]
    wi, wic = oracle.tables(N, q, g["psi_inv"])
    assert wi.tolist() == g["w_inv"] and wic.tolist() == g["w_inv_con"]
    a = np.array(g["a"], dtype=np.uint64)
    assert oracle.fwd(a, q, w, wc).tolist() == g["fwd"]
    assert oracle.fwd_lazy(a, q, w, wc).tolist() == g["fwd_lazy"]
    # the butterfly network computes the transform it is supposed to: out[i] = a(psi^(2*bitrev(i)+1))
    assert oracle.fwd_definition(a, q, g["psi"]).tolist() == g["fwd"]


def test_oracle_parameters_are_the_fixture_rule(oracle, golden_cases, golden_synth):
    """psi is the smallest primitive 2N-th root, w_inv and n_inv its inverses (tests/test_cases.h:113-142)."""
    for g in golden_cases:
        if g["m"] > 14 and g["q"] > (1 << 40):
            continue  # the exhaustive minimum search is slow for the biggest cases; covered by the others
        N = 1 << g["m"]
        assert oracle.is_prime(g["q"]) and (g["q"] - 1) % (2 * N) == 0
        assert oracle.min_root(N, g["q"]) == g["w"]
        assert oracle.invmod(g["w"], g["q"]) == g["w_inv"]
        assert oracle.invmod(N, g["q"]) == g["n_inv"]
    for s in golden_synth:
        assert oracle.powmod(s["psi"], 1 << s["m"], s["q"]) == s["q"] - 1
        assert oracle.invmod(s["psi"], s["q"]) == s["psi_inv"]


def test_oracle_against_live_reference(oracle, reference, golden_cases, case_tables):
    """Where the reference build is present (oracle/_ref), compare byte for byte, lazy values included."""
    if not reference.available:
        pytest.skip("oracle/_ref not built here")
    assert [c["q"] for c in reference.cases()] == [g["q"] for g in golden_cases]
    for idx in (0, 2, 7, 12, 13, 15, 17, 18):
        t = case_tables(idx)
        w, wc = reference.tables(t.N, t.q, t.psi)
        assert np.array_equal(w, t.w) and np.array_equal(wc, t.w_con)
        a = oracle.uniform(t.N, 4 * t.q, 900 + idx)
        assert np.array_equal(reference.fwd_lazy(a, t.q, w, wc), oracle.fwd_lazy(a, t.q, w, wc))
        assert np.array_equal(reference.fwd(a, t.q, w, wc), oracle.fwd(a, t.q, w, wc))
        assert np.array_equal(reference.fwd_seal(a % np.uint64(t.q), t.q, w, wc), oracle.fwd(a, t.q, w, wc))
        b = oracle.uniform(t.N, 2 * t.q, 901 + idx)
        assert np.array_equal(reference.inv(b, t.q, t.n_inv, t.w_inv, t.w_inv_con),
                              oracle.inv(b, t.q, t.n_inv, t.w_inv, t.w_inv_con))


def test_oracle_polymul_definition(oracle):
    N, q, psi = 64, 7681, None
    for cand in range(2, q):
        if oracle.powmod(cand, N, q) == q - 1:
            psi = cand
            break
    w, wc = oracle.tables(N, q, psi)
    wi, wic = oracle.tables(N, q, oracle.invmod(psi, q))
    a, b = oracle.uniform(N, q, 1), oracle.uniform(N, q, 2)
    prod = oracle.pointwise_mul(oracle.fwd(a, q, w, wc), oracle.fwd(b, q, w, wc), q)
    c = oracle.inv(prod, q, oracle.invmod(N, q), wi, wic)
    assert np.array_equal(c, oracle.negacyclic_mul(a, b, q))


# ---- product host logic (no GPU needed): the C replacements of pre_compute.h ----------------------------

@pytest.mark.parametrize("idx", [0, 1, 6, 9, 12, 13, 16, 18])
def test_host_table_builders_match_reference_tables(ntt, oracle, golden_cases, idx):
    g = golden_cases[idx]
    N, q = 1 << g["m"], g["q"]
    h = lambda v: "%016x" % oracle.fnv(v)
    w = ntt.calc_w(g["w"], N, q)
    assert h(w) == g["w_fnv"]
    assert h(ntt.calc_w_con(w, q)) == g["w_con_fnv"]
    wi = ntt.calc_w(g["w_inv"], N, q)
    assert h(wi) == g["w_inv_fnv"] and h(ntt.calc_w_con(wi, q)) == g["w_inv_con_fnv"]
    assert ntt.calc_ninv_con(g["n_inv"], q) == g["n_inv_con"]
    assert ntt.inv_mod(N, q) == g["n_inv"] and ntt.inv_mod(g["w"], q) == g["w_inv"]
    assert ntt.is_prime(q)
    if g["m"] <= 14 and q < (1 << 40):
        assert ntt.min_primitive_root(N, q) == g["w"]


def test_host_math_helpers(ntt, oracle):
    assert ntt.lib.ntt_b200_bit_rev_idx(1, 14) == 1 << 13
    for i in (0, 1, 5, 1000, 16383):
        assert ntt.lib.ntt_b200_bit_rev_idx(i, 14) == oracle.L.oracle_bitrev(i, 14)
    q = 0x1FFFFFC800001
    assert ntt.is_prime(q) and not ntt.is_prime(q + 2) and not ntt.is_prime(1)
    assert ntt.pow_mod(3, q - 1, q) == 1
    assert ntt.min_primitive_root(1 << 14, q) == 20456969886
    # 52-bit word size (the IFMA tables of the reference) goes through the same builder
    w = ntt.calc_w(62, 256, 7681)
    assert np.array_equal(ntt.calc_w_con(w, 7681, 52), (w.astype(object) * (1 << 52) // 7681).astype(np.uint64))
