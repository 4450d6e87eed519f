"""GPU tests of the round-2 entry points: the lazy forward, the fused forward-multiply-inverse (device and host
buffers), squaring through negacyclic_mul, and the multi-GPU C API (device list in, batch sharded, no collective).
All comparisons are bit-exact against the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import CaseTables

pytestmark = pytest.mark.gpu
Q49 = 0x1FFFFFC800001


def _torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _dev(a, device=0):
    torch = _torch()
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to("cuda:%d" % device)


def _host(t):
    return t.cpu().numpy().view(np.uint64)


def _setup(oracle, m, q=Q49):
    N = 1 << m
    psi = oracle.min_root(N, q)
    return N, psi, CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))


def _q50(oracle, N):
    q = (1 << 50) - ((1 << 50) - 1) % (2 * N)
    while not oracle.is_prime(q) or q > (1 << 50) - 2048:
        q -= 2 * N
    return q


@pytest.mark.parametrize("m,bits", [(14, 49), (14, 50), (13, 49), (12, 50), (16, 49), (10, 49), (14, 55)])
def test_lazy_forward_contract(ntt, oracle, m, bits):
    """ntt_b200_fwd_lazy_batch: output within the reference's lazy window [0,4q) and equal to fwd_ntt_ref_harvey
    after reduce_4q_to_q (tests/test_correctness.c:267-269).  On the FP64 ring kernel the values really are
    unreduced (some >= q), i.e. the entry point is not an alias of the reduced one."""
    N = 1 << m
    if bits == 49:
        q = Q49
    elif bits == 50:
        q = _q50(oracle, N)
    else:
        q = (1 << 55) - ((1 << 55) - 1) % (2 * N)
        while not oracle.is_prime(q):
            q -= 2 * N
    N, psi, t = _setup(oracle, m, q)
    batch = 64
    a = oracle.uniform(batch * N, 4 * q, 21).reshape(batch, N)
    plan = ntt.Plan.from_psi(N, q, psi)
    d = _dev(a)
    plan.fwd_lazy(d, batch)
    lazy = _host(d)
    want = oracle.fwd_batch(a, q, t.w, t.w_con)
    assert (lazy < 4 * q).all()
    assert np.array_equal(lazy % np.uint64(q), want)
    if bits <= 50 and m >= 12:
        assert (lazy < 2 * q).all() and (lazy >= q).any(), "the FP64 ring kernel should skip the sign correction"
    plan.close()


@pytest.mark.parametrize("m,bits", [(14, 49), (13, 49), (14, 50), (16, 49), (11, 49), (14, 55)])
def test_fwd_mul_inv_device(ntt, oracle, m, bits):
    """a <- INTT(NTT(a) .* m) with ONE multiplier broadcast over the batch, against the oracle pipeline."""
    N = 1 << m
    if bits == 49:
        q = Q49
    elif bits == 50:
        q = _q50(oracle, N)
    else:
        q = (1 << 55) - ((1 << 55) - 1) % (2 * N)
        while not oracle.is_prime(q):
            q -= 2 * N
    N, psi, t = _setup(oracle, m, q)
    batch = 37
    a = oracle.uniform(batch * N, q, 5).reshape(batch, N)
    mult = oracle.uniform(N, q, 6)                              # NTT-domain multiplier, canonical residues
    fa = oracle.fwd_batch(a, q, t.w, t.w_con)
    prod = oracle.pointwise_mul(fa, np.ascontiguousarray(np.broadcast_to(mult, (batch, N))), q).reshape(batch, N)
    want = oracle.inv_batch(prod, q, t.n_inv, t.w_inv, t.w_inv_con)
    plan = ntt.Plan.from_psi(N, q, psi)
    d, dm = _dev(a), _dev(mult)
    plan.fwd_mul_inv(d, dm, batch)
    assert np.array_equal(_host(d), want)
    # NULL multiplier: plain round trip
    d = _dev(a)
    plan.fwd_mul_inv(d, None, batch)
    assert np.array_equal(_host(d), a)
    # host-buffer form (pinned and pageable), chunked pipeline: make the batch span several chunks
    torch = _torch()
    big = max(batch, (96 << 20) // (N * 8) + 3)
    ab = oracle.uniform(big * N, q, 8).reshape(big, N)
    fb = oracle.fwd_batch(ab, q, t.w, t.w_con)
    pb = oracle.pointwise_mul(fb, np.ascontiguousarray(np.broadcast_to(mult, (big, N))), q).reshape(big, N)
    wb = oracle.inv_batch(pb, q, t.n_inv, t.w_inv, t.w_inv_con)
    pinned = torch.from_numpy(ab.copy().view(np.int64)).pin_memory()
    plan.fwd_mul_inv_host(pinned, dm, big)
    assert np.array_equal(pinned.numpy().view(np.uint64), wb)
    pageable = ab.copy()
    plan.fwd_mul_inv_host(pageable, dm, big)
    assert np.array_equal(pageable, wb)
    # the separate host calls still agree with the oracle on the same multi-chunk batch
    h = ab.copy()
    plan.fwd_host(h, big)
    assert np.array_equal(h, fb)
    plan.inv_host(h, big)
    assert np.array_equal(h, ab)
    plan.close()


@pytest.mark.parametrize("m,bits", [(13, 49), (14, 49), (10, 49), (13, 55)])
def test_negacyclic_square(ntt, oracle, m, bits):
    """d_a == d_b (squaring) used to return garbage: now one forward, a pointwise square, the inverse."""
    N = 1 << m
    q = Q49
    if bits == 55:
        q = (1 << 55) - ((1 << 55) - 1) % (2 * N)
        while not oracle.is_prime(q):
            q -= 2 * N
    N, psi, t = _setup(oracle, m, q)
    batch = 5
    a = oracle.uniform(batch * N, q, 9).reshape(batch, N)
    fa = oracle.fwd_batch(a, q, t.w, t.w_con)
    want = oracle.inv_batch(oracle.pointwise_mul(fa, fa, q).reshape(batch, N), q, t.n_inv, t.w_inv, t.w_inv_con)
    plan = ntt.Plan.from_psi(N, q, psi)
    d = _dev(a)
    plan.negacyclic_mul(d, d, d, batch)                         # c = a = b
    assert np.array_equal(_host(d), want)
    d, c = _dev(a), _dev(np.zeros_like(a))
    plan.negacyclic_mul(c, d, d, batch)                         # separate output
    assert np.array_equal(_host(c), want)
    if m <= 10:                                                 # schoolbook cross-check of the first polynomial
        assert np.array_equal(want[0], oracle.negacyclic_mul(a[0], a[0], q))
    plan.close()


def test_multi_gpu_c_api(ntt, oracle):
    """ntt_b200_multi_*: every visible GPU gets a contiguous shard; device-resident and host-buffer forms;
    results equal the oracle row for row whatever the device count (1 on the test box, 2..8 on a bigger one)."""
    torch = _torch()
    ndev = torch.cuda.device_count()
    m = 13
    N, psi, t = _setup(oracle, m)
    q = Q49
    batch = 8 * ndev + 5                                        # ragged: shards differ by one
    a = oracle.uniform(batch * N, q, 31).reshape(batch, N)
    want = oracle.fwd_batch(a, q, t.w, t.w_con)
    mp = ntt.MultiPlan(N, q, psi, list(range(ndev)))
    shards = [mp.shard(batch, i) for i in range(ndev)]
    assert sum(c for _, c in shards) == batch and shards[0][0] == 0
    assert all(shards[i][0] + shards[i][1] == shards[i + 1][0] for i in range(ndev - 1))
    bufs = [_dev(a[f:f + c], i) for i, (f, c) in enumerate(shards)]
    mp.fwd(bufs, [c for _, c in shards])
    mp.sync()
    got = np.concatenate([_host(b) for b in bufs])
    assert np.array_equal(got, want)
    mp.inv(bufs, [c for _, c in shards])
    mp.sync()
    assert np.array_equal(np.concatenate([_host(b) for b in bufs]), a)
    # host-buffer forms: one host thread per device
    h = a.copy()
    mp.fwd_host(h, batch)
    assert np.array_equal(h, want)
    mp.inv_host(h, batch)
    assert np.array_equal(h, a)
    mult = oracle.uniform(N, q, 32)
    dms = [_dev(mult, i) for i in range(ndev)]
    prod = oracle.pointwise_mul(want, np.ascontiguousarray(np.broadcast_to(mult, (batch, N))), q).reshape(batch, N)
    want2 = oracle.inv_batch(prod, q, t.n_inv, t.w_inv, t.w_inv_con)
    mp.fwd_mul_inv_host(h, dms, batch)
    assert np.array_equal(h, want2)
    mp.close()


def test_rns_limbs_sharded_over_devices(ntt, oracle):
    """ntt_b200_fwd_rns_multi / inv: limb l on device l mod ndev, each with its own modulus (config 3 shape)."""
    torch = _torch()
    ndev = torch.cuda.device_count()
    m, limbs, per = 13, 6, 3
    N = 1 << m
    qs, q = [], (1 << 50) + 1
    while len(qs) < limbs:
        q -= 2 * N
        if oracle.is_prime(q) and q <= (1 << 50) - 2048:
            qs.append(q)
    plans, bufs, wants, ins = [], [], [], []
    for l, ql in enumerate(qs):
        dev = l * ndev // limbs                                  # contiguous runs of limbs per device
        psi = oracle.min_root(N, ql)
        t = CaseTables(oracle, m, ql, psi, oracle.invmod(psi, ql), oracle.invmod(N, ql))
        a = oracle.uniform(per * N, ql, 40 + l).reshape(per, N)
        plans.append(ntt.Plan.from_psi(N, ql, psi, device=dev))
        bufs.append(_dev(a, dev))
        ins.append(a)
        wants.append(oracle.fwd_batch(a, ql, t.w, t.w_con))
    ntt.rns_multi(plans, bufs, per)
    for i in range(ndev):
        ntt.device_sync(i)
    for l in range(limbs):
        assert np.array_equal(_host(bufs[l]), wants[l]), "limb %d" % l
    ntt.rns_multi(plans, bufs, per, inverse=True)
    for i in range(ndev):
        ntt.device_sync(i)
    for l in range(limbs):
        assert np.array_equal(_host(bufs[l]), ins[l]), "limb %d round trip" % l
    for p in plans:
        p.close()


@pytest.mark.parametrize("m", [12, 14, 16])
def test_unordered_contract(ntt, oracle, m):
    """The reference's own check for its unordered variant (tests/test_correctness.c:179-209): repair the order, then
    memcmp with fwd_ntt_ref_harvey.  Plus what a pointwise consumer relies on: inv_unordered(fwd_unordered(a) .* fwd_unordered(b))
    is the negacyclic product."""
    N, psi, t = _setup(oracle, m)
    q, batch = Q49, 3
    a = oracle.uniform(batch * N, q, 61).reshape(batch, N)
    b = oracle.uniform(batch * N, q, 62).reshape(batch, N)
    plan = ntt.Plan.from_psi(N, q, psi)
    perm = np.array([plan.unordered_index(i) for i in range(N)], dtype=np.int64)
    assert sorted(perm.tolist()) == list(range(N)) and plan.unordered_index(N) == 2**64 - 1
    da, db = _dev(a), _dev(b)
    plan.fwd_unordered(da, batch)
    plan.fwd_unordered(db, batch)
    fa = _host(da)
    fixed = np.empty_like(fa)
    fixed[:, perm] = fa                                         # fix_a_order
    want_a = oracle.fwd_batch(a, q, t.w, t.w_con)
    assert np.array_equal(fixed, want_a)
    plan.pointwise_mul(da, da, db, batch)
    plan.inv_unordered(da, batch)
    fb = oracle.fwd_batch(b, q, t.w, t.w_con)
    want = oracle.inv_batch(oracle.pointwise_mul(want_a, fb, q).reshape(batch, N), q, t.n_inv, t.w_inv, t.w_inv_con)
    assert np.array_equal(_host(da), want)
    plan.close()


@pytest.mark.parametrize("bits,batch", [(49, 1), (49, 37), (50, 5), (49, 300)])
def test_one_kernel_polymul(ntt, oracle, bits, batch):
    """ntt_b200_negacyclic_mul_batch at N = 2^13 runs as ONE kernel (csrc/ntt_polymul_fp.cuh): every product equals
    the oracle pipeline (forward x2, pointwise product, inverse), the composed path gives identical bytes, and the
    aliasing forms (c = a, c = b, a = b) work.  batch 300 > 148 CTAs: several pairs per CTA through the slot ring."""
    m, N = 13, 1 << 13
    q = Q49 if bits == 49 else _q50(oracle, N)
    N, psi, t = _setup(oracle, m, q)
    a = oracle.uniform(batch * N, 4 * q, 71).reshape(batch, N)      # forward contract [0,4q)
    b = oracle.uniform(batch * N, 4 * q, 72).reshape(batch, N)
    a[0, :] = 4 * q - 1
    b[0, :] = 4 * q - 1
    fa, fb = oracle.fwd_batch(a, q, t.w, t.w_con), oracle.fwd_batch(b, q, t.w, t.w_con)
    want = oracle.inv_batch(oracle.pointwise_mul(fa, fb, q).reshape(batch, N), q, t.n_inv, t.w_inv, t.w_inv_con)
    want_sq = oracle.inv_batch(oracle.pointwise_mul(fa, fa, q).reshape(batch, N), q, t.n_inv, t.w_inv, t.w_inv_con)
    plan = ntt.Plan.from_psi(N, q, psi)
    da, db, dc = _dev(a), _dev(b), _dev(np.zeros_like(a))
    plan.negacyclic_mul(dc, da, db, batch)
    assert np.array_equal(_host(dc), want), "one-kernel product differs from the oracle pipeline"
    assert np.array_equal(_host(da), a) and np.array_equal(_host(db), b), "the fused kernel leaves its operands alone"
    da2 = _dev(a)
    plan.negacyclic_mul(da2, da2, db, batch)                      # c aliases a
    assert np.array_equal(_host(da2), want)
    db2 = _dev(b)
    plan.negacyclic_mul(db2, da, db2, batch)                      # c aliases b
    assert np.array_equal(_host(db2), want)
    plan.negacyclic_mul(dc, da, da, batch)                        # squaring
    assert np.array_equal(_host(dc), want_sq)
    try:                                                          # the composed path: identical bytes
        ntt.configure("polymul", 0)
        da3, db3, dc3 = _dev(a % np.uint64(q)), _dev(b % np.uint64(q)), _dev(np.zeros_like(a))
        plan.negacyclic_mul(dc3, da3, db3, batch)
        assert np.array_equal(_host(dc3), want)
    finally:
        ntt.configure("polymul", 1)
    if batch == 1:                                                # schoolbook cross-check, independent of any NTT
        assert np.array_equal(want[0], oracle.negacyclic_mul(a[0] % np.uint64(q), b[0] % np.uint64(q), q))
    plan.close()


@pytest.mark.parametrize("m,limbs,per,bits", [(16, 5, 3, 50), (14, 4, 9, 49), (15, 48, 2, 50), (17, 3, 1, 49),
                                               (14, 48, 1, 49), (14, 2, 200, 49), (15, 7, 33, 50)])
def test_rns_single_launch_matches_oracle_and_per_limb_path(ntt, oracle, m, limbs, per, bits):
    """RNS batches of N >= 2^14 in the FP64 range run as ONE launch per kernel over all limbs (k_ring_fp<.., MULTI>,
    k_strided_multi): every limb equals the oracle with its own modulus, the inverse returns the input, and the
    launch-per-limb path (NTT_B200_NO_RING... switched by configure("fp64", 0) -> integer kernels) gives the same bytes."""
    N = 1 << m
    top = (1 << bits) - (2048 if bits == 50 else 1024)
    qs, q = [], (1 << bits) + 1
    q -= (q - 1) % (2 * N)
    while len(qs) < limbs:
        q -= 2 * N
        if q <= top and oracle.is_prime(q):
            qs.append(q)
    plans, tabs, ins = [], [], []
    for l, ql in enumerate(qs):
        psi = oracle.min_root(N, ql)
        plans.append(ntt.Plan.from_psi(N, ql, psi))
        tabs.append(CaseTables(oracle, m, ql, psi, oracle.invmod(psi, ql), oracle.invmod(N, ql)))
        ins.append(oracle.uniform(per * N, 4 * ql, 90 + l).reshape(per, N))     # forward contract [0,4q)
    a = np.stack(ins)
    d = _dev(a)
    ntt.fwd_rns(plans, d, per)
    f = _host(d).reshape(limbs, per, N)
    for l in range(limbs):
        assert np.array_equal(f[l], oracle.fwd_batch(ins[l], qs[l], tabs[l].w, tabs[l].w_con)), "limb %d forward" % l
    ntt.inv_rns(plans, d, per)
    back = _host(d).reshape(limbs, per, N)
    for l in range(limbs):
        assert np.array_equal(back[l], ins[l] % np.uint64(qs[l])), "limb %d round trip" % l
    try:                                        # the integer kernels, one launch per limb: identical bytes
        ntt.configure("fp64", 0)
        d2 = _dev(a)
        ntt.fwd_rns(plans, d2, per)
        assert np.array_equal(_host(d2).reshape(limbs, per, N), f)
    finally:
        ntt.configure("fp64", 1)
    for p in plans:
        p.close()
