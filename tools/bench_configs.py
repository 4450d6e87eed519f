#!/usr/bin/env python
"""Timings of BASELINE configs 3 and 4 on one GPU (they are parity-test cases in tests/, this adds the numbers):
  config 3  CKKS-style RNS batch: N = 2^16, 48 limbs (largest 49-bit primes = 1 mod 2^17), B polynomials per limb
  config 4  negacyclic polynomial multiply: N = 2^13, batch 16384 (fwd x2, pointwise, inverse; product fused)
Results are sanity-checked without the test oracle (round trip, agreement of the FP64 and integer kernels, a
few schoolbook coefficients); bit-exact parity lives in tests/.  python tools/bench_configs.py [rns|polymul]"""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ntt = importlib.import_module("optimized-number-theoretic-transform-implementations_b200")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
# under torchrun the RNS limbs are sharded contiguously across the ranks (no collective on the data path)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sharding = importlib.import_module("optimized-number-theoretic-transform-implementations_b200.sharding")

def timed(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

if which in ("rns", "all"):
    m, limbs, per = 16, 48, 32
    N = 1 << m
    qs, q = [], (1 << 49) + 1
    q -= (q - 1) % (2 * N)
    while len(qs) < limbs:
        q -= 2 * N
        if ntt.is_prime(q) and q <= (1 << 49) - 1024: qs.append(q)
    lb, le = sharding.shard_range(limbs, rank, world)
    all_limbs, qs, limbs = limbs, qs[lb:le], le - lb
    plans, psis = [], []
    for q in qs:
        x = 2
        while True:
            psi = ntt.pow_mod(x, (q - 1) // (2 * N), q)
            if ntt.pow_mod(psi, N, q) == q - 1: break
            x += 1
        psis.append(psi); plans.append(ntt.Plan.from_psi(N, q, psi, device=local))
    rng = np.random.default_rng(2)
    a = np.stack([rng.integers(0, q, size=(per, N), dtype=np.uint64) for q in qs])
    d = torch.from_numpy(a.view(np.int64)).cuda()
    ntt.fwd_rns(plans, d, per); f = d.cpu().numpy().view(np.uint64)
    ntt.configure("fp64", 0)                                   # integer kernels must agree with the FP64 ones
    for l in (0, limbs - 1):
        dl = torch.from_numpy(a[l].view(np.int64)).cuda(); plans[l].fwd(dl, per)
        assert np.array_equal(f[l], dl.cpu().numpy().view(np.uint64)), "RNS limb %d: FP64 and integer kernels differ" % l
    ntt.configure("fp64", 1)
    ntt.inv_rns(plans, d, per); assert np.array_equal(d.cpu().numpy().view(np.uint64), a)
    if world > 1: dist.barrier()
    ms_f = timed(lambda: ntt.fwd_rns(plans, d, per)); ms_i = timed(lambda: ntt.inv_rns(plans, d, per))
    if world > 1:
        tt = torch.tensor([ms_f, ms_i], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms_f, ms_i = tt.tolist()
    n = all_limbs * per
    if rank == 0:
        print(json.dumps({"config": "RNS N=2^16 x %d limbs x %d polys, 49-bit primes, limbs sharded over %d GPU(s)" % (all_limbs, per, world),
                          "fwd_ms": ms_f, "inv_ms": ms_i, "fwd_ntt_per_s": n / ms_f * 1e3, "inv_ntt_per_s": n / ms_i * 1e3,
                          "fwd_frac_hbm_single_pass_per_gpu": n * 2 * N * 8 / (ms_f * 1e-3) / 6537.3e9 / world}))
    for p in plans: p.close()
    del d

if which in ("polymul", "all") and rank == 0:
    m, batch, q, psi = 13, 16384, 0x1FFFFFC800001, 94912374482
    N = 1 << m
    plan = ntt.Plan.from_psi(N, q, psi)
    rng = np.random.default_rng(3)
    a = rng.integers(0, q, size=(batch, N), dtype=np.uint64); b = rng.integers(0, q, size=(batch, N), dtype=np.uint64)
    da, db = torch.from_numpy(a.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()
    dc = torch.empty_like(da)
    plan.negacyclic_mul(dc, da, db, batch); c = dc.cpu().numpy().view(np.uint64)
    for r in (0, 9999):                                        # schoolbook X^N = -1 product, a few coefficients
        ar, br = [int(v) for v in a[r]], [int(v) for v in b[r]]
        for k in (0, 1, N // 2, N - 1):
            want = (sum(ar[i] * br[k - i] for i in range(k + 1)) - sum(ar[i] * br[N + k - i] for i in range(k + 1, N))) % q
            assert int(c[r, k]) == want, "polymul row %d coefficient %d" % (r, k)
    def step():
        plan.negacyclic_mul(da, da, db, batch)   # in place: result in da, db is work space
    ms = timed(step)
    ntt.configure("fp64", 0); ms_int = timed(step); ntt.configure("fp64", 1)
    print(json.dumps({"config": "negacyclic polymul N=2^13 batch %d" % batch, "ms": ms, "polymul_per_s": batch / ms * 1e3,
                      "frac_hbm_3N8": batch * 3 * N * 8 / (ms * 1e-3) / 6537.3e9, "ms_integer_unfused": ms_int}))
    plan.close()

if world > 1:
    dist.destroy_process_group()
