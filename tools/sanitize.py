"""Small end-to-end run of every kernel path for compute-sanitizer (memcheck / synccheck / initcheck)."""
import importlib, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ntt = importlib.import_module("optimized-number-theoretic-transform-implementations_b200")
q = 0x1FFFFFC800001
for m, batch in ((12, 9), (13, 5), (14, 3), (15, 2), (8, 3), (14, 330), (13, 700)):  # the last two: several polynomials per CTA
    N = 1 << m
    psi = ntt.min_primitive_root(N, q) if m <= 8 else None
    if psi is None:
        x = 2
        while True:
            psi = ntt.pow_mod(x, (q - 1) // (2 * N), q)
            if ntt.pow_mod(psi, N, q) == q - 1: break
            x += 1
    plan = ntt.Plan.from_psi(N, q, psi)
    a = np.random.default_rng(m).integers(0, q, size=(batch, N), dtype=np.uint64)
    for ring, fp in ((1, 1), (1, 0), (0, 0)):
        ntt.configure("ring", ring); ntt.configure("fp64", fp)
        d = torch.from_numpy(a.view(np.int64)).cuda()
        plan.fwd(d, batch); plan.inv(d, batch); torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().view(np.uint64), a), (m, ring, fp)
    ntt.configure("ring", 1); ntt.configure("fp64", 1)
    d2 = torch.from_numpy(a.view(np.int64)).cuda(); d3 = torch.from_numpy(a.view(np.int64)).cuda()
    plan.negacyclic_mul(d2, d2, d3, batch); torch.cuda.synchronize()
    plan.close()
# round-2 paths: lazy forward, forward-multiply-inverse with a broadcast operand, one launch over all RNS limbs
for m, limbs, per in ((14, 3, 5), (16, 4, 2)):
    N = 1 << m
    qs, qq = [], (1 << 49) + 1
    qq -= (qq - 1) % (2 * N)
    while len(qs) < limbs:
        qq -= 2 * N
        if ntt.is_prime(qq) and qq <= (1 << 49) - 1024: qs.append(qq)
    plans = [ntt.Plan.from_psi(N, ql, ntt.min_primitive_root(N, ql)) for ql in qs]
    a = np.stack([np.random.default_rng(l).integers(0, ql, size=(per, N), dtype=np.uint64) for l, ql in enumerate(qs)])
    d = torch.from_numpy(a.view(np.int64)).cuda()
    ntt.fwd_rns(plans, d, per); ntt.inv_rns(plans, d, per); torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint64), a), ("rns", m)
    mult = torch.from_numpy(np.random.default_rng(9).integers(0, qs[0], size=N, dtype=np.uint64).view(np.int64)).cuda()
    d = torch.from_numpy(a[0].view(np.int64)).cuda()
    plans[0].fwd_lazy(d, per); plans[0].fwd_mul_inv(d, mult, per); torch.cuda.synchronize()
    for p in plans: p.close()
# tail stages fused with the exchange (all ranks' slices on this GPU)
import ctypes as C
fs = importlib.import_module("optimized-number-theoretic-transform-implementations_b200.fourstep")
m, G = 16, 4
N = 1 << m
x = 2
while True:
    psi = ntt.pow_mod(x, (q - 1) // (2 * N), q)
    if ntt.pow_mod(psi, N, q) == q - 1: break
    x += 1
a = np.random.default_rng(3).integers(0, q, size=N, dtype=np.uint64)
parts = [fs.DistributedNtt(N, q, psi, r, G) for r in range(G)]
slices = [torch.from_numpy(np.ascontiguousarray(a[p::G]).view(np.int64)).cuda() for p in range(G)]
ptrs = (C.c_void_p * G)(*[t.data_ptr() for t in slices])
blocks = [torch.empty(N // G, dtype=torch.int64, device="cuda") for _ in range(G)]
for p in range(G): parts[p].local.fwd(slices[p], 1)
for r in range(G): parts[r].full.fwd_tail_gather(ptrs, blocks[r], 2, r, batch=1)
for r in range(G): parts[r].full.inv_tail_scatter(ptrs, blocks[r], 2, r, batch=1)
for p in range(G): parts[p].local.inv(slices[p], 1)
torch.cuda.synchronize()
back = np.empty(N, dtype=np.uint64)
for p in range(G): back[p::G] = slices[p].cpu().numpy().view(np.uint64)
assert np.array_equal(back, a)
print("sanitize run ok")
