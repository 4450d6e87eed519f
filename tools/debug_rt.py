import importlib, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ntt = importlib.import_module("optimized-number-theoretic-transform-implementations_b200")
from oracle.pyoracle import Oracle
o = Oracle()
N, q, psi = 1 << 14, 0x1FFFFFC800001, 20456969886
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
plan = ntt.Plan.from_psi(N, q, psi)
a = o.uniform(batch * N, q, 1).reshape(batch, N)
for trial in range(3):
    d = torch.from_numpy(a.view(np.int64)).cuda()
    plan.fwd(d, batch); torch.cuda.synchronize()
    f = d.cpu().numpy().view(np.uint64).copy()
    plan.inv(d, batch); torch.cuda.synchronize()
    r = d.cpu().numpy().view(np.uint64)
    bad = np.argwhere(r != a)
    print("trial", trial, "mismatches", len(bad))
    if len(bad):
        rows = np.unique(bad[:, 0]); print(" rows", rows[:20], "n rows", len(rows))
        for (i, j) in bad[:8]:
            diff = (int(r[i, j]) - int(a[i, j])) % q
            print("  ", i, j, hex(int(r[i, j])), hex(int(a[i, j])), "diff mod q", diff, "r>=q", int(r[i,j]) >= q)
        cols = bad[:, 1]; print(" col min/max", cols.min(), cols.max(), "distinct cols", len(np.unique(cols)))

w, wc = o.tables(N, q, psi)
psi_inv = o.invmod(psi, q); wi, wic = o.tables(N, q, psi_inv); n_inv = o.invmod(N, q)
rows = [151, 206, 0]
sub = a[rows].copy()
d = torch.from_numpy(sub.view(np.int64)).cuda()
plan.fwd(d, len(rows)); torch.cuda.synchronize()
f = d.cpu().numpy().view(np.uint64).copy()
fo = o.fwd(sub, q, w, wc)
print("forward mismatches vs oracle per row", [(int((f[i] != fo[i]).sum())) for i in range(len(rows))])
d2 = torch.from_numpy(fo.view(np.int64)).cuda()
plan.inv(d2, len(rows)); torch.cuda.synchronize()
r2 = d2.cpu().numpy().view(np.uint64)
print("inverse(oracle fwd) mismatches per row", [(int((r2[i] != sub[i]).sum())) for i in range(len(rows))])
bad = np.argwhere(f != fo)
for (i, j) in bad[:6]:
    print("  fwd bad", rows[i], j, hex(int(f[i, j])), hex(int(fo[i, j])), (int(f[i,j]) - int(fo[i,j])) % q)
