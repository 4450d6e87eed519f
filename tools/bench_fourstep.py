#!/usr/bin/env python
"""Timing of BASELINE config 5: one forward+inverse NTT of size N = 2^22 spread over the ranks of a torchrun job
(NCCL all-to-all over NVLink).  python -m torch.distributed.run --nproc-per-node G tools/bench_fourstep.py"""
import importlib, json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "optimized-number-theoretic-transform-implementations_b200"
ntt = importlib.import_module(PKG); fs = importlib.import_module(PKG + ".fourstep")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = int(sys.argv[1]) if len(sys.argv) > 1 else 22
N, q = 1 << m, 0x1FFFFFC800001
x = 2
while True:
    psi = ntt.pow_mod(x, (q - 1) // (2 * N), q)
    if ntt.pow_mod(psi, N, q) == q - 1: break
    x += 1
plan = fs.DistributedNtt(N, q, psi, rank, world, device=local)
a = np.random.default_rng(5).integers(0, q, size=N, dtype=np.uint64)
sl0 = torch.from_numpy(np.ascontiguousarray(a[rank::world]).view(np.int64)).cuda()
d = dist if world > 1 else None
def step(sl):
    blk = plan.forward(sl, d)
    return plan.inverse(blk, d)
for _ in range(3): out = step(sl0.clone())
assert torch.equal(out, sl0), "round trip failed"
steps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1: dist.barrier()
torch.cuda.synchronize(); e0.record()
sl = sl0.clone()
for _ in range(steps): sl = step(sl)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(ms.item())
    print(json.dumps({"config": "N=2^%d forward+inverse over %d GPU(s), cyclic<->block all-to-all" % (m, world), "ms_per_pair": ms,
                      "single_direction_ntt_per_s": 2e3 / ms, "algorithmic_GBps": 2 * 2 * N * 8 / (ms * 1e-3) / 1e9,
                      "hbm_bound_ntt_per_s_per_gpu": 6537.3e9 / (2 * N * 8)}))
plan.close()
if world > 1:
    # the same pair with the exchange fused into the tail kernels (peer loads/stores over NVLink, GPU-side barrier)
    fused = fs.FusedDistributedNtt(N, q, psi, rank, world, local, dist)
    fused.px.load_slice(a[rank::world])
    block = torch.empty(N // world, dtype=torch.int64, device="cuda")
    for _ in range(3):
        fused.forward(block); fused.inverse(block)
    torch.cuda.synchronize()
    assert np.array_equal(fused.px.read_slice(), a[rank::world]) and not fused.px.timed_out(), "fused round trip failed"
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        fused.forward(block); fused.inverse(block)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms.item())
        print(json.dumps({"config": "N=2^%d forward+inverse over %d GPU(s), exchange fused into the tail kernels (peer memory)" % (m, world),
                          "ms_per_pair": ms, "single_direction_ntt_per_s": 2e3 / ms}))
    # and the pair captured once as a CUDA graph (eight short launches: the host launch path is the bottleneck)
    graph = fused.capture_pair(block)
    for _ in range(3): graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(fused.px.read_slice(), a[rank::world]) and not fused.px.timed_out(), "graph round trip failed"
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(steps): graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms.item())
        print(json.dumps({"config": "N=2^%d forward+inverse over %d GPU(s), fused exchange, pair replayed as a CUDA graph" % (m, world),
                          "ms_per_pair": ms, "single_direction_ntt_per_s": 2e3 / ms}))
    del graph
    fused.close()
    dist.destroy_process_group()
