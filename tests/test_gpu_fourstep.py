"""GPU tests of the distributed (multi-GPU) transform: BASELINE config 5, one polynomial spread over G ranks with
a single exchange step.  On one GPU every rank is emulated in turn with the same kernels and index arithmetic;
with >= 2 GPUs visible the NCCL path runs as well (spawned processes, 127.0.0.1 rendezvous)."""
import importlib
import os
import socket

import numpy as np
import pytest

from conftest import PKG, CaseTables

pytestmark = pytest.mark.gpu
Q49 = 0x1FFFFFC800001


def _root(oracle, N, q):
    x = 2
    while True:
        c = oracle.powmod(x, (q - 1) // (2 * N), q)
        if oracle.powmod(c, N, q) == q - 1:
            return c
        x += 1


@pytest.mark.parametrize("m,world", [(16, 2), (18, 4), (20, 8), (22, 8)])
def test_distributed_transform_emulated_on_one_gpu(ntt, oracle, m, world):
    fourstep = importlib.import_module(PKG + ".fourstep")
    N, q = 1 << m, Q49
    psi = _root(oracle, N, q)
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    a = oracle.uniform(N, q, 4)
    got, back = fourstep.emulate_forward_single_gpu(N, q, psi, a, world)
    assert np.array_equal(got, oracle.fwd(a, q, t.w, t.w_con)), "distributed forward differs from the oracle"
    assert np.array_equal(back, a), "distributed inverse(forward(a)) != a"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, m, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fourstep = importlib.import_module(PKG + ".fourstep")
    from oracle.pyoracle import Oracle
    orc = Oracle()
    N, q = 1 << m, Q49
    psi = _root(orc, N, q)
    a = orc.uniform(N, q, 4)
    plan = fourstep.DistributedNtt(N, q, psi, rank, world, device=rank)
    sl = torch.from_numpy(np.ascontiguousarray(a[rank::world]).view(np.int64)).cuda()
    block = plan.forward(sl, dist)
    torch.cuda.synchronize()
    fwd_block = block.cpu().numpy().view(np.uint64).copy()
    back = plan.inverse(block, dist)
    torch.cuda.synchronize()
    out[rank] = (fwd_block, back.cpu().numpy().view(np.uint64).copy())
    plan.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [18, 22])
def test_distributed_transform_nccl(ntt, oracle, m):
    import torch
    import torch.multiprocessing as mp
    world = min(8, torch.cuda.device_count())
    world = 1 << (world.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    N, q = 1 << m, Q49
    psi = _root(oracle, N, q)
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    a = oracle.uniform(N, q, 4)
    want = oracle.fwd(a, q, t.w, t.w_con)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), m, out), nprocs=world, join=True)
    got = np.concatenate([out[r][0] for r in range(world)])
    assert np.array_equal(got, want)
    back = np.empty(N, dtype=np.uint64)
    for r in range(world):
        back[r::world] = out[r][1]
    assert np.array_equal(back, a)


# ---- exchange fused into the tail kernels (peer loads / stores instead of a collective) -------------------

def _prime_below(oracle, bits, N):
    q = (1 << bits) - ((1 << bits) - 1) % (2 * N)
    while not oracle.is_prime(q):
        q -= 2 * N
    return q


@pytest.mark.parametrize("m,world,qbits", [(16, 2, 49), (18, 4, 49), (20, 8, 49), (14, 32, 49), (14, 32, 55),
                                           (22, 32, 55), (16, 16, 56)])
def test_peer_gather_scatter_emulated_on_one_gpu(ntt, oracle, m, world, qbits):
    """ntt_b200_fwd_tail_gather / ntt_b200_inv_tail_scatter with every rank's slice on the same GPU (the "peer"
    pointers are local): same result as the reference transform, no all-to-all and no interleave copy.
    qbits = 55 / 56: the inverse tail's own lazy bounds (the size-N plan schedules a renormalisation inside the
    tail's five stages for q above about 2^51.7, which a single register network would skip)."""
    import ctypes as C
    import torch
    fourstep = importlib.import_module(PKG + ".fourstep")
    N, G = 1 << m, world
    q = Q49 if qbits == 49 else _prime_below(oracle, qbits, N)
    g = G.bit_length() - 1
    psi = _root(oracle, N, q)
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    a = oracle.uniform(N, 4 * q, 6)                       # forward contract [0,4q)
    a[: N // 64] = 4 * q - 1
    parts = [fourstep.DistributedNtt(N, q, psi, r, G) for r in range(G)]
    slices = [torch.from_numpy(np.ascontiguousarray(a[p::G]).view(np.int64)).cuda() for p in range(G)]
    ptrs = (C.c_void_p * G)(*[s.data_ptr() for s in slices])
    for p in range(G):
        parts[p].local.fwd(slices[p], 1)
    blocks = [torch.empty(N // G, dtype=torch.int64, device="cuda") for _ in range(G)]
    for r in range(G):
        parts[r].full.fwd_tail_gather(ptrs, blocks[r], g, r)
    got = torch.cat(blocks).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, oracle.fwd(a, q, t.w, t.w_con)), "gather + tail differs from the oracle"
    for s in slices:
        s.fill_(-1)
    for r in range(G):
        parts[r].full.inv_tail_scatter(ptrs, blocks[r], g, r)
    back = np.empty(N, dtype=np.uint64)
    for p in range(G):
        parts[p].local.inv(slices[p], 1)
        back[p::G] = slices[p].cpu().numpy().view(np.uint64)
    assert np.array_equal(back, a % np.uint64(q)), "tail + scatter + local inverse != input"
    for d in parts:
        d.close()


def _peer_worker(rank, world, port, m, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fourstep = importlib.import_module(PKG + ".fourstep")
    from oracle.pyoracle import Oracle
    orc = Oracle()
    N, q = 1 << m, Q49
    psi = _root(orc, N, q)
    a = orc.uniform(N, q, 4)
    plan = fourstep.FusedDistributedNtt(N, q, psi, rank, world, rank, dist)
    block = torch.empty(N // world, dtype=torch.int64, device="cuda")
    plan.px.load_slice(a[rank::world])
    fwd_block = None
    for it in range(3):                                   # forward / inverse pairs reuse the buffers and the flags
        plan.forward(block)
        if it == 0:
            torch.cuda.synchronize()
            fwd_block = block.cpu().numpy().view(np.uint64).copy()
        plan.inverse(block)
    graph = plan.capture_pair(block)                      # the same pair replayed as a CUDA graph
    for it in range(2):
        graph.replay()
    torch.cuda.synchronize()
    del graph
    out[rank] = (fwd_block, plan.px.read_slice(), plan.px.timed_out())
    plan.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [18, 22])
def test_distributed_transform_peer_memory(ntt, oracle, m):
    """One process per GPU, slices mapped through CUDA IPC, GPU-side flag barrier: forward blocks equal the
    reference transform and three forward/inverse round trips return the input."""
    import torch
    import torch.multiprocessing as mp
    world = min(8, torch.cuda.device_count())
    world = 1 << (world.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    N, q = 1 << m, Q49
    psi = _root(oracle, N, q)
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    a = oracle.uniform(N, q, 4)
    want = oracle.fwd(a, q, t.w, t.w_con)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_worker, args=(world, _free_port(), m, out), nprocs=world, join=True)
    assert not any(out[r][2] for r in range(world)), "peer barrier timed out"
    got = np.concatenate([out[r][0] for r in range(world)])
    assert np.array_equal(got, want)
    back = np.empty(N, dtype=np.uint64)
    for r in range(world):
        back[r::world] = out[r][1]
    assert np.array_equal(back, a)


@pytest.mark.parametrize("m,world,batch", [(16, 4, 3), (20, 8, 5), (14, 32, 2)])
def test_peer_gather_scatter_batched_emulated(ntt, oracle, m, world, batch):
    """ntt_b200_fwd_tail_gather_batch / inv_tail_scatter_batch: `batch` transforms per launch, slices and blocks
    stored one after the other; every polynomial equals the reference transform and the round trip is exact."""
    import ctypes as C
    import torch
    fourstep = importlib.import_module(PKG + ".fourstep")
    N, G, q = 1 << m, world, Q49
    g = G.bit_length() - 1
    psi = _root(oracle, N, q)
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    a = oracle.uniform(batch * N, 4 * q, 16).reshape(batch, N)
    parts = [fourstep.DistributedNtt(N, q, psi, r, G) for r in range(G)]
    slices = [torch.from_numpy(np.ascontiguousarray(a[:, p::G]).view(np.int64)).cuda() for p in range(G)]   # [batch][N/G]
    ptrs = (C.c_void_p * G)(*[s.data_ptr() for s in slices])
    for p in range(G):
        parts[p].local.fwd(slices[p], batch)
    blocks = [torch.empty(batch * (N // G), dtype=torch.int64, device="cuda") for _ in range(G)]
    for r in range(G):
        parts[r].full.fwd_tail_gather(ptrs, blocks[r], g, r, batch=batch)
    got = torch.stack([b.view(batch, N // G) for b in blocks], dim=1).reshape(batch, N).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, oracle.fwd_batch(a, q, t.w, t.w_con)), "batched gather + tail differs from the oracle"
    for s in slices:
        s.fill_(-1)
    for r in range(G):
        parts[r].full.inv_tail_scatter(ptrs, blocks[r], g, r, batch=batch)
    back = np.empty((batch, N), dtype=np.uint64)
    for p in range(G):
        parts[p].local.inv(slices[p], batch)
        back[:, p::G] = slices[p].view(batch, N // G).cpu().numpy().view(np.uint64)
    assert np.array_equal(back, a % np.uint64(q)), "batched tail + scatter + local inverse != input"
    for d in parts:
        d.close()


def _peer_batch_worker(rank, world, port, m, batch, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fourstep = importlib.import_module(PKG + ".fourstep")
    from oracle.pyoracle import Oracle
    orc = Oracle()
    N, q = 1 << m, Q49
    psi = _root(orc, N, q)
    a = orc.uniform(batch * N, q, 24).reshape(batch, N)
    plan = fourstep.FusedDistributedNtt(N, q, psi, rank, world, rank, dist, batch=batch)
    block = torch.empty(batch * (N // world), dtype=torch.int64, device="cuda")
    plan.px.load_slice(np.ascontiguousarray(a[:, rank::world]).reshape(-1))
    plan.forward(block)
    torch.cuda.synchronize()
    fwd_block = block.cpu().numpy().view(np.uint64).copy()
    plan.inverse(block)
    plan.forward(block)
    plan.inverse(block)
    plan.check()
    out[rank] = (fwd_block, plan.px.read_slice())
    plan.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("m,batch", [(18, 4), (22, 3)])
def test_distributed_transform_peer_memory_batched(ntt, oracle, m, batch):
    """The batched exchange on real GPUs (CUDA IPC, GPU-side barrier): every polynomial of the batch equals the
    reference transform, two round trips return the input."""
    import torch
    import torch.multiprocessing as mp
    world = min(8, torch.cuda.device_count())
    world = 1 << (world.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    N, q = 1 << m, Q49
    psi = _root(oracle, N, q)
    t = CaseTables(oracle, m, q, psi, oracle.invmod(psi, q), oracle.invmod(N, q))
    a = oracle.uniform(batch * N, q, 24).reshape(batch, N)
    want = oracle.fwd_batch(a, q, t.w, t.w_con)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_batch_worker, args=(world, _free_port(), m, batch, out), nprocs=world, join=True)
    got = np.stack([out[r][0].reshape(batch, N // world) for r in range(world)], axis=1).reshape(batch, N)
    assert np.array_equal(got, want)
    back = np.empty((batch, N), dtype=np.uint64)
    for r in range(world):
        back[:, r::world] = out[r][1].reshape(batch, N // world)
    assert np.array_equal(back, a)
