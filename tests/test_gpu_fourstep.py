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
