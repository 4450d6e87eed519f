"""CPU tests of the multi-rank host logic (world_size 2 over gloo): contiguous sharding of independent
polynomials / RNS limbs with no data-path collective, and the max-over-ranks timing reduction bench.py uses."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, limbs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module(PKG + ".sharding")
    b, e = sh.shard_range(total, rank, world)
    lb, le = sh.shard_range(limbs, rank, world)
    mine = torch.zeros(total, dtype=torch.int64)
    mine[b:e] = 1
    dist.all_reduce(mine)  # test-only check that shards tile the batch exactly once
    step_ms = 10.0 + 5.0 * rank  # rank 1 is slower: the job time is the max
    out[rank] = dict(cover=bool((mine == 1).all()), n=e - b, limbs=(lb, le),
                     max_ms=sh.reduce_max(step_ms, dist), sum_units=sh.reduce_sum(e - b, dist))
    dist.destroy_process_group()


@pytest.mark.parametrize("total,limbs", [(4096, 48), (4097, 7), (3, 1)])
def test_world2_sharding_and_timing_reduce(total, limbs):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, total, limbs, out), nprocs=world, join=True)
    assert all(out[r]["cover"] for r in range(world))
    assert sum(out[r]["n"] for r in range(world)) == total
    assert abs(out[0]["n"] - out[1]["n"]) <= 1
    assert out[0]["limbs"][1] == out[1]["limbs"][0] and out[1]["limbs"][1] == limbs
    assert out[0]["max_ms"] == out[1]["max_ms"] == 15.0
    assert out[0]["sum_units"] == total


def test_shard_range_properties():
    sh = importlib.import_module(PKG + ".sharding")
    for total in (0, 1, 7, 48, 4096):
        for world in (1, 2, 4, 8):
            edges = [sh.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            if total:
                assert all(sh.limb_owner(l, total, world) == next(r for r, (b, e) in enumerate(edges) if b <= l < e)
                           for l in range(0, total, max(1, total // 5)))
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)


# ---- the exchange step of the distributed transform (BASELINE config 5), on CPU tensors over gloo -------------

def _exchange_worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fs = importlib.import_module(PKG + ".fourstep")
    a = torch.arange(n, dtype=torch.int64) * 7 + 3          # a[e] = 7e + 3: position is recoverable
    cyclic = a[rank::world].clone()
    block = fs.exchange_cyclic_to_blocks(cyclic, world, dist)
    back = fs.exchange_blocks_to_cyclic(block, world, dist)
    n_local = n // world
    out[rank] = dict(block_ok=bool(torch.equal(block, a[rank * n_local:(rank + 1) * n_local])),
                     back_ok=bool(torch.equal(back, cyclic)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [16, 4096])
def test_world2_cyclic_block_exchange(n):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port, n, out), nprocs=world, join=True)
    assert all(out[r]["block_ok"] and out[r]["back_ok"] for r in range(world))
