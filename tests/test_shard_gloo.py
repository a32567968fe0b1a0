"""world_size-2 gloo tests (CPU) of the host-side sharding logic.

What can be checked without a GPU: the partition of global env ids, that the per-shard Philox
draw tables (oracle/philox.py restates the device layout) concatenate to the unsharded table
bit for bit, and the max/sum reductions bench.py uses for its timing.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import philox as ph


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, n_locusts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import golds_rl_gym_b200 as pkg
    sh = pkg.submodule("sharding")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, n = sh.shard_envs(total, world, rank)
        ids = sh.global_env_ids(total, world, rank)
        # the shard's reset draws, keyed by GLOBAL env id (csrc/swarm_philox.cuh layout)
        x0 = np.stack([ph.uniform_pairs(1234, int(e), 0, ph.STREAM_X0, n_locusts) for e in ids])
        nz = np.stack([ph.normal_pairs(1234, int(e), 0, ph.STREAM_NOISE_X, 10, n_locusts) for e in ids])
        allx = sh.gather_shards(torch.as_tensor(x0))
        allnz = sh.gather_shards(torch.as_tensor(nz))
        allids = sh.gather_shards(ids)
        gflat = torch.full((1000,), float(rank + 1))
        sh.allreduce_mean_(gflat)
        assert torch.allclose(gflat, torch.full((1000,), (1 + world) / 2.0))
        slow = sh.max_over_ranks(1.0 + rank)
        tot = sh.sum_over_ranks(float(n))
        if rank == 0:
            q.put(dict(first=first, n=n, ids=allids.numpy(), x0=allx.numpy(), nz=allnz.numpy(), slow=slow, tot=tot))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_is_a_partition_and_rng_is_shard_invariant():
    world, total, N = 2, 6, 16
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = q.get()
    assert np.array_equal(got["ids"], np.arange(total))
    assert got["slow"] == 2.0 and got["tot"] == float(total)
    ref_x0 = np.stack([ph.uniform_pairs(1234, e, 0, ph.STREAM_X0, N) for e in range(total)])
    ref_nz = np.stack([ph.normal_pairs(1234, e, 0, ph.STREAM_NOISE_X, 10, N) for e in range(total)])
    assert np.array_equal(got["x0"], ref_x0)        # bit-exact: same global ids -> same streams
    assert np.array_equal(got["nz"], ref_nz)


def test_shard_envs_refuses_uneven_split(pkg):
    sh = pkg.submodule("sharding")
    assert sh.shard_envs(4096, 8, 3) == (1536, 512)
    assert sh.shard_envs(32, 1, 0) == (0, 32)
    with pytest.raises(ValueError):
        sh.shard_envs(10, 4, 0)         # np.split raises for the reference too (runners.py:66)
    with pytest.raises(ValueError):
        sh.shard_envs(8, 2, 2)
