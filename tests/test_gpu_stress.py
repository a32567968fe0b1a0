"""GPU stress of the multi-kernel / multi-stream step paths (was scripts/stress_follow.py): random batch sizes incl.
multi-wave ones, random swarm and grid sizes, many steps, two envs stepping on two streams at the same time, eager and
graph-replayed -- the follower-kernel shape must equal the in-kernel shapes bit for bit and leave its flags clean."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M(pkg, cuda):
    return pkg.submodule("envs.multiagent")


def test_follower_stress_two_streams(M):
    nat = M.nat
    rnd = random.Random(0)
    keys = ("x", "xa", "grid", "positions", "reward", "done_u8", "episode", "elapsed")
    for trial in range(10):
        E = rnd.choice([1, 3, 37, 300, 1500, 5000])
        N = rnd.choice([160, 176, 200, 256, 300, 512, 700])
        G = rnd.choice([20, 83, 84])
        lim = rnd.choice([2, 5, 128])
        a = M.BatchedSwarmEnv(E, n_locusts=N, grid_size=G, seed=trial, max_episode_steps=lim, binding="ctypes",
                              tuning=nat.TUNE_RASTER_FOLLOW)
        b = M.BatchedSwarmEnv(E, n_locusts=N, grid_size=G, seed=trial, max_episode_steps=lim, binding="ctypes",
                              tuning=rnd.choice([nat.TUNE_RASTER_WARPS, nat.TUNE_RASTER_SELF]))
        a.reset(); b.reset()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        act = torch.randn(E, 10, 2, device="cuda").clamp(-0.9, 0.9)
        torch.cuda.synchronize()
        for t in range(12):
            with torch.cuda.stream(s1):
                a.step(act)
            with torch.cuda.stream(s2):
                b.step(act)
        torch.cuda.synchronize()
        for k in keys:
            assert torch.equal(getattr(a, k), getattr(b, k)), (trial, E, N, G, k)
        assert int(a.work.sum()) == 0, (trial, E, N, G)


def test_two_followers_on_two_streams_concurrently(M):
    """Two independent env batches, both taking the two-kernel step, driven from two streams at once: each call uses the
    side stream / events of ITS caller stream, so neither can wait on the other's fork event."""
    nat = M.nat
    E, N = 600, 256
    envs = [M.BatchedSwarmEnv(E, n_locusts=N, seed=5, max_episode_steps=4, tuning=nat.TUNE_RASTER_FOLLOW) for _ in range(2)]
    ref = M.BatchedSwarmEnv(E, n_locusts=N, seed=5, max_episode_steps=4, tuning=nat.TUNE_RASTER_SELF)
    for e in envs + [ref]:
        e.reset()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    act = torch.randn(E, 10, 2, device="cuda").clamp(-0.9, 0.9)
    torch.cuda.synchronize()
    for t in range(20):
        for e, s in zip(envs, streams):
            with torch.cuda.stream(s):
                e.step(act)
        ref.step(act)
    torch.cuda.synchronize()
    for e in envs:
        for k in ("x", "grid", "positions", "reward", "episode"):
            assert torch.equal(getattr(e, k), getattr(ref, k)), k
        assert int(e.work.sum()) == 0
