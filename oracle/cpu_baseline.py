"""Time the CPU oracle (the NumPy port of the reference env + rasteriser) on the host cores.

TEST/BENCH INFRASTRUCTURE: used only by bench.py's ``cpu_baseline`` leg and ``--impl reference``.
The reference itself is Python and lives outside the repo (it cannot travel to the GPU box), so
the timed code is this repo's restatement ("kind": "port").  Like the reference's runner pool
(fed_gym/agents/paac/runners.py:14-19) the env batch is split over P worker processes; each
worker owns its envs and runs reset + the step/rasterise loop on them.
"""
import multiprocessing as mp
import os
import time

import numpy as np


def _worker(args):
    n_envs, n_locusts, steps, warmup, seed, grid = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from . import swarm_oracle as so
    rs = np.random.RandomState(seed)
    ds = [so.draw_reset_numpy(rs, n_locusts) for _ in range(n_envs)]
    draws = [np.stack([d[i] for d in ds]) for i in range(5)]
    x, xa = so.reset_injected(*draws)
    na, nx = draws[3][:, 10], draws[4][:, 10]

    def one_step():
        a = rs.normal(size=(n_envs, so.N_AGENTS, 2))
        so.clip_actions_(a.reshape(-1, 2))
        so.step(x, xa, a, na, nx)
        for e in range(n_envs):
            so.rasterize(x[e], xa[e], grid)

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    return time.perf_counter() - t0


def time_port(n_locusts, steps, warmup=1, envs_per_proc=2, procs=None, grid=84, seed=0):
    """Returns dict(env_steps_per_s, seconds, procs, envs, steps).  All host cores by default."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    jobs = [(envs_per_proc, n_locusts, steps, warmup, seed + 1000 * i, grid) for i in range(procs)]
    with ctx.Pool(procs) as pool:
        times = pool.map(_worker, jobs)
    wall = max(times)
    envs = procs * envs_per_proc
    return dict(env_steps_per_s=envs * steps / wall, seconds=wall, procs=procs, envs=envs, steps=steps)
