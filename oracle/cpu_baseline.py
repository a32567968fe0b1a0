"""Time the CPU side of the swarm hot path on the host cores.

TEST/BENCH INFRASTRUCTURE: used only by bench.py's ``cpu_baseline`` leg and ``--impl reference``.

  time_reference   the UNMODIFIED reference (fed_gym/envs/multiagent.py SwarmEnv + fed_gym/agents/
                   state_processors.py SwarmStateProcessor.process_state, staged under oracle/_ref by
                   oracle/make_ref.py and loaded through oracle/ref_loader.py)  -> "kind": "reference"
  time_port        this repo's vectorised NumPy restatement (oracle/swarm_oracle.py), ~10x faster than the
                   reference's per-locust Python loop                               -> "kind": "port"

Like the reference's runner pool (fed_gym/agents/paac/runners.py:14-19, SURVEY.md 8d "CPU baseline beside
it") the env batch is split over P worker processes, OMP_NUM_THREADS=1; each worker owns its envs and runs
reset + the step/process_state loop on them with clipped N(0,1) actions.
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


def _ref_worker(args):
    n_envs, n_locusts, steps, warmup, seed, grid = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from . import ref_loader as rl
    from . import swarm_oracle as so
    _, sp = rl.load_reference()
    np.random.seed(seed)
    envs = [rl.make_reference_env(n_locusts) for _ in range(n_envs)]
    proc = sp.SwarmStateProcessor(grid_size=grid)
    for env in envs:
        env.reset()
    rs = np.random.RandomState(seed + 1)

    def one_step():
        for env in envs:
            a = rs.normal(size=(env.N_AGENTS, 2))
            so.clip_actions_(a)                      # emulator_runner.py:113-118
            state, _, _, _ = env.step(a)
            proc.process_state(state)

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    return time.perf_counter() - t0


def reference_staged():
    from . import ref_loader as rl
    return rl.reference_available()


def time_reference(n_locusts, steps, warmup=1, envs_per_proc=1, procs=None, grid=84, seed=0):
    """The reference's own SwarmEnv.step + SwarmStateProcessor.process_state on all host cores.
    Returns dict(env_steps_per_s, seconds, procs, envs, steps)."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    jobs = [(envs_per_proc, n_locusts, steps, warmup, seed + 1000 * i, grid) for i in range(procs)]
    with ctx.Pool(procs) as pool:
        times = pool.map(_ref_worker, jobs)
    wall = max(times)
    envs = procs * envs_per_proc
    return dict(env_steps_per_s=envs * steps / wall, seconds=wall, procs=procs, envs=envs, steps=steps)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    import platform
    return platform.processor() or platform.machine()


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
