"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (container only; /root/reference needed).

    python -m oracle.make_golden

The reference holds no golden vectors of its own, so these fixtures are
outputs of the unmodified reference code (fed_gym/envs/multiagent.py, state_processors.py,
paac/emulator_runner.py) loaded by oracle/ref_loader.py, with its random draws recorded so
that the GPU box (where the reference does not exist) can replay the same inputs.
All arrays are float64 unless noted; actions are float32-representable.
"""
import os

import numpy as np

from . import ref_loader as rl
from . import swarm_oracle as so

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sparse(grid):
    idx = np.argwhere(grid != 0).astype(np.int16)
    return idx, grid[grid != 0]


def clipped_actions(rs, steps, A=10):
    a = rs.normal(size=(steps, A, 2)).astype(np.float32).astype(np.float64)
    for t in range(steps):
        so.clip_actions_(a[t])
    return a.astype(np.float32).astype(np.float64)


def trajectory(n_locusts, draws, actions, snap_steps, proc):
    """Run the reference env from injected draws; record states/grids at snap_steps, all rewards."""
    env = rl.make_reference_env(n_locusts)
    st = rl.reference_reset_injected(env, *draws)
    out = {"x": [], "xa": [], "grid_idx": [], "grid_val": [], "pos": [], "reward": [], "done": []}

    def snap(state):
        g = proc.process_state(state)
        i, v = sparse(g)
        out["x"].append(state[0].copy()); out["xa"].append(state[1].copy())
        out["grid_idx"].append(i); out["grid_val"].append(v); out["pos"].append(proc.positions.copy())

    if 0 in snap_steps:
        snap(st)
    for t in range(actions.shape[0]):
        st, r, d, _ = env.step(actions[t])
        out["reward"].append(r); out["done"].append(d)
        if (t + 1) in snap_steps:
            snap(st)
    assert env.t == env.N_BURN_IN      # Q1: the noise row index never advances after reset
    return out


def pack(prefix, d, tr, snap_steps):
    d[prefix + "snap_steps"] = np.array(snap_steps)
    d[prefix + "x"] = np.stack(tr["x"]); d[prefix + "xa"] = np.stack(tr["xa"])
    d[prefix + "pos"] = np.stack(tr["pos"])
    d[prefix + "reward"] = np.array(tr["reward"]); d[prefix + "done"] = np.array(tr["done"])
    for k, (i, v) in enumerate(zip(tr["grid_idx"], tr["grid_val"])):
        d[prefix + "grid_idx_%d" % k] = i
        d[prefix + "grid_val_%d" % k] = v


def c1_1000():
    """BASELINE.json configs[0] / SURVEY 8d C1 as specified: the reference SwarmEnv (N=80, seed 192 = Swarm-eval-v0), 1000
    steps of a random policy = 7 full 128-step episodes + 104 steps with gym's TimeLimit emulated by hand (the raw class has
    none, fed_gym/__init__.py:21-33), actions RandomState(0).normal clipped like transform_actions_for_env.  Every
    reset re-seeds the global numpy RNG (multiagent.py:47-48), so all 8 episodes start from the same state.  Kept: the
    actions, all rewards and done flags, the state at every episode start and after the first 16 steps of every
    episode, and the reference's own steps/s on this container's CPU."""
    import time
    env = rl.make_reference_env(80, seed=192)
    rs = np.random.RandomState(0)
    acts = clipped_actions(rs, 1000)
    st = env.reset()
    starts_x, starts_xa, x16, xa16 = [st[0].copy()], [st[1].copy()], [], []
    rewards, dones, ep_of_step = np.zeros(1000), np.zeros(1000, dtype=bool), np.zeros(1000, dtype=np.int32)
    elapsed, ep, spent = 0, 0, 0.0
    for t in range(1000):
        t0 = time.perf_counter()
        st, r, d, _ = env.step(acts[t])
        spent += time.perf_counter() - t0
        elapsed += 1
        done = bool(d) or elapsed >= 128                # gym 0.9.4 TimeLimit._step
        rewards[t], dones[t], ep_of_step[t] = r, done, ep
        if elapsed == 16:
            x16.append(st[0].copy()); xa16.append(st[1].copy())
        if done:
            assert env.t == env.N_BURN_IN
            t0 = time.perf_counter()
            st = env.reset()
            spent += time.perf_counter() - t0
            elapsed, ep = 0, ep + 1
            starts_x.append(st[0].copy()); starts_xa.append(st[1].copy())
    assert ep == 7 and elapsed == 104 and dones.sum() == 7
    np.savez_compressed(os.path.join(OUT, "c1_1000_seed192_n80.npz"), actions=acts, reward=rewards, done=dones,
                        episode=ep_of_step, start_x=np.stack(starts_x), start_xa=np.stack(starts_xa),
                        x16=np.stack(x16), xa16=np.stack(xa16),
                        ref_steps_per_s=np.array(1000.0 / spent), ref_seconds=np.array(spent))
    print("c1_1000: reference %.1f steps/s (%.2f s incl. 7 resets)" % (1000.0 / spent, spent))


def main():
    os.makedirs(OUT, exist_ok=True)
    ma, sp = rl.load_reference()
    proc = sp.SwarmStateProcessor(grid_size=84)

    # ---- C1: Swarm-eval-v0 (seed 192), N=80, one full 128-step episode ------------------------
    np.random.seed(192)
    draws = so.draw_reset_numpy(np.random, 80)
    env = rl.make_reference_env(80, seed=192)
    st = env.reset()
    chk = rl.reference_reset_injected(rl.make_reference_env(80), *draws)
    assert np.array_equal(st[0], chk[0]) and np.array_equal(st[1], chk[1]), "draw order mismatch"
    acts = clipped_actions(np.random.RandomState(0), 128)
    snaps = [0, 1, 16, 64, 128]
    d = dict(x0=draws[0], xa0=draws[1], burn=draws[2], agent_noise=draws[3], particle_noise=draws[4], actions=acts)
    pack("", d, trajectory(80, draws, acts, snaps, proc), snaps)
    np.savez_compressed(os.path.join(OUT, "c1_seed192_n80.npz"), **d)

    # ---- N=64 and N=256: reset + 16 steps, every state kept (teacher forcing) -----------------
    for N in (64, 256):
        rs = np.random.RandomState(1000 + N)
        draws = so.draw_reset_numpy(rs, N)
        acts = clipped_actions(rs, 16)
        snaps = list(range(17))
        d = dict(x0=draws[0], xa0=draws[1], burn=draws[2], agent_noise=draws[3], particle_noise=draws[4], actions=acts)
        pack("", d, trajectory(N, draws, acts, snaps, proc), snaps)
        np.savez_compressed(os.path.join(OUT, "traj16_n%d.npz" % N), **d)

    # ---- rasteriser edge cases through the reference's process_state --------------------------
    cases = []
    rs = np.random.RandomState(7)
    for G in (84, 20):
        p = sp.SwarmStateProcessor(grid_size=G)
        for k in range(6):
            N = [80, 64, 256, 5, 80, 33][k]
            x = rs.rand(N, 2) * [2.0, 1.0] + [rs.rand() * 8, 0.0]
            xa = rs.rand(10, 2) * [4.0, 8.0] + [x[:, 0].mean() - 2.0, -1.0]
            if k >= 1:
                m = so.sequential_mean_x(x, xa)
                # place points exactly on edges: solve by fixed-point so the mean stays the same
                ex = so.box_edges(m - 1.5, m + 1.5, G)
                ey = so.box_edges(0.0, 6.0, G)
                xa[0] = [ex[0], ey[0]]; xa[1] = [ex[-1], ey[-1]]; xa[2] = [ex[G // 2], ey[G // 3]]
                xa[3] = [np.nextafter(ex[-1], np.inf), 1.0]; xa[4] = [np.nextafter(ex[0], -np.inf), 7.0]
                x[0] = [ex[G - 1], ey[G - 1]]; x[1] = [ex[1], 0.0]; x[2] = [ex[3], 6.0]
                x[3] = [ex[5], np.nextafter(6.0, 7.0)]
            if k == 4:
                x[:, :] = x[0]            # all locusts in ONE cell (count == N)
            if k == 5:
                x[:, 1] = 0.0             # everybody grounded
            g = p.process_state([x, xa])
            i, v = sparse(g)
            cases.append((G, x, xa, i, v, p.positions.copy()))
    d = {"n_cases": np.array(len(cases))}
    for k, (G, x, xa, i, v, pos) in enumerate(cases):
        d["G_%d" % k] = np.array(G); d["x_%d" % k] = x; d["xa_%d" % k] = xa
        d["grid_idx_%d" % k] = i; d["grid_val_%d" % k] = v; d["pos_%d" % k] = pos
    np.savez_compressed(os.path.join(OUT, "raster_cases.npz"), **d)

    # ---- SwarmRunner statics -------------------------------------------------------------------
    Runner = rl.load_reference_runner()
    rs = np.random.RandomState(11)
    a = (rs.normal(size=(40, 2)) * 1.2).astype(np.float32)
    a[0] = [3, 4]; a[1] = [0.6, 0.8]; a[2] = [0, 0]; a[3] = [1, 0]
    a_in = a.copy()
    a64 = a.astype(np.float64)
    ret = Runner.transform_actions_for_env(a64)
    assert ret is a64
    env = rl.make_reference_env(80, seed=5)
    g = proc.process_state(env.reset())
    loc = np.array(Runner.get_local_states(g, proc.positions))
    hot = np.argwhere(loc[:, :, :, 2] != 0).astype(np.int16)
    gi, gv = sparse(g)
    np.savez_compressed(os.path.join(OUT, "runner_statics.npz"), clip_in=a_in, clip_out=a64,
                        ls_grid_idx=gi, ls_grid_val=gv, ls_pos=proc.positions.copy(), ls_hot=hot,
                        ls_shape=np.array(loc.shape))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "c1":      # only the 1000-step C1 fixture (leaves the others byte-identical)
        os.makedirs(OUT, exist_ok=True)
        c1_1000()
    else:
        main()
        c1_1000()
