"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle / golden vectors.

Tolerances (BASELINE.json north_star): particle states and rewards within rtol 1e-5 in FP32
(metric: max|a-b| <= 1e-5 * max|b| per env, see _util.rel_err); occupancy grids, positions, done
flags, reset indices and Philox integers bit-exact.
"""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import philox as ph
from oracle import swarm_oracle as so

from _util import dense_grid, draws_of, golden, rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-5            # north_star tolerance for FP32 particle states / rewards
STEP_TOL = 5e-6        # teacher-forced single-step bound actually asserted (expected ~1e-7..1e-6)
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


@pytest.fixture(scope="module")
def M(pkg, cuda):
    return pkg.submodule("envs.multiagent")


def stack_draws(seeds, N):
    ds = [so.draw_reset_numpy(np.random.RandomState(s), N) for s in seeds]
    return [np.stack([d[i] for d in ds]) for i in range(5)]


def clipped(rs, shape):
    a = rs.normal(size=shape).astype(np.float32).astype(np.float64)
    so.clip_actions_(a.reshape(-1, 2))
    return a.astype(np.float32)


def to_dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def load_state(env, x, xa, na, nx):
    env.x.copy_(to_dev(x)); env.xa.copy_(to_dev(xa))
    env.noise_a.copy_(to_dev(na)); env.noise_x.copy_(to_dev(nx))
    env._was_reset = True


# --------------------------------------------------------------------------------------- dynamics
@pytest.mark.parametrize("mode", ["fast", "precise"])
@pytest.mark.parametrize("N", [64, 80, 256])
@pytest.mark.parametrize("E", [1, 32])
def test_step_teacher_forced_16(M, E, N, mode):
    """Every one of 16 steps starts from the ORACLE state: x, xa, v, reward, done vs oracle."""
    seeds = [7000 + 13 * e + N for e in range(E)]
    draws = stack_draws(seeds, N)
    x, xa = so.reset_injected(*draws)
    na, nx = draws[3][:, 10], draws[4][:, 10]
    env = M.BatchedSwarmEnv(E, n_locusts=N, max_episode_steps=0, math_mode=mode, auto_reset=False, rasterize=False)
    v_out = torch.zeros(E, N, 2, dtype=torch.float32, device="cuda")
    rs = np.random.RandomState(99)
    worst = 0.0
    for t in range(16):
        a = clipped(rs, (E, 10, 2))
        load_state(env, x, xa, na, nx)
        (gx, gxa), r, d, _ = env.step(to_dev(a), v_out=v_out)
        x_old = x.copy()
        rew, done = so.step(x, xa, a.astype(np.float64), na, nx)   # advances the oracle in place
        v_ref, _ = so.forces(x_old, xa)                            # old x, NEW xa (multiagent.py:39)
        gx, gxa, r, d = gx.cpu().numpy(), gxa.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        gv = v_out.cpu().numpy()
        assert np.array_equal(gxa, xa), "agent update is FP64 and must be exact"
        for e in range(E):
            err = rel_err(gx[e], x[e])
            worst = max(worst, err)
            assert err <= STEP_TOL, (t, e, err)
            assert rel_err(gv[e], v_ref[e]) <= RTOL, (t, e)
        assert np.all(np.abs(r - rew) <= RTOL * np.abs(rew))
        assert np.array_equal(d, done)
    assert worst <= STEP_TOL


@pytest.mark.parametrize("mode", ["fast", "precise"])
@pytest.mark.parametrize("name", ["traj16_n64.npz", "traj16_n256.npz", "c1_seed192_n80.npz"])
def test_forces_match_oracle(M, name, mode):
    """swarm_forces (v_calculate) on golden reference states: v within 1e-5 of max|v|, reward rtol 1e-5."""
    d = golden(name)
    E, n = d["x"].shape[0], d["x"].shape[1]
    env = M.BatchedSwarmEnv(E, n_locusts=n, math_mode=mode, rasterize=False)
    v, r = env.forces(to_dev(d["x"]), to_dev(d["xa"]))
    v_ref, r_ref = so.forces(d["x"], d["xa"])
    v, r = v.cpu().numpy(), r.cpu().numpy()
    for e in range(E):
        assert rel_err(v[e], v_ref[e]) <= RTOL
    assert np.all(np.abs(r - r_ref) <= RTOL * np.abs(r_ref))
    assert (r < 0).all()


@pytest.mark.parametrize("mode", ["fast", "precise"])
@pytest.mark.parametrize("N", [1, 31, 33, 96, 100, 128, 192, 320, 448, 512, 513, 1100])
def test_forces_and_step_every_tiling(M, N, mode):
    """Every force-kernel shape: 32-wide tiles (odd / even tile counts), 64-wide super-tiles (1, 2, 3,
    5, 7, 8 of them), ordered pairs with 2 and 4 targets per thread, ragged last tiles.  v_calculate and one
    full step against the oracle on a post-burn-in state."""
    E = 3
    draws = stack_draws([4200 + N + e for e in range(E)], N)
    x, xa = so.reset_injected(*draws)
    na, nx = draws[3][:, 10], draws[4][:, 10]
    env = M.BatchedSwarmEnv(E, n_locusts=N, max_episode_steps=0, math_mode=mode, auto_reset=False, rasterize=True)
    v, r = env.forces(to_dev(x), to_dev(xa))
    v_ref, r_ref = so.forces(x, xa)
    v, r = v.cpu().numpy(), r.cpu().numpy()
    for e in range(E):
        assert rel_err(v[e], v_ref[e]) <= RTOL, (N, e)
    assert np.all(np.abs(r - r_ref) <= RTOL * np.abs(r_ref))
    a = clipped(np.random.RandomState(N), (E, 10, 2))
    load_state(env, x, xa, na, nx)
    (gx, gxa), gr, gd, _ = env.step(to_dev(a))
    rew, done = so.step(x, xa, a.astype(np.float64), na, nx)
    gx, gxa = gx.cpu().numpy(), gxa.cpu().numpy()
    assert np.array_equal(gxa, xa)
    for e in range(E):
        assert rel_err(gx[e], x[e]) <= STEP_TOL, (N, e)
        g, p_ = so.rasterize(gx[e], gxa[e], 84)
        assert np.array_equal(env.grid[e].cpu().numpy(), g.astype(np.float32))
        assert np.array_equal(env.positions[e].cpu().numpy(), p_)
    assert np.all(np.abs(gr.cpu().numpy() - rew) <= RTOL * np.abs(rew))


# Free-running parity.  Seeds 31000..31127 (128 of them) for N = 64 and N = 80: injected reset (10 burn-in steps), then 16
# free-running steps from the oracle's post-reset state.  FP32 pair forces cannot hold rtol 1e-5 on EVERY seed -- the
# dynamics amplify a perturbation by kappa = 2..800 over 16 steps (close pairs) -- and WHICH of the sensitive
# seeds end up above 1e-5 depends on the last bits of the summation order (31021 + 31125 with round 1's order, 31122 +
# 31125 with round 2's).  So the contract is stated on the ORACLE's sensitivity, which no kernel change moves:
#   * SENSITIVE[N] lists the seeds whose kappa is >= 100 (11 / 14 of the 128; checked against the oracle in the test);
#     every other seed -- the strict list, 117 + 114 seeds -- must pass rtol 1e-5 in fast math;
#   * a sensitive seed may exceed 1e-5 only within err <= 2e-7 * kappa (2e-7 being the single-step error bound observed
#     in test_step_teacher_forced_16);
#   * the pass fraction may not drop below the measured level (126 / 128 at N = 80, 128 / 128 at N = 64) minus one seed.
SENSITIVE = {64: {31023, 31035, 31049, 31062, 31064, 31080, 31081, 31085, 31087, 31106, 31119},
             80: {31011, 31012, 31021, 31023, 31040, 31057, 31065, 31078, 31081, 31088, 31112, 31121, 31122, 31125},
             256: None}                                          # None: report only (N = 256 has no strict list)
KNOWN_CHAOTIC = SENSITIVE
MIN_PASS = {64: 127.0 / 128, 80: 125.0 / 128, 256: 0.0}


def smooth_kappa(draws, E, actions_fn, eps=1e-10):
    """Amplification of a perturbation of the oracle's post-reset state over the same 16 steps.  Only smooth directions
    are perturbed: x everywhere, y of airborne locusts (lifting a grounded locust off y = 0 switches its wind on -- a
    different trajectory, not a sensitivity)."""
    x0, xa0 = so.reset_injected(*draws)
    pert = eps * np.random.RandomState(1).normal(size=x0.shape)
    pert[..., 1] *= x0[..., 1] > 0
    xp, xq, xap, xaq = x0 + pert, x0.copy(), xa0.copy(), xa0.copy()
    rs = np.random.RandomState(5)
    for t in range(16):
        a = actions_fn(rs).astype(np.float64)
        so.step(xp, xap, a, draws[3][:, 10], draws[4][:, 10])
        so.step(xq, xaq, a, draws[3][:, 10], draws[4][:, 10])
    return np.array([np.abs(xp[e] - xq[e]).max() / eps for e in range(E)])


@pytest.mark.parametrize("N", [64, 80, 256])
def test_reset_and_free_running_16_seed_list(M, N):
    E = 128
    seeds = [31000 + e for e in range(E)]
    draws = stack_draws(seeds, N)
    inj = M.InjectedDraws(*draws, device="cuda")
    out = {"seeds": [seeds[0], seeds[-1]], "sensitive_seeds_kappa_ge_100": sorted(SENSITIVE[N] or [])}
    for mode in ("fast", "precise"):
        env = M.BatchedSwarmEnv(E, n_locusts=N, max_episode_steps=0, math_mode=mode, auto_reset=False, rasterize=False)
        gx, gxa = env.reset(draws=inj)
        x, xa = so.reset_injected(*draws)
        e_reset = np.array([rel_err(gx[e].cpu().numpy(), x[e]) for e in range(E)])
        assert np.array_equal(env.noise_x.cpu().numpy(), draws[4][:, 10])
        assert (env.elapsed.cpu().numpy() == 0).all() and (env.episode.cpu().numpy() == 1).all()
        # 16 free-running steps from the ORACLE's post-reset state
        load_state(env, x, xa, draws[3][:, 10], draws[4][:, 10])
        rs = np.random.RandomState(5)
        worst_tf = 0.0
        for t in range(16):
            a = clipped(rs, (E, 10, 2))
            env.step(to_dev(a))
            so.step(x, xa, a.astype(np.float64), draws[3][:, 10], draws[4][:, 10])
        e16 = np.array([rel_err(env.x[e].cpu().numpy(), x[e]) for e in range(E)])
        kappa = smooth_kappa(draws, E, lambda r: clipped(r, (E, 10, 2)))
        bad = [int(seeds[e]) for e in np.nonzero(e16 > RTOL)[0]]
        out[mode] = dict(reset_median=float(np.median(e_reset)), reset_p90=float(np.percentile(e_reset, 90)),
                         reset_max=float(e_reset.max()), reset_pass=float((e_reset <= RTOL).mean()),
                         s16_median=float(np.median(e16)), s16_p90=float(np.percentile(e16, 90)),
                         s16_max=float(e16.max()), s16_pass=float((e16 <= RTOL).mean()),
                         kappa_median=float(np.median(kappa)), kappa_p90=float(np.percentile(kappa, 90)),
                         kappa_max=float(kappa.max()),
                         bad_seeds={s_: dict(err=float(e16[s_ - seeds[0]]), kappa=float(kappa[s_ - seeds[0]])) for s_ in bad})
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, "parity_distribution_n%d.json" % N), "w") as f:
            json.dump(out, f, indent=1)
        if KNOWN_CHAOTIC[N] is None:
            # N = 256 from a U[0,1)^2 start is a dense, strongly repelling swarm: the oracle itself amplifies a 1e-10
            # perturbation by kappa ~ 1e7..1e10 over 16 steps, so no arithmetic narrower than the oracle's can track it
            # free-running.  Reported (profiles/), not asserted; what holds at N = 256 is the teacher-forced per-step
            # bound (test_step_teacher_forced_16, test_golden_trajectory_teacher_forced) -- re-measured here.
            assert np.median(kappa) >= 1e4, out[mode]
            xt, xat = so.reset_injected(*draws)
            rs = np.random.RandomState(5)
            worst = 0.0
            for t in range(16):
                a = clipped(rs, (E, 10, 2))
                load_state(env, xt, xat, draws[3][:, 10], draws[4][:, 10])
                env.step(to_dev(a))
                so.step(xt, xat, a.astype(np.float64), draws[3][:, 10], draws[4][:, 10])
                worst = max(worst, max(rel_err(env.x[e].cpu().numpy(), xt[e]) for e in range(E)))
            out[mode]["teacher_forced_16_worst_step"] = worst
            with open(os.path.join(OUT, "parity_distribution_n%d.json" % N), "w") as f:
                json.dump(out, f, indent=1)
            assert worst <= STEP_TOL, worst
            continue
        assert np.median(e16) <= 2e-6 and np.median(e_reset) <= 2e-6, out[mode]
        assert (e16 <= RTOL).mean() >= MIN_PASS[N] - (1.0 / 128 if mode == "precise" else 0.0), out[mode]
        assert (e_reset <= RTOL).mean() >= 0.98, out[mode]
        assert {int(seeds[e]) for e in np.nonzero(kappa >= 100.0)[0]} == SENSITIVE[N]      # the list is the oracle's, not ours
        for s_ in bad:          # every failing seed is explained by the oracle's own sensitivity
            e = s_ - seeds[0]
            assert e16[e] <= 2e-7 * kappa[e], (s_, e16[e], kappa[e])
        if mode == "fast":
            assert set(bad) <= SENSITIVE[N], "seeds off the strict list fail rtol 1e-5: %s" % sorted(set(bad) - SENSITIVE[N])


@pytest.mark.parametrize("N", [64, 256])
def test_golden_trajectory_teacher_forced(M, N):
    """Against the REFERENCE's own outputs (tests/golden/traj16_n*.npz): each step from golden state t
    must land on golden state t+1; reward = golden reward."""
    d = golden("traj16_n%d.npz" % N)
    env = M.BatchedSwarmEnv(16, n_locusts=N, max_episode_steps=0, auto_reset=False)
    na = np.broadcast_to(d["agent_noise"][10], (16, 10, 2))
    nx = np.broadcast_to(d["particle_noise"][10], (16, N, 2))
    load_state(env, d["x"][:16], d["xa"][:16], na, nx)
    (gx, gxa), r, done, _ = env.step(to_dev(d["actions"].astype(np.float32)))
    gx, gxa, r = gx.cpu().numpy(), gxa.cpu().numpy(), r.cpu().numpy()
    for t in range(16):
        assert rel_err(gx[t], d["x"][t + 1]) <= STEP_TOL, t
        assert np.array_equal(gxa[t], d["xa"][t + 1])
    assert np.all(np.abs(r - d["reward"]) <= RTOL * np.abs(d["reward"]))
    assert not done.any().item()
    # fused rasterise of (nearly) the golden states: agent cells exact (the agent update is FP64-exact); the grid is
    # bit-exact for the DEVICE positions, and differs from the golden grid only where a locust sits within the FP32
    # step error of a bin edge (a handful of cells at most, each by one count)
    grid, pos = env.grid.cpu().numpy(), env.positions.cpu().numpy()
    assert np.array_equal(pos, d["pos"][1:17])
    moved = 0
    for t in range(16):
        g_dev, p_dev = so.rasterize(gx[t], gxa[t], 84)
        assert np.array_equal(grid[t], g_dev.astype(np.float32)) and np.array_equal(pos[t], p_dev), t
        ref = dense_grid(d["grid_idx_%d" % (t + 1)], d["grid_val_%d" % (t + 1)]).astype(np.float32)
        assert np.array_equal(grid[t][..., 1], ref[..., 1]), t                     # agent channel: exact
        diff = np.abs(grid[t][..., 0] - ref[..., 0])
        assert diff.max() <= 1.0 / N + 1e-7                                         # one locust at most per cell
        moved += int((diff > 0).sum())
        assert abs(float(grid[t][..., 0].sum()) - float(ref[..., 0].sum())) <= 1.0 / N + 1e-6      # nobody lost
    assert moved <= 8, moved


# --------------------------------------------------------------------------------------- rasteriser
def test_rasterizer_bit_exact_on_reference_states(M):
    """Grids / positions bit-exact vs the reference's process_state on identical FP64 positions."""
    for name, N in (("traj16_n64.npz", 64), ("traj16_n256.npz", 256), ("c1_seed192_n80.npz", 80)):
        d = golden(name)
        E = d["x"].shape[0]
        env = M.BatchedSwarmEnv(E, n_locusts=N)
        env.x.copy_(to_dev(d["x"])); env.xa.copy_(to_dev(d["xa"]))
        box = torch.zeros(E, 4, dtype=torch.float64, device="cuda")
        grid, pos = env.observe(box=box)
        grid, pos = grid.cpu().numpy(), pos.cpu().numpy()
        for k in range(E):
            ref = dense_grid(d["grid_idx_%d" % k], d["grid_val_%d" % k]).astype(np.float32)
            assert np.array_equal(grid[k], ref), (name, k)
            assert np.array_equal(pos[k], d["pos"][k]), (name, k)
            m = so.sequential_mean_x(d["x"][k], d["xa"][k])
            assert box[k].cpu().numpy().tolist() == [m - 1.5, m + 1.5, 0.0, 6.0]


def test_rasterizer_edge_cases_through_state_processor(pkg):
    """SwarmStateProcessor facade on the adversarial golden cases (points on edges, one-cell pile-up,
    everyone grounded, G=84 and G=20, N in {5,33,64,80,256})."""
    SP = pkg.submodule("agents.state_processors").SwarmStateProcessor
    d = golden("raster_cases.npz")
    for k in range(int(d["n_cases"])):
        G = int(d["G_%d" % k])
        proc = SP(grid_size=G)
        g = proc.process_state([d["x_%d" % k], d["xa_%d" % k]])
        ref = dense_grid(d["grid_idx_%d" % k], d["grid_val_%d" % k], G)
        assert g.shape == (G, G, 2) and g.dtype == np.float64
        assert np.array_equal(g.astype(np.float32), ref.astype(np.float32)), k
        assert proc.positions.dtype == np.uint8 and np.array_equal(proc.positions, d["pos_%d" % k]), k


def test_rasterizer_verified_parallel_mean_next_to_edges(M):
    """The rasteriser bins against a PARALLEL mean and walks numpy's sequential one only when some point is too close to
    an edge for the parallel mean to be trusted (env_raster).  States with a locust / an agent ON one of numpy's own
    edges (first, inner, last = the histogram's closed right edge), a few ulps beside it, and windows far from the
    origin: without the box (verified parallel mean + fall-back) and with it (always sequential) vs the oracle."""
    rs = np.random.RandomState(7)
    G = 84
    for N in (5, 64, 80, 256):
        xs, xas = [], []
        for trial in range(8):
            x = rs.rand(N, 2) * [2.5, 1.5] + [rs.uniform(-40.0, 40.0) * (trial % 3), 0.0]
            xa = rs.rand(10, 2) * [3.4, 3.0] + [x[:, 0].mean() - 1.7, 0.0]
            j, i = rs.randint(N), (0, G, rs.randint(1, G))[trial % 3]
            k = rs.randint(10)
            for _ in range(40):          # fixed point: the window moves with the point that is put on its edge
                m = so.sequential_mean_x(x, xa)
                edges = np.linspace(m - 1.5, m + 1.5, G + 1)
                x[j, 0] = edges[i]
                if trial >= 4:
                    xa[k, 0] = edges[(i * 7 + 3) % (G + 1)]
            for ulps in (0, 1, -1, 2, -3):
                y, ya = x.copy(), xa.copy()
                for _ in range(abs(ulps)):
                    y[j, 0] = np.nextafter(y[j, 0], np.inf if ulps > 0 else -np.inf)
                    ya[k, 0] = np.nextafter(ya[k, 0], -np.inf if ulps > 0 else np.inf)
                xs.append(y); xas.append(ya)
        xs, xas = np.stack(xs), np.stack(xas)
        E = len(xs)
        env = M.BatchedSwarmEnv(E, n_locusts=N)
        env.x.copy_(to_dev(xs)); env.xa.copy_(to_dev(xas))
        g1, p1 = env.observe()
        g1, p1 = g1.cpu().numpy().copy(), p1.cpu().numpy().copy()
        box = torch.zeros(E, 4, dtype=torch.float64, device="cuda")
        g2, p2 = env.observe(box=box)
        g2, p2 = g2.cpu().numpy(), p2.cpu().numpy()
        for e in range(E):
            go, po = so.rasterize(xs[e], xas[e], G)
            assert np.array_equal(g1[e], go.astype(np.float32)) and np.array_equal(p1[e], po), (N, e)
            assert np.array_equal(g2[e], go.astype(np.float32)) and np.array_equal(p2[e], po), (N, e)


def test_fused_raster_equals_standalone_and_oracle(M):
    E, N = 64, 80
    env = M.BatchedSwarmEnv(E, n_locusts=N, seed=11)
    env.reset()
    rs = np.random.RandomState(1)
    for t in range(5):
        env.step(to_dev(clipped(rs, (E, 10, 2)) * 3.0))       # big actions: agents leave the box
        g1, p1 = env.grid.clone(), env.positions.clone()
        g2, p2 = env.observe()
        assert torch.equal(g1, g2) and torch.equal(p1, p2)
    x, xa = env.x.cpu().numpy(), env.xa.cpu().numpy()
    g, p = g2.cpu().numpy(), p2.cpu().numpy()
    for e in range(E):
        go, po = so.rasterize(x[e], xa[e], 84)
        assert np.array_equal(g[e], go.astype(np.float32)) and np.array_equal(p[e], po)


# --------------------------------------------------------------------------------------- done / reset
def test_timelimit_autoreset_matches_oracle_runner(M):
    E, N, LIM = 6, 16, 5
    table = {}

    def draws(e, ep):
        if (e, ep) not in table:
            table[(e, ep)] = so.draw_reset_numpy(np.random.RandomState(1000 * e + ep), N)
        return table[(e, ep)]

    def batch(ep):
        ds = [draws(e, ep) for e in range(E)]
        return M.InjectedDraws(*[np.stack([d[i] for d in ds]) for i in range(5)], device="cuda")

    orc = so.OracleRunner(E, N, draws, grid_size=84, max_episode_steps=LIM)
    orc.reset()
    env = M.BatchedSwarmEnv(E, n_locusts=N, max_episode_steps=LIM, auto_reset=True)
    env.reset(draws=batch(0))
    rs = np.random.RandomState(8)
    for t in range(1, 13):
        a = clipped(rs, (E, 10, 2))
        (gx, gxa), r, d, _ = env.step(to_dev(a), reset_draws=batch((t - 1) // LIM + 1))
        rew, done = orc.step(a.astype(np.float64))
        assert np.array_equal(d.cpu().numpy(), done), t
        assert np.array_equal(env.elapsed.cpu().numpy(), orc.elapsed), t
        assert np.array_equal(env.episode.cpu().numpy(), orc.episode), t
        assert np.all(np.abs(r.cpu().numpy() - rew) <= RTOL * np.abs(rew))
        for e in range(E):
            assert rel_err(gx[e].cpu().numpy(), orc.x[e]) <= 1e-4      # free-running incl. burn-in, N=16
        # teacher-force the oracle onto the device state so bins cannot flip, then compare observations
        orc.x[...] = gx.cpu().numpy(); orc.xa[...] = gxa.cpu().numpy()
        grids, pos = orc.observe()
        assert np.array_equal(env.grid.cpu().numpy(), grids.astype(np.float32)), t
        assert np.array_equal(env.positions.cpu().numpy(), pos), t
    assert done.sum() == 0 and orc.episode.max() == 3


def test_step_before_reset_raises(M):
    env = M.BatchedSwarmEnv(2, n_locusts=8)
    with pytest.raises(TypeError):
        env.step(torch.zeros(2, 10, 2, device="cuda"))
    with pytest.raises(ValueError):
        env.reset(); env.step(torch.zeros(3, 10, 2, device="cuda"))


def test_reward_nonnegative_sets_done(M):
    """done = reward >= 0 (multiagent.py:44) needs zero energy: one grounded locust far from a
    single far agent, no wind/gravity."""
    env = M.BatchedSwarmEnv(1, n_locusts=1, n_agents=1, max_episode_steps=0, auto_reset=False, rasterize=False)
    env.params.wind = 0.0; env.params.gravity = 0.0; env.params.noise = 0.0
    load_state(env, np.zeros((1, 1, 2)), np.full((1, 1, 2), 1e4), np.zeros((1, 1, 2)), np.zeros((1, 1, 2)))
    _, r, d, _ = env.step(torch.zeros(1, 1, 2, device="cuda"))
    assert r.item() == 0.0 and bool(d.item()) is True


# --------------------------------------------------------------------------------------- Philox
def test_philox_device_integers_bit_exact(pkg):
    nat = pkg._native
    lib = nat.load()
    rs = np.random.RandomState(0)
    ctr = rs.randint(0, 2 ** 32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    key = rs.randint(0, 2 ** 32, size=(4096, 2), dtype=np.uint64).astype(np.uint32)
    for i, (c, k, _) in enumerate(ph.KAT):
        ctr[i], key[i] = c, k
    dc, dk = to_dev(ctr.view(np.int32)), to_dev(key.view(np.int32))
    out = torch.zeros(4096, 4, dtype=torch.int32, device="cuda")
    nat.check(lib.swarm_philox_raw(ctypes.c_void_p(dc.data_ptr()), ctypes.c_void_p(dk.data_ptr()),
                                   ctypes.c_void_p(out.data_ptr()), 4096, None))
    got = out.cpu().numpy().view(np.uint32)
    ref = np.stack(ph.philox4x32_10(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], key[:, 0], key[:, 1]), axis=1)
    assert np.array_equal(got, ref)
    for i, (_, _, o) in enumerate(ph.KAT):
        assert got[i].tolist() == list(o)


def test_philox_reset_equals_injected_reset_and_oracle(M):
    E, N, seed, off = 16, 80, 20261018, 5
    a = M.BatchedSwarmEnv(E, n_locusts=N, seed=seed, env_id_offset=off)
    b = M.BatchedSwarmEnv(E, n_locusts=N, seed=seed, env_id_offset=off)
    for episode in range(2):
        inj = a.philox_draws()                     # the draws a's NEXT reset will use
        a.reset()
        b.reset(draws=inj)
        for k in ("x", "xa", "noise_x", "noise_a"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
        host = [t.cpu().numpy() for t in (inj.x0, inj.xa0, inj.burn_actions, inj.agent_noise, inj.particle_noise)]
        x, xa = so.reset_injected(*host)
        # 10 free-running burn-in steps from a dense uniform start are chaotic: distribution, not max
        errs = np.array([rel_err(a.x[e].cpu().numpy(), x[e]) for e in range(E)])
        assert np.median(errs) <= 2e-6 and (errs <= RTOL).mean() >= 0.8, errs
        for e in range(E):      # exported draws vs the NumPy restatement of the draw layout
            ref = ph.reset_draws(seed, off + e, episode, N)
            assert np.array_equal(host[0][e], ref[0]) and np.array_equal(host[1][e], ref[1])   # uniforms: bit-exact
            for got, want in zip((host[2][e], host[3][e], host[4][e]), ref[2:]):
                assert np.abs(got - want).max() <= 2e-6                                          # FP32 Box-Muller
    pn = host[4]
    assert abs(pn.mean()) < 0.02 and abs(pn.std() - 1.0) < 0.02


def test_shard_invariance_bitwise(M):
    """Global env ids key the RNG: 8 envs on one 'rank' == 4+4 envs on two 'ranks', bit for bit."""
    N, seed = 64, 77
    whole = M.BatchedSwarmEnv(8, n_locusts=N, seed=seed)
    parts = [M.BatchedSwarmEnv(4, n_locusts=N, seed=seed, env_id_offset=4 * r) for r in range(2)]
    whole.reset()
    [p.reset() for p in parts]
    rs = np.random.RandomState(3)
    for t in range(130):                 # crosses the 128-step auto-reset
        a = to_dev(clipped(rs, (8, 10, 2)))
        whole.step(a)
        for r, p in enumerate(parts):
            p.step(a[4 * r:4 * r + 4].contiguous())
    for k in ("x", "xa", "grid", "positions", "reward", "elapsed", "episode"):
        cat = torch.cat([getattr(p, k) for p in parts])
        assert torch.equal(getattr(whole, k), cat), k
    assert whole.episode.cpu().tolist() == [2] * 8 and whole.elapsed.cpu().tolist() == [2] * 8


# --------------------------------------------------------------------------------------- PAAC boundary
def test_runner_statics_match_golden(pkg):
    R = pkg.submodule("agents.paac.emulator_runner").SwarmRunner
    d = golden("runner_statics.npz")
    a = to_dev(d["clip_in"])
    ret = R.transform_actions_for_env(a)
    assert ret is a                                                  # in place, same object
    ref = d["clip_out"]
    assert np.abs(a.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    assert np.array_equal(a.cpu().numpy()[2], d["clip_in"][2])      # |a| < 1 rows untouched
    g = to_dev(dense_grid(d["ls_grid_idx"], d["ls_grid_val"]).astype(np.float32))
    pos = to_dev(d["ls_pos"])
    loc = R.get_local_states(g, pos).cpu().numpy()
    assert loc.shape == tuple(d["ls_shape"])
    ref = so.local_states(g.cpu().numpy(), d["ls_pos"])
    assert np.array_equal(loc, ref)
    assert np.array_equal(np.argwhere(loc[:, :, :, 2] != 0), d["ls_hot"])


def test_grid_runners_contract(pkg, M):
    """Six shared variables (runners.py:42-43, paac.py:263-274): shapes, dtypes, in-place updates."""
    GR = pkg.submodule("agents.paac.runners").GridRunners
    R = pkg.submodule("agents.paac.emulator_runner").SwarmRunner
    E = 8
    env = M.BatchedSwarmEnv(E, n_locusts=80, seed=3, max_episode_steps=4)
    runners = GR(env, 8, None, R, None, 84)
    runners.start()
    states, hist, pos, rew, over, act = runners.get_shared_variables()
    assert states.shape == (E, 10, 84, 84, 3) and states.dtype == torch.float32
    assert pos.shape == (E, 10, 2) and pos.dtype == torch.uint8
    assert rew.shape == (E, 10) and over.shape == (E, 10) and act.shape == (E, 10, 2)
    shared = (states, pos, rew, over, act)
    ptrs = [t.data_ptr() for t in shared]
    first = states.clone()
    assert torch.equal(states[..., :2], env.grid[:, None].expand(E, 10, 84, 84, 2))
    assert states[..., 2].sum().item() == E * 10
    for t in range(1, 6):
        act.copy_(torch.randn(E, 10, 2, device="cuda"))
        R.transform_actions_for_env(act)
        runners.update_environments()
        runners.wait_updated()
        assert (rew[:, 0] < 0).all() and torch.equal(rew, rew[:, :1].expand(E, 10))
        assert over.sum().item() == (E * 10 if t == 4 else 0)
    again = runners.get_shared_variables()
    assert ptrs == [again[i].data_ptr() for i in (0, 2, 3, 4, 5)]          # updated in place, never rebound
    assert not torch.equal(first, states)
    assert env.episode.cpu().tolist() == [2] * E and env.elapsed.cpu().tolist() == [1] * E


# --------------------------------------------------------------------------------------- gym facade
def test_reference_swarm_tests_on_facade(pkg, M):
    """tests/env_tests.py:10-44 (SwarmTests.run_env_test, bounding_box_test) against the facade."""
    SP = pkg.submodule("agents.state_processors").SwarmStateProcessor
    env = M.SwarmEnv()
    env.reset()
    for _ in range(20):
        state, reward, done, _ = env.step(np.random.normal(size=(env.N_AGENTS, 2)))
    assert len(state) == 2 and state[0].shape == (env.N_LOCUSTS, 2) and state[1].shape == (env.N_AGENTS, 2)
    assert reward < 0. and not done
    proc = SP()
    env = M.SwarmEnv()
    env.reset()
    for _ in range(200):
        state, reward, done, _ = env.step(np.zeros((10, 2)))
        proc.process_state(state)
    max_x, max_y = state[0].max(axis=0)
    min_x, min_y = state[0].min(axis=0)
    bb = proc._get_bounding_box(state[0])
    assert max_x < bb[0][1] and min_x > bb[0][0] and max_y < bb[1][1] and min_y >= bb[1][0]


def test_seeded_facade_matches_reference_golden(M):
    """SwarmEnv(seed=192) (= Swarm-eval-v0) consumes numpy's global RNG in the reference's order:
    post-reset state and the first step agree with the reference's golden episode."""
    d = golden("c1_seed192_n80.npz")
    env = M.make("Swarm-eval-v0")
    st = env.reset()
    assert env.t == 10
    assert rel_err(st[0], d["x"][0]) <= RTOL and rel_err(st[1], d["xa"][0]) <= 1e-12
    st2, r, done, _ = env.step(d["actions"][0])
    assert st2 is st and rel_err(st[0], d["x"][1]) <= RTOL
    assert abs(r - d["reward"][0]) <= RTOL * abs(d["reward"][0]) and not done
    for t in range(1, 128):
        _, r, done, _ = env.step(d["actions"][t])
    assert done                                   # TimeLimit(128) of the registered id
    # static helpers, each through the C ABI
    x = d["x"][3].copy(); xa = d["xa"][3].copy()
    v, rew = M.SwarmEnv.v_calculate(x, xa, 0.5, 10, 1, -1)
    v_ref, r_ref = so.forces(x[None], xa[None])
    assert rel_err(v, v_ref[0]) <= RTOL and abs(rew - r_ref[0]) <= RTOL * abs(r_ref[0])
    rr = np.linspace(0, 5, 50)
    assert np.abs(M.SwarmEnv.s(rr, 0.5, 10) - so.s_potential(rr)).max() < 1e-15
    p = np.array([[0.0, -0.1], [1.0, 0.5], [2.0, 0.0]]); w = np.array([[1.0, -1.0], [1.0, -2.0], [3.0, 2.0]])
    p2, w2 = p.copy(), w.copy()
    noise = np.full((3, 2), 1e-4)
    out = M.SwarmEnv.x_update(p, w, 0.05, noise)
    so.move_(p2, w2, noise, 0.05)
    assert out is p and np.array_equal(p, p2) and np.array_equal(w, w2)


# --------------------------------------------------------------------------------------- host buffers
@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_matches_device_step(M, pinned):
    """swarm_step_host: pinned buffers take the zero-copy path (the kernel reads the actions from and writes
    reward/done to host memory), pageable ones the staged-copy path; both equal the device-resident step."""
    E, N = 300, 64
    a = M.BatchedSwarmEnv(E, n_locusts=N, seed=5, max_episode_steps=2)
    b = M.BatchedSwarmEnv(E, n_locusts=N, seed=5, max_episode_steps=2)
    a.reset(); b.reset()
    pin = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    h_act = pin(torch.randn(E, 10, 2).clamp_(-0.5, 0.5))
    h_rew = pin(torch.zeros(E))
    h_done = pin(torch.zeros(E, dtype=torch.uint8))
    for t in range(3):
        a.step_host(h_act, h_rew, h_done)
        _, r, d, _ = b.step(h_act.cuda())
        torch.cuda.synchronize()
        assert torch.equal(h_rew, r.cpu()) and torch.equal(h_done.bool(), d.cpu())
        assert torch.equal(a.x, b.x) and torch.equal(a.grid, b.grid)


# --------------------------------------------------------------------------------------- full size
def test_full_size_properties_c4(M):
    """BASELINE config 4 size (4096 x 256): size-independent properties instead of an oracle run."""
    E, N = 4096, 256
    env = M.BatchedSwarmEnv(E, n_locusts=N, seed=1234)
    env.reset()
    assert (env.x[..., 1] >= 0).all() and torch.isfinite(env.x).all()
    a = torch.randn(E, 10, 2, device="cuda")
    v = torch.zeros(E, N, 2, device="cuda")
    sd = env.state_dict()
    (x, xa), r, d, _ = env.step(a, clip=True, v_out=v)
    assert (a.norm(dim=-1) <= 1.0 + 1e-6).all()                       # clipped in place
    x1, g1, p1, r1 = x.clone(), env.grid.clone(), env.positions.clone(), r.clone()
    # (i) determinism: same state + same actions -> bitwise identical
    env.load_state_dict(sd)
    env.step(a, v_out=v)
    assert torch.equal(env.x, x1) and torch.equal(env.grid, g1) and torch.equal(env.reward, r1)
    # (ii) reward is -mean |v|^2 of the velocities the kernel used
    r_chk = -(v.double() ** 2).sum(-1).mean(-1)
    assert torch.allclose(r1.double(), r_chk, rtol=1e-6)
    assert (r1 < 0).all() and not d.any() and (env.x[..., 1] >= 0).all()
    # (iii) histogram conservation: channel sums == fraction of points inside the box (FP64 recount)
    m = (torch.cat([env.x[..., 0], env.xa[..., 0]], 1).cumsum(1)[:, -1] / (N + 10))[:, None]
    def inside(p):
        return ((p[..., 0] >= m - 1.5) & (p[..., 0] <= m + 1.5) & (p[..., 1] >= 0) & (p[..., 1] <= 6)).sum(1)
    cl = (env.grid[..., 0].double().sum((1, 2)) * N).round().long()
    ca = (env.grid[..., 1].double().sum((1, 2)) * 10).round().long()
    assert (cl - inside(env.x)).abs().max() <= 1 and (ca - inside(env.xa)).abs().max() <= 1   # edge ties only
    assert (cl == inside(env.x)).float().mean() > 0.99
    assert (env.positions < 84).all()
    # (iv) permutation invariance: relabelling locusts leaves the occupancy grid bit-identical and
    #      the forces equal up to FP32 summation order
    perm = torch.randperm(N, device="cuda")
    v0, r0 = env.forces()
    v0, r0 = v0.clone(), r0.clone()
    env.x.copy_(x1[:, perm])
    g2, p2 = env.observe()
    assert torch.equal(g2, g1) and torch.equal(p2, p1)
    v2, r2 = env.forces()
    assert torch.allclose(r2, r0, rtol=1e-4)
    scale = v0.abs().amax(dim=(1, 2), keepdim=True)
    assert ((v2 - v0[:, perm]).abs() <= 1e-4 * scale).all()


def test_torch_extension_binding_equals_ctypes_binding(M):
    """torch.ops.swarm_b200.* and the ctypes binding drive the same kernels: bitwise-identical rollouts."""
    E, N = 16, 80
    a = M.BatchedSwarmEnv(E, n_locusts=N, seed=11, binding="torch")
    b = M.BatchedSwarmEnv(E, n_locusts=N, seed=11, binding="ctypes")
    assert a.ops is not None and b.ops is None
    a.reset(); b.reset()
    assert torch.equal(a.x, b.x) and torch.equal(a.noise_x, b.noise_x)
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    for t in range(130):                                   # crosses the 128-step auto-reset
        act = torch.randn(E, 10, 2, device="cuda", generator=gen)
        act2 = act.clone()
        _, ra, da, _ = a.step(act, clip=True)
        _, rb, db, _ = b.step(act2, clip=True)
        assert torch.equal(act, act2)
    torch.cuda.synchronize()
    for k in ("x", "xa", "noise_x", "noise_a", "elapsed", "episode", "grid", "positions", "reward", "done_u8"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert torch.equal(a.local_states(), b.local_states())
    va, _ = a.forces(); vb, _ = b.forces()
    assert torch.equal(va, vb)
    with pytest.raises(ValueError):
        a.step(torch.zeros(E, 10, 2, device="cuda", dtype=torch.float16))


@pytest.mark.parametrize("G", [2, 20, 83, 84, 85, 255])
@pytest.mark.parametrize("A", [1, 10, 32])
def test_rasterizer_grid_sizes_and_agent_counts(M, G, A):
    """Fused and standalone rasteriser for other grid sizes (reference default grid_size=20; odd sizes take the
    plain-store zero fill instead of the TMA one; 255 is the uint8 limit) and agent counts, bit-exact vs the
    oracle, over two consecutive steps (the counter table must come back clean)."""
    E, N = 5, 48
    rs = np.random.RandomState(G * 100 + A)
    if G == 255 and A == 32:        # 32 agents need 32-bit counters: 255^2 of them do not fit in shared memory
        with pytest.raises(Exception):
            M.BatchedSwarmEnv(E, n_locusts=N, n_agents=A, grid_size=G)
        return
    env = M.BatchedSwarmEnv(E, n_locusts=N, n_agents=A, grid_size=G, max_episode_steps=0, auto_reset=False)
    x = rs.rand(E, N, 2) * [2.5, 1.2]
    x[:, ::3, 1] = 0.0                                   # grounded locusts share y-bin 0
    xa = rs.rand(E, A, 2) * [4.0, 7.0] - [0.5, 0.5]       # some agents outside the box
    zeros_a, zeros_x = np.zeros((E, A, 2)), np.zeros((E, N, 2))
    load_state(env, x, xa, zeros_a, zeros_x)
    for t in range(2):
        a = clipped(rs, (E, A, 2))
        (gx, gxa), _, _, _ = env.step(to_dev(a))
        gx, gxa = gx.cpu().numpy(), gxa.cpu().numpy()
        fused_g, fused_p = env.grid.cpu().numpy().copy(), env.positions.cpu().numpy().copy()
        g2, p2 = env.observe()
        assert np.array_equal(fused_g, g2.cpu().numpy()) and np.array_equal(fused_p, p2.cpu().numpy())
        for e in range(E):
            g, p_ = so.rasterize(gx[e], gxa[e], G)
            assert np.array_equal(fused_g[e], g.astype(np.float32)), (G, A, t, e)
            assert np.array_equal(fused_p[e], p_), (G, A, t, e)


def test_step_under_cuda_graph_replay_equals_eager(M):
    """The rollout contract 'no host round trip': 8 captured steps replayed 20 times == 160 eager steps, bitwise
    (crosses the 128-step auto-reset inside the graph)."""
    E, N = 64, 64
    acts = [torch.randn(E, 10, 2, device="cuda").clamp(-0.7, 0.7).contiguous() for _ in range(8)]
    a = M.BatchedSwarmEnv(E, n_locusts=N, seed=21)
    b = M.BatchedSwarmEnv(E, n_locusts=N, seed=21)
    a.reset(); b.reset()
    a.step(acts[0]); b.step(acts[0])                     # warm-up outside the capture (one-time attribute calls)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(8):
            a.step(acts[i])
    for _ in range(20):
        g.replay()
    for _ in range(20):
        for i in range(8):
            b.step(acts[i])
    torch.cuda.synchronize()
    for k in ("x", "xa", "elapsed", "episode", "grid", "positions", "reward", "done_u8"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert int(a.episode[0]) == 2


@pytest.mark.parametrize("G", [20, 83, 84])
def test_follower_rasteriser_large_swarm_grid_sizes(M, G):
    """The two-kernel step (k_step + k_raster_follow on the side stream, forced through SwarmParams.tuning): observations
    bit-exact vs the oracle for TMA-able and odd grid sizes, over consecutive steps including an auto-reset, and
    identical to the raster warps inside k_step."""
    nat = M.nat
    E, N = 37, 176
    a = M.BatchedSwarmEnv(E, n_locusts=N, grid_size=G, seed=9, max_episode_steps=3, binding="ctypes",
                          tuning=nat.TUNE_RASTER_FOLLOW)
    b = M.BatchedSwarmEnv(E, n_locusts=N, grid_size=G, seed=9, max_episode_steps=3, binding="ctypes",
                          tuning=nat.TUNE_RASTER_WARPS)
    b.state_c.work, b.state_c.work_words = None, 0           # static assignment, no work queue
    a.reset(); b.reset()
    rs = np.random.RandomState(G)
    for t in range(5):
        act = to_dev(clipped(rs, (E, 10, 2)))
        a.step(act); b.step(act.clone())
        torch.cuda.synchronize()
        assert torch.equal(a.x, b.x) and torch.equal(a.grid, b.grid) and torch.equal(a.positions, b.positions)
        assert torch.equal(a.done_u8, b.done_u8) and int(a.work.sum()) == 0
        gx, gxa = a.x.cpu().numpy(), a.xa.cpu().numpy()
        for e in range(0, E, 6):
            g, p_ = so.rasterize(gx[e], gxa[e], G)
            assert np.array_equal(a.grid[e].cpu().numpy(), g.astype(np.float32)), (G, t, e)
            assert np.array_equal(a.positions[e].cpu().numpy(), p_)
    assert int(a.episode[0]) == 2


@pytest.mark.parametrize("N", [40, 64, 80, 128, 176, 256, 320, 512, 700])
def test_launch_shapes_bitwise_identical(M, N):
    """swarm_step picks its launch shape by batch size: the rasteriser runs in a follower kernel, in raster warps of the
    step kernel or on the step's own threads (with the agents' pull evaluated before or after the tile passes, a filler
    warp or plain stores for the zeros, programmatic dependent launch or not).  Every shape must give the same BITS -- a
    trajectory must not depend on how many envs share a GPU -- and agree with the oracle."""
    nat = M.nat
    E = 21
    ref = None
    rs = np.random.RandomState(N)
    acts = [to_dev(clipped(rs, (E, 10, 2))) for _ in range(5)]
    for place in (nat.TUNE_RASTER_WARPS, nat.TUNE_RASTER_SELF, nat.TUNE_RASTER_FOLLOW, 0):
        env = M.BatchedSwarmEnv(E, n_locusts=N, seed=77, max_episode_steps=3, tuning=place)
        env.reset()
        v0, r0 = env.forces()
        outs = [v0.clone(), r0.clone()]
        for t in range(5):                          # crosses an auto-reset (limit 3) inside the step kernel
            env.step(acts[t].clone())
            outs += [env.x.clone(), env.xa.clone(), env.grid.clone(), env.positions.clone(), env.reward.clone(),
                     env.done_u8.clone(), env.episode.clone(), env.elapsed.clone()]
        torch.cuda.synchronize()
        assert int(env.work.sum()) == 0
        if ref is None:
            ref = outs
            x, xa = env.x.cpu().numpy(), env.xa.cpu().numpy()
            for e in range(0, E, 5):
                g, p_ = so.rasterize(x[e], xa[e], 84)
                assert np.array_equal(env.grid[e].cpu().numpy(), g.astype(np.float32))
                assert np.array_equal(env.positions[e].cpu().numpy(), p_)
        else:
            for i, (a, b) in enumerate(zip(ref, outs)):
                assert torch.equal(a, b), (N, place >> 4, i)


def test_add_wind_false_matches_reference_semantics(M):
    """SwarmEnv._step(v_action, add_wind=False) (multiagent.py:30,35-36): the action is used as it is, the locusts still
    feel the wind U.  Checked against the oracle by folding the wind into the action."""
    N = 80
    draws = stack_draws([901], N)
    x, xa = so.reset_injected(*draws)
    na, nx = draws[3][:, 10], draws[4][:, 10]
    env = M.BatchedSwarmEnv(1, n_locusts=N, max_episode_steps=0, auto_reset=False, rasterize=False)
    a = clipped(np.random.RandomState(2), (1, 10, 2)).astype(np.float64)
    load_state(env, x, xa, na, nx)
    (gx, gxa), r, _, _ = env.step(to_dev(a), add_wind=False)
    a_ref = a.copy()
    a_ref[..., 0] -= 1.0                                  # the oracle always adds WIND_SPEED = 1 (exact in FP64 here?)
    rew, _ = so.step(x, xa, a_ref, na, nx)
    # (a - 1) + 1 may differ from a in the last bit: compare the agents to 1e-15 instead of exactly
    assert np.abs(gxa.cpu().numpy() - xa).max() <= 1e-15
    assert rel_err(gx.cpu().numpy()[0], x[0]) <= STEP_TOL
    assert abs(float(r[0]) - rew[0]) <= RTOL * abs(rew[0])
    # facade: SwarmEnv()._step(a, add_wind=False) no longer raises
    f = M.SwarmEnv(seed=5)
    f.reset()
    st, rr, dd, _ = f._step(np.zeros((10, 2)), add_wind=False)
    assert rr < 0 and not dd


def test_stagger_episodes_spreads_the_resets(M):
    """BatchedSwarmEnv.stagger_episodes (opt-in; the reference keeps every emulator in lock-step): with E = 4 x limit envs
    exactly 4 episodes end on every step, every env resets once per `limit` steps, and an env's trajectory within an
    episode is the one the un-staggered batch produces from the same state (the stagger only moves the TimeLimit)."""
    E, N, LIM = 64, 16, 16
    env = M.BatchedSwarmEnv(E, n_locusts=N, seed=8, max_episode_steps=LIM)
    env.reset()
    el = env.stagger_episodes().cpu().numpy()
    assert sorted(el.tolist()) == sorted([(e * LIM) // E for e in range(E)])
    act = torch.zeros(E, 10, 2, device="cuda")
    resets = np.zeros(E, dtype=int)
    for t in range(2 * LIM):
        _, _, done, _ = env.step(act)
        d = done.cpu().numpy()
        assert d.sum() == E // LIM, (t, d.sum())
        resets += d
    assert (resets == 2).all()
    assert (env.episode.cpu().numpy() == 3).all()
