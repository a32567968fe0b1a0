"""Stress of the two-kernel step (k_step + k_raster_follow): random sizes, many steps, graph and eager, two envs on two
streams at once; checks against the in-kernel-raster shape bit for bit.  Run under `timeout`."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import golds_rl_gym_b200 as pkg
M = pkg.submodule("envs.multiagent")
random.seed(0)
for trial in range(12):
    E = random.choice([1, 3, 37, 300, 1500, 5000])
    N = random.choice([160, 176, 200, 256, 300, 512, 700])
    G = random.choice([20, 83, 84])
    a = M.BatchedSwarmEnv(E, n_locusts=N, grid_size=G, seed=trial, max_episode_steps=random.choice([2, 5, 128]), binding="ctypes")
    b = M.BatchedSwarmEnv(E, n_locusts=N, grid_size=G, seed=trial, max_episode_steps=a.params.max_episode_steps, binding="ctypes")
    b.state_c.work = None
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
    ok = all(torch.equal(getattr(a, k), getattr(b, k)) for k in ("x", "xa", "grid", "positions", "reward", "done_u8", "episode"))
    assert ok and int(a.work.sum()) == 0, (trial, E, N, G)
    print("trial", trial, E, N, G, "ok", flush=True)
print("stress ok")
