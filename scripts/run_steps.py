"""Eager steps of one (E, N) shape -- a target for ncu / compute-sanitizer.   python scripts/run_steps.py E N [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import golds_rl_gym_b200 as pkg
M = pkg.submodule("envs.multiagent")
E, N = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
env = M.BatchedSwarmEnv(E, n_locusts=N, seed=1234, max_episode_steps=128)
env.reset()
a = torch.randn(E, 10, 2, device="cuda").clamp(-0.7, 0.7).contiguous()
v = torch.empty(E, N, 2, dtype=torch.float32, device="cuda")
r = torch.empty(E, dtype=torch.float32, device="cuda")
for _ in range(steps):
    env.step(a)
    env.forces(v=v, reward=r)
torch.cuda.synchronize()
print("plan", env.plan())
print("ok", E, N, steps, float(env.reward.mean()))
