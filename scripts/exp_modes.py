"""Compare the two ways swarm_step produces the observation: follower kernel (work given) vs raster warps inside k_step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import golds_rl_gym_b200 as pkg
M = pkg.submodule("envs.multiagent")

def run(E, N, fused, steps=60):
    env = M.BatchedSwarmEnv(E, n_locusts=N, seed=1, binding="ctypes")
    if fused:
        env.state_c.work = None
    env.reset()
    a = torch.randn(E, 10, 2, device="cuda").clamp(-0.7, 0.7)
    for _ in range(10): env.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): env.step(a)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

for E, N in [(4096, 96), (4096, 128), (4096, 160), (4096, 192), (4096, 256), (2048, 384), (1024, 512), (512, 700), (256, 1024)]:
    print(E, N, "auto: %.1f us   in-kernel raster warps: %.1f us" % (run(E, N, False), run(E, N, True)))
