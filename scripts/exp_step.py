"""Experiment: time the fused step with / without the raster group (C4 and C2 shapes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import golds_rl_gym_b200 as pkg
M = pkg.submodule("envs.multiagent")

def run(E, N, raster, steps=100):
    env = M.BatchedSwarmEnv(E, n_locusts=N, seed=1, rasterize=raster, auto_reset=True)
    env.reset()
    a = torch.randn(E, 10, 2, device="cuda").clamp(-0.7, 0.7)
    for _ in range(10): env.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): env.step(a)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

for E, N in [(4096, 256), (1024, 64), (4096, 64), (4096, 80), (32, 80), (1024, 512), (256, 1024), (64, 2048)]:
    print(E, N, "raster: %.1f us   no raster: %.1f us" % (run(E, N, True), run(E, N, False)))
