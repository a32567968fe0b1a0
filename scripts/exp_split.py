"""Where a FOLLOW-shape batch-step goes: the step with and without the rasteriser (CUDA graph of 8 steps, steady state),
the force kernel alone and the standalone rasteriser alone.   python scripts/exp_split.py 4096x256 2048x256 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import golds_rl_gym_b200 as pkg
from sweep import timed
M = pkg.submodule("envs.multiagent")

def graph(env, acts, **kw):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(8): env.step(acts[i], **kw)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(8): env.step(acts[i], **kw)
    return g

for s in sys.argv[1:]:
    E, N = (int(v) for v in s.split("x"))
    env = M.BatchedSwarmEnv(E, n_locusts=N, seed=1234, max_episode_steps=128, binding=os.environ.get("SWARM_BINDING", "ctypes"))
    env.reset()
    acts = []
    for i in range(8):
        a = torch.randn(E, 10, 2, device="cuda"); n = a.norm(dim=-1, keepdim=True)
        acts.append(torch.where(n >= 1.0, a / n, a).contiguous())
    out = {}
    for name, kw in (("step+raster", {}), ("step only", {"rasterize": False})):
        g = graph(env, acts, **kw)
        best = 1e9
        for r in range(5):
            env.reset(); g.replay()
            best = min(best, timed(g.replay, 10) / 80)
        out[name] = best * 1e3
    f = lambda: env.observe()
    for _ in range(3): f()
    out["k_rasterize alone"] = min(timed(f, 20) / 20 for _ in range(3)) * 1e3
    v = torch.empty(E, N, 2, dtype=torch.float32, device="cuda"); r = torch.empty(E, dtype=torch.float32, device="cuda")
    f = lambda: env.forces(v=v, reward=r)
    for _ in range(3): f()
    out["k_forces alone"] = min(timed(f, 20) / 20 for _ in range(3)) * 1e3
    print(s, "  ".join("%s %.2f us" % kv for kv in out.items()), flush=True)
