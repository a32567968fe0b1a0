"""Step latency of the fused batch-step over a list of (E, N) shapes: CUDA graph of 8 steps replayed, steady state
(no episode boundary inside the timed region) and, with --boundary, the whole 128-step episode incl. the auto-reset.

    python scripts/sweep.py 512x256 1024x256:ks2:self 4096x256 1024x64 1024x80 [--boundary] [--json out.jsonl]
(ExN[:ksK][:follow|warps|self] forces the launch shape through SwarmParams.tuning)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import golds_rl_gym_b200 as pkg

M = pkg.submodule("envs.multiagent")


def graph_of(env, acts, n=8):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n):
            env.step(acts[i % len(acts)])
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(n):
            env.step(acts[i % len(acts)])
    return g


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def run(E, N, boundary, reps=5, **kw):
    if not kw.get("tuning"):
        kw.pop("tuning", None)          # (the round-1 tree under .r1_baseline has no such argument)
    env = M.BatchedSwarmEnv(E, n_locusts=N, seed=1234, max_episode_steps=128, **kw)
    env.reset()
    acts = []
    for i in range(8):
        a = torch.randn(E, 10, 2, device="cuda")
        n = a.norm(dim=-1, keepdim=True)
        acts.append(torch.where(n >= 1.0, a / n, a).contiguous())
    g = graph_of(env, acts)         # 16 steps so far
    out = {"E": E, "N": N}
    # steady state: 8 replays = 64 steps inside one episode, best of reps
    best = 1e9
    for r in range(reps):
        env.reset()
        g.replay()                  # 8 steps of warm-up
        best = min(best, timed(g.replay, 10) / 80)
    out["us_steady"] = best * 1e3
    if boundary:                    # 2 whole episodes = 256 steps = 32 replays
        env.reset()
        out["us_episode"] = timed(g.replay, 32) / 256 * 1e3
        # worst single step of an episode: per-replay (8-step) timing
        env.reset()
        ts = [timed(g.replay, 1) / 8 * 1e3 for _ in range(32)]
        out["us_replay_max"] = max(ts)
        out["us_replay_med"] = sorted(ts)[len(ts) // 2]
    v = torch.empty(E, N, 2, dtype=torch.float32, device="cuda")
    r = torch.empty(E, dtype=torch.float32, device="cuda")
    f = lambda: env.forces(v=v, reward=r)
    for _ in range(3):
        f()
    out["us_forces"] = min(timed(f, 20) / 20 for _ in range(3)) * 1e3
    pairs = E * N * (N + 10)
    # unordered-pair modes (N <= 512): 3 MUFU per unordered pair, else 4
    out["xu_frac_steady"] = (1.5 if N <= 512 else 2.0) * pairs / (out["us_steady"] * 1e-6) / (148 * 15.93 * 1965e6)
    out["locust_updates_per_s"] = E * N / (out["us_steady"] * 1e-6)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("shapes", nargs="+")
    ap.add_argument("--boundary", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--binding", default="torch")
    args = ap.parse_args()
    PLACE = {"follow": 1 << 4, "warps": 2 << 4, "self": 3 << 4}
    for s in args.shapes:
        parts = s.split(":")
        E, N = (int(v) for v in parts[0].split("x"))
        tuning = 0
        for t in parts[1:]:
            tuning |= PLACE.get(t, 0)
        res = run(E, N, args.boundary, binding=args.binding, tuning=tuning)
        res["shape"] = s
        print(json.dumps(res), flush=True)
        if args.json:
            with open(args.json, "a") as fh:
                fh.write(json.dumps(res) + "\n")
