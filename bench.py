#!/usr/bin/env python
"""Benchmark of the swarm hot path (BASELINE.json metric: swarm env-steps/s & locust-updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one batch-step: every env of the rank's shard advances once through the fused kernel
(SwarmEnv._step + TimeLimit + SwarmRunner auto-reset + 84x84 compact rasterise).  Weak scaling:
each GPU owns E envs (global env ids keep the Philox streams shard-invariant); there is no
data-path collective.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3]: the configuration the metric's target is quoted on
    "c4": dict(E=4096, N=256, name="C4 large-swarm stress: 4096 envs x 256 locusts per GPU, all-pairs forces"),
    # BASELINE.json configs[1]
    "c2": dict(E=1024, N=64, name="C2 batched swarm env: 1024 envs x 64 locusts per GPU"),
}
PAAC = {
    # BASELINE.json configs[2] / configs[4]
    "paac3": dict(E=32, name="C3 PAAC conv training as in scripts/train_paac_conv.py (--height=84 --clip_norm=1), 32 emulators"),
    "paac5": dict(E=1024, name="C5 PAAC conv training, 1024 emulators per GPU (8192 over 8 GPUs), NCCL gradient all-reduce"),
}
A, G = 10, 84
FLOPS_PER_PAIR = 18


def ncu_traffic(kernels=("k_step", "k_raster_follow")):
    """dram read+write bytes per batch-step (the step kernel + its concurrent follower), from the committed
    ncu --set full summary (profiles/rNN_summary.json, newest round)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        for name in sorted(os.listdir(pdir)):
            if name.endswith("_summary.json"):
                with open(os.path.join(pdir, name)) as f:
                    summ = json.load(f)
                tot = 0.0
                for k in kernels:
                    d = summ.get(k, {})
                    if "dram__bytes_read.sum" in d:
                        tot += (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e6   # ncu: Mbyte
                if tot:
                    best = tot
    except Exception:
        pass
    return best


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), sm_max_mhz=float(p["sm_max_mhz"]), source="measured")
    except Exception:
        return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler(object):
    """nvidia-smi polled every 20 ms from BEFORE the warm-up (it needs ~0.2 s to start); only the samples whose
    timestamps fall inside the timed region count."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self, t_begin, t_end):
        import datetime
        if self.proc is None:
            return None
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return None
        rows, mx = [], 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), [n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
        if not rows:
            return None
        inside = [r for r in rows if t_begin - 0.02 <= r[0] <= t_end + 0.02]
        where = "timed region"
        if not inside:                      # region shorter than the polling period: fall back to the busiest samples
            inside = sorted(rows, key=lambda r: r[1])[len(rows) // 2:]
            where = "whole run (timed region shorter than the 20 ms polling period)"
        reasons = sorted({n for r in inside for n in r[2]})
        return dict(sm_mhz=statistics.median([r[1] for r in inside]), sm_max_mhz=mx, reasons=reasons,
                    samples=len(inside), sampled="nvidia-smi -lms 20, " + where)


def run_reference(args, wl):
    """--impl reference: the CPU port of the reference env + rasteriser on all host cores, on the same
    workload shape; each step is a bounded sample (2 envs per core)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    res = cb.time_port(wl["N"], steps=args.steps, warmup=args.warmup, envs_per_proc=2)
    value = res["env_steps_per_s"] * wl["N"]
    unit = "locust-updates/s"
    sample = "%d envs (2 per core) x %d steps of %s, NumPy FP64 port, reset excluded" % (res["envs"], res["steps"], wl["name"])
    print(json.dumps({
        "impl": "reference", "metric": "locust_updates_per_sec", "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * res["seconds"] / max(1, res["steps"]), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"] + " (bounded sample)", "envs_timed": res["envs"], "n_locusts": wl["N"],
                   "n_agents": A, "grid": G},
        "env_steps_per_sec": res["env_steps_per_s"],
        "cpu_baseline": {"value": value, "unit": unit, "cores": res["procs"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def paac_frames_per_sec(E_per_gpu, updates, warmup_updates, world, rank, local):
    """PAAC frames/s (= T*E / loop time, paac.py:397-401) of the device-resident learner, CUDA-event timed."""
    import torch
    import golds_rl_gym_b200 as pkg
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import train_paac_conv as tp
    args = tp.get_arg_parser().parse_args(["--height=84", "--clip_norm=1", "-ec", str(E_per_gpu * world)])
    net_creator, env_creator = tp.get_network_and_environment_creator(args)
    learner = pkg.submodule("agents.paac.paac").GridPAACLearner(net_creator, env_creator, args)
    learner.start()
    for _ in range(warmup_updates):
        learner.update()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(updates):
        learner.update()
    e1.record()
    torch.cuda.synchronize()
    sh = pkg.submodule("sharding")
    ms = sh.max_over_ranks(e0.elapsed_time(e1), device=torch.device("cuda", local))
    frames = updates * learner.max_local_steps * learner.total_emulators
    return dict(frames_per_sec=frames / (ms * 1e-3), ms_per_update=ms / updates, updates=updates,
                emulators=learner.total_emulators, local_steps=learner.max_local_steps,
                policy_batch=learner.real_batch_size,
                observation="compact (grid + positions; conv1 factorised over the shared channels)" if learner.compact_obs
                else "expanded (E,A,84,84,3)", net_dtype="fp32 (cuDNN TF32 convolutions, FP32 dense)",
                launch="one CUDA graph per update (rollout + returns + backward + all-reduce + Adam)")


def run_paac(args, wl):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = paac_frames_per_sec(wl["E"], max(args.steps, 8), max(args.warmup, 3), world, rank, local)
    if rank == 0:
        print(json.dumps({
            "metric": "paac_frames_per_sec", "value": res["frames_per_sec"], "unit": "frames/s", "n_gpus": world,
            "steps": res["updates"], "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_update"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": res["net_dtype"],
            "data": "synthetic", "config": {"workload": wl["name"], "emulators_per_gpu": wl["E"], "n_locusts": 80,
                                            "n_agents": A, "grid": G, "local_steps": res["local_steps"],
                                            "parallelism": "env-sharded x%d, flat-gradient NCCL all-reduce" % world},
            "paac": res, "gpu_launches": res["updates"] * res["local_steps"] * 3}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + sorted(PAAC))
    ap.add_argument("--math", default="fast", choices=["fast", "precise"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary C2 figure (1024 envs x 64 locusts)")
    ap.add_argument("--no-paac", action="store_true", help="skip the secondary PAAC frames/s figure (config 3)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of replaying a CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the CPU-baseline sample")
    args = ap.parse_args()
    if args.workload in PAAC:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference learner needs TensorFlow 1.4 (not in this image)"}))
            return
        return run_paac(args, PAAC[args.workload])
    wl = WORKLOADS[args.workload]
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    import golds_rl_gym_b200 as pkg
    M = pkg.submodule("envs.multiagent")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, N = wl["E"], wl["N"]
    K, W = args.steps, args.warmup

    env = M.BatchedSwarmEnv(E, n_locusts=N, n_agents=A, grid_size=G, max_episode_steps=128, seed=1234,
                            env_id_offset=rank * E, device=dev, math_mode=args.math, auto_reset=True, rasterize=True)
    env.reset()
    # inputs resident in HBM: a ring of pre-clipped N(0,1) actions (seeded per rank)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    ring = []
    for _ in range(8):
        a = torch.randn(E, A, 2, device=dev, generator=gen)
        n = a.norm(dim=-1, keepdim=True)
        ring.append(torch.where(n >= 1.0, a / n, a).contiguous())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wall = [0.0, 0.0]         # host clock around the last timed loop (for the clock sampler)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall[0] = time.time()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        wall[1] = time.time()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local) if rank == 0 else None      # started early: nvidia-smi takes a moment to come up
    step = lambda i: env.step(ring[i & 7])
    for i in range(W):
        step(i)
    # The device-resident rollout replays a CUDA graph of 8 consecutive steps (one per action tensor of
    # the ring): same kernels, same arguments, no per-step host work.  K is rounded up to a multiple of 8.
    graph = None
    if not args.no_graph:
        torch.cuda.synchronize()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(8):
                step(i)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for i in range(8):
                step(i)
        K = ((K + 7) // 8) * 8
    if graph is not None:
        ms = timed(lambda i: graph.replay(), K // 8)
    else:
        ms = timed(step, K)
    clocks = sampler.stop(wall[0], wall[1]) if sampler else None

    env_steps = world * E * K / (ms * 1e-3)
    pairs_per_launch = E * N * (N + A)
    pk = peaks()
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12          # TFLOP/s
    achieved = pairs_per_launch * FLOPS_PER_PAIR / (ms / K * 1e-3) / 1e12
    # algorithmic HBM bytes of the fused launch: state r/w + frozen noise + actions + obs
    bytes_per_env = (N + A) * 16 * 3 + A * 8 + 5 + 8 + (G * G * 2 * 4 + A * 2)
    hbm = E * bytes_per_env / (ms / K * 1e-3) / 1e9

    # force kernel alone (swarm_forces = SwarmEnv.v_calculate for the batch)
    v = torch.empty(E, N, 2, dtype=torch.float32, device=dev)
    r = torch.empty(E, dtype=torch.float32, device=dev)
    force = lambda i: env.forces(v=v, reward=r)
    for i in range(3):
        force(i)
    ms_f = timed(force, max(10, K // 4)) / max(10, K // 4)
    # rasteriser alone
    rast = lambda i: env.observe()
    for i in range(3):
        rast(i)
    ms_r = timed(rast, max(10, K // 4)) / max(10, K // 4)

    # end to end through the host-buffer C-ABI call: actions from pinned host memory, reward/done back
    h_act = [a.cpu().pin_memory() for a in ring]
    h_rew = torch.zeros(E, dtype=torch.float32).pin_memory()
    h_done = torch.zeros(E, dtype=torch.uint8).pin_memory()
    e2e_step = lambda i: env.step_host(h_act[i & 7], h_rew, h_done)
    for i in range(16):          # every host buffer of the ring once or more (the library caches a graph per argument set)
        e2e_step(i)
    barrier()
    w0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    torch.cuda.synchronize()
    w_ms = (time.perf_counter() - w0) * 1e3
    if world > 1:
        t = torch.tensor([w_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w_ms = float(t.item())
    e2e_val = world * E * K * N / (w_ms * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as cb
        probe = cb.time_port(N, steps=4, warmup=1, envs_per_proc=2)
        steps_cpu = int(max(8, min(20000, args.cpu_seconds * probe["steps"] / max(probe["seconds"], 1e-6))))
        res = cb.time_port(N, steps=steps_cpu, warmup=1, envs_per_proc=2)
        cpu = {"value": res["env_steps_per_s"] * N, "unit": "locust-updates/s", "cores": res["procs"], "kind": "port",
               "sample": "%d envs (2 per core) x %d steps, NumPy FP64 port of SwarmEnv.step + process_state, "
                         "one process per core; %.1f s" % (res["envs"], res["steps"], res["seconds"]),
               "env_steps_per_sec": res["env_steps_per_s"]}

    paac = None
    if world == 1 and not args.no_paac:
        try:
            del env
            torch.cuda.empty_cache()
            paac = paac_frames_per_sec(PAAC["paac3"]["E"], 200, 20, 1, 0, local)
            paac["config"] = PAAC["paac3"]["name"]
        except Exception as exc:      # the secondary figure must never take the headline down
            paac = {"error": repr(exc)}

    # BASELINE.json configs[1] (C2) next to the headline workload: measured by a child process of this very script
    secondary = None
    if world == 1 and rank == 0 and args.workload == "c4" and not args.no_secondary:
        try:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "c2", "--steps", "512", "--warmup",
                                  str(W), "--no-cpu-baseline", "--no-paac", "--no-secondary"], capture_output=True, text=True,
                                 timeout=300)
            d = json.loads(res.stdout.strip().splitlines()[-1])
            secondary = {"c2": {k: d[k] for k in ("value", "unit", "ms_per_step", "env_steps_per_sec", "e2e", "config")}}
        except Exception as exc:
            secondary = {"c2": {"error": repr(exc)}}

    if rank == 0:
        out = {
            "metric": "locust_updates_per_sec", "value": env_steps * N, "unit": "locust-updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 pair forces / f64 integrator state", "data": "synthetic",
            "config": {"workload": wl["name"] + "; fused step + TimeLimit(128) + auto-reset + 84x84 compact rasterise",
                       "envs_per_gpu": E, "n_locusts": N, "n_agents": A, "grid": G, "math": args.math,
                       "actions": "N(0,1) clipped to unit norm, 8 pre-generated device tensors",
                       "launch": "python per step" if args.no_graph else "CUDA graph of 8 steps replayed",
                       "l2": ("no explicit flush: each step streams %.0f MB of observations, more than the 126 MB L2"
                              if E * G * G * 8 > 126e6 else
                              "no explicit flush: each step streams %.0f MB of observations (smaller than the 126 MB L2, "
                              "which may absorb part of the write-back: secondary figure only)") % (E * G * G * 2 * 4 / 1e6),
                       "parallelism": "env-sharded x%d, no collective" % world},
            "env_steps_per_sec": env_steps, "pairs_per_sec": env_steps * N * (N + A),
            "roofline": {"bound": "fp32", "kernel": "k_step (fused step + rasterise)", "achieved": achieved, "peak": fp32_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "traffic": ncu_traffic() if args.workload == "c4" else None,
                         "xu_pipe": {"achieved_mufu_per_s": 2.0 * pairs_per_launch / (ms / K * 1e-3),
                                     "peak_mufu_per_s": 148 * 15.93 * pk["sm_max_mhz"] * 1e6,
                                     "frac": 2.0 * pairs_per_launch / (ms / K * 1e-3) / (148 * 15.93 * pk["sm_max_mhz"] * 1e6),
                                     "note": "the saturated pipe: 4 MUFU per UNORDERED pair = 2 per ordered pair; "
                                             "15.93 MUFU/clk/SM measured (scripts/microbench.cu)"},
                         "note": "compute-bound kernel (190 flop/B): algorithmic 18 flop/pair x %d ordered pairs/launch over the "
                                 "mean launch duration of the timed region; peak = 148 SM x 128 lanes x 2 x %.0f MHz (%s clock); "
                                 "traffic = ncu dram read+write bytes per launch (profiles/)"
                                 % (pairs_per_launch, pk["sm_max_mhz"], pk["source"])},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": hbm / pk["hbm_gbs"], "bytes_per_launch": E * bytes_per_env, "peak_source": pk["source"]},
            "force_kernel": {"ms": ms_f, "pairs_per_sec": pairs_per_launch / (ms_f * 1e-3),
                             "tflops_algorithmic": pairs_per_launch * FLOPS_PER_PAIR / (ms_f * 1e-3) / 1e12,
                             "frac_fp32_peak": pairs_per_launch * FLOPS_PER_PAIR / (ms_f * 1e-3) / 1e12 / fp32_peak},
            "raster_kernel": {"ms": ms_r, "gbs": E * (G * G * 2 * 4 + A * 2 + (N + A) * 16) / (ms_r * 1e-3) / 1e9,
                              "frac_hbm_peak": E * (G * G * 2 * 4 + A * 2 + (N + A) * 16) / (ms_r * 1e-3) / 1e9 / pk["hbm_gbs"]},
            "e2e": {"value": e2e_val, "unit": "locust-updates/s", "h2d_bytes_per_step": E * A * 2 * 4,
                    "d2h_bytes_per_step": E * 5, "ms_per_step": w_ms / K,
                    "note": "swarm_step_host with pinned HOST buffers, one call + stream sync per step, wall clock: the kernel "
                            "reads the step's actions from host memory over PCIe (cp.async prefetch, zero-copy) and posts "
                            "reward+done into host memory; observations stay in HBM for the device-resident policy"},
            "gpu_launches": K * (2 if N >= 160 else 1),      # k_step (+ k_raster_follow for large swarms) per step
            "clocks": clocks,
        }
        if cpu:
            out["cpu_baseline"] = cpu
        if paac is not None:
            out["paac"] = paac
        if secondary is not None:
            out["secondary"] = secondary
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
