"""Per-phase timeline of one batch-step (debug facility swarm_debug_trace).  Needs a tracing build of the library:
    python scripts/build_variants.py trace=-DSWARM_TRACE
    SWARM_B200_LIB=$PWD/.variants/libswarm_trace.so python scripts/trace_step.py 512x256[:follow|warps|self] ...
For every CTA's first env: ns since the kernel's first CTA entered, per phase (median / p90 / max over CTAs)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import golds_rl_gym_b200 as pkg
M = pkg.submodule("envs.multiagent")
nat = M.nat
PLACE = {"follow": 1 << 4, "warps": 2 << 4, "self": 3 << 4}
PH = ["entry", "zfill", "loaded", "staged", "tiles", "forces", "stepped", "stored", "raster", "mean", "binned", "zeros", "done",
      "-", "-", "-"]
lib = nat.load()
for s in sys.argv[1:]:
    parts = s.split(":")
    E, N = (int(v) for v in parts[0].split("x"))
    tuning = 0
    raster = "noraster" not in parts[1:]
    for t in parts[1:]:
        tuning |= PLACE.get(t, 0)
    env = M.BatchedSwarmEnv(E, n_locusts=N, seed=1234, max_episode_steps=128, tuning=tuning, binding="ctypes")
    env.reset()
    print("plan", env.plan())
    a = torch.randn(E, 10, 2, device="cuda").clamp(-0.7, 0.7).contiguous()
    for _ in range(6):
        env.step(a, rasterize=raster)
    buf = torch.zeros(2 * E * 32, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    lib.swarm_debug_trace(ctypes.c_void_p(buf.data_ptr()), buf.numel())
    env.step(a, rasterize=raster)
    torch.cuda.synchronize()
    lib.swarm_debug_trace(None, 0)
    t = buf.cpu().numpy().reshape(2 * E, 16, 2)
    gt = t[:, :, 0].astype(np.float64)
    t0 = gt[gt > 0].min()
    life = (t[:E, 7, 0] - t[:E, 0, 0]).astype(np.float64)
    ent = np.sort(t[:E, 0, 0].astype(np.float64) - t0)
    print("== %s  CTA lifetime entry->stored: med %.0f p10 %.0f p90 %.0f ns; span %.0f ns; mean resident CTAs/SM %.2f; entries by 10%% quantile: %s" % (
        s, np.median(life), np.percentile(life, 10), np.percentile(life, 90), (t[:E, 7, 0].max() - t0), life.sum() / (t[:E, 7, 0].max() - t0) / 148,
        " ".join("%.0f" % ent[int(q * (E - 1) / 10)] for q in range(11))))
    print("== %s  (globaltimer ns since first entry; SM cycles between consecutive phases of the same record in [])" % s)
    for name, rows in (("step CTAs", t[:E]), ("follower envs", t[E:])):
        if not (rows[:, :, 0] > 0).any():
            continue
        for ph in range(16):
            v = rows[:, ph, 0].astype(np.float64)
            ok = v > 0
            if not ok.any():
                continue
            d = v[ok] - t0
            line = "  %-13s %-8s n=%5d  med %8.0f  p90 %8.0f  max %8.0f" % (name, PH[ph], ok.sum(), np.median(d), np.percentile(d, 90), d.max())
            prev = [q for q in range(ph) if (rows[:, q, 0] > 0).any()]
            if prev:
                q = prev[-1]
                both = ok & (rows[:, q, 0] > 0)
                cyc = (rows[both, ph, 1] - rows[both, q, 1]).astype(np.float64)
                line += "   [+%6.0f cyc med since %s, max %6.0f]" % (np.median(cyc), PH[q], cyc.max())
            print(line)
    del env
