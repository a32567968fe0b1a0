"""Experiments: build variants of libswarm_b200.so with compile-time switches into .variants/ (git-ignored, shipped to the
GPU box) so that one gpurun call can A/B them:  python scripts/build_variants.py NAME=-DFLAG[,-DFLAG2] ...
Use with SWARM_B200_LIB=.variants/libswarm_NAME.so python scripts/sweep.py --binding ctypes ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, ".variants"); os.makedirs(out, exist_ok=True)
src = os.path.join(ROOT, "golds-rl-gym_b200", "csrc", "swarm_b200.cu")
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-shared"]
    cmd += [f for f in flags.split(",") if f] + ["-o", os.path.join(out, "libswarm_%s.so" % name), src]
    procs.append((name, subprocess.Popen(cmd)))
for name, p in procs:
    print(name, "rc", p.wait())
