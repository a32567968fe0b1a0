#!/bin/bash
# round 2, call 2: parity of the K-split / self-raster shapes + their speed
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > $O/pytest_gpu_2.log 2>&1; echo "pytest_rc=$?" >> $O/pytest_gpu_2.log
tail -15 $O/pytest_gpu_2.log
timeout 900 python scripts/sweep.py 512x256:ks1:follow 512x256:ks1:self 512x256:ks2:self 512x256:ks2:warps 512x256:ks4:self 512x256 \
   1024x256:ks1:self 1024x256:ks2:self 1024x256:ks1:follow 1024x256 2048x256 2048x256:ks1:self 4096x256 4096x256:ks1:self 4096x256:ks1:warps \
   1024x64:ks1:warps 1024x64:ks1:self 1024x64:ks2:self 1024x64:ks4:self 1024x64:ks4:warps 1024x64 \
   1024x80:warps 1024x80:self 1024x80 4096x80 4096x80:self 32x80 32x80:warps 4096x64 4096x64:ks1:self 4096x64:ks1:warps \
   --json $O/sweep1.jsonl > $O/sweep1.log 2>&1
echo "sweep rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r2/sweep1.jsonl"):
    d=json.loads(l); print("%-22s steady %7.2f us  forces %7.2f us  xu %.3f" % (d["shape"], d["us_steady"], d["us_forces"], d["xu_frac_steady"]))
PY
