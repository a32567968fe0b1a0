#!/bin/bash
# PDL on / off / round-1 on the same box:  bash scripts/gpu_pdl.sh TAG "shapes"
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
TAG=$1; SHAPES=$2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_production.py tests/test_paac_gpu.py -m gpu -x -q --timeout 300 > $O/pytest_$TAG.log 2>&1; echo "pytest_rc=$?" >> $O/pytest_$TAG.log
tail -4 $O/pytest_$TAG.log
rm -f $O/pdl_${TAG}_*.jsonl
timeout 600 python scripts/sweep.py $SHAPES --boundary --json $O/pdl_${TAG}_on.jsonl > $O/pdl_${TAG}_on.log 2>&1 || tail -3 $O/pdl_${TAG}_on.log
SWARM_B200_NO_PDL=1 timeout 600 python scripts/sweep.py $SHAPES --boundary --json $O/pdl_${TAG}_off.jsonl > $O/pdl_${TAG}_off.log 2>&1
R1=$(echo "$SHAPES" | tr ' ' '\n' | cut -d: -f1 | sort -u | tr '\n' ' ')
(cd .r1_baseline && timeout 600 python scripts/sweep.py $R1 --boundary --json ../$O/pdl_${TAG}_r1.jsonl > ../$O/pdl_${TAG}_r1.log 2>&1)
python - <<PY
import json
def load(f): return {d.get("shape", "%dx%d" % (d["E"], d["N"])): d for d in map(json.loads, open(f))}
on, off, r1 = load("$O/pdl_${TAG}_on.jsonl"), load("$O/pdl_${TAG}_off.jsonl"), load("$O/pdl_${TAG}_r1.jsonl")
print("%-20s %10s %10s %10s   %s" % ("shape", "r1", "no PDL", "PDL", "episode-avg r1 / PDL   max-8-step r1 / PDL"))
for s in "$SHAPES".split():
    b = r1[s.split(":")[0]]
    print("%-20s %10.2f %10.2f %10.2f (%+5.1f %%)   %7.2f / %7.2f (%+5.1f %%)   %7.2f / %7.2f" % (s, b["us_steady"], off[s]["us_steady"], on[s]["us_steady"],
          100 * (on[s]["us_steady"] / b["us_steady"] - 1), b["us_episode"], on[s]["us_episode"], 100 * (on[s]["us_episode"] / b["us_episode"] - 1),
          b["us_replay_max"], on[s]["us_replay_max"]))
PY
