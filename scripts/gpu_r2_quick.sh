#!/bin/bash
# quick loop: parity of the launch shapes + traces + a short sweep.   bash scripts/gpu_r2_quick.sh TAG "trace shapes" "sweep shapes" ["pytest -k expr"]
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
TAG=$1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_production.py -m gpu -x -q --timeout 300 ${4:+-k "$4"} > $O/pytest_$TAG.log 2>&1; echo "pytest_rc=$?" >> $O/pytest_$TAG.log
tail -6 $O/pytest_$TAG.log
if [ -n "${2:-}" ]; then timeout 300 python scripts/trace_step.py $2 > $O/trace_$TAG.log 2>&1; grep -v "^plan" $O/trace_$TAG.log; fi
if [ -n "${3:-}" ]; then rm -f $O/sweep_$TAG.jsonl $O/sweep_r1_$TAG.jsonl; timeout 600 python scripts/sweep.py $3 --json $O/sweep_$TAG.jsonl > $O/sweep_$TAG.log 2>&1
 if [ -d .r1_baseline ]; then R1=$(echo "$3" | tr ' ' '\n' | cut -d: -f1 | sort -u | tr '\n' ' '); (cd .r1_baseline && timeout 600 python scripts/sweep.py $R1 --json ../$O/sweep_r1_$TAG.jsonl > ../$O/sweep_r1_$TAG.log 2>&1); fi
 python - <<PY
import json, os
r1 = {}
if os.path.exists("$O/sweep_r1_$TAG.jsonl"):
    r1 = {(d["E"], d["N"]): d for d in map(json.loads, open("$O/sweep_r1_$TAG.jsonl"))}
for l in open("$O/sweep_$TAG.jsonl"):
    d=json.loads(l); b = r1.get((d["E"], d["N"]))
    print("%-24s steady %7.2f us  forces %7.2f us  xu %.3f %s" % (d["shape"], d["us_steady"], d["us_forces"], d["xu_frac_steady"],
          ("  | r1 same box: %7.2f us (%+.1f %%)" % (b["us_steady"], 100*(d["us_steady"]/b["us_steady"]-1))) if b else ""))
PY
fi
