#!/bin/bash
# quick loop: parity of the launch shapes + traces + a short sweep.   bash scripts/gpu_r2_quick.sh TAG "trace shapes" "sweep shapes"
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
TAG=$1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -x -q --timeout 300 > $O/pytest_$TAG.log 2>&1; echo "pytest_rc=$?" >> $O/pytest_$TAG.log
tail -4 $O/pytest_$TAG.log
if [ -n "${2:-}" ]; then timeout 300 python scripts/trace_step.py $2 > $O/trace_$TAG.log 2>&1; grep -v "^plan" $O/trace_$TAG.log; fi
if [ -n "${3:-}" ]; then rm -f $O/sweep_$TAG.jsonl; timeout 600 python scripts/sweep.py $3 --json $O/sweep_$TAG.jsonl > $O/sweep_$TAG.log 2>&1; python - <<PY
import json
for l in open("$O/sweep_$TAG.jsonl"):
    d=json.loads(l); print("%-22s steady %7.2f us  forces %7.2f us  xu %.3f" % (d["shape"], d["us_steady"], d["us_forces"], d["xu_frac_steady"]))
PY
fi
