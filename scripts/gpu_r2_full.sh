#!/bin/bash
# full check: every GPU test, smoke, the default bench line (+ reference arm), A/B against the round-1 kernels
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
TAG=${1:-full}
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > $O/pytest_$TAG.log 2>&1; echo "pytest_rc=$?" >> $O/pytest_$TAG.log
tail -5 $O/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke_rc=$?" >> $O/smoke_$TAG.log; tail -3 $O/smoke_$TAG.log
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -c 3000 $O/bench_$TAG.json; tail -5 $O/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_$TAG.json 2>&1; echo "ref rc=$?"; tail -c 600 $O/bench_ref_$TAG.json
if [ -d .r1_baseline ]; then
  (cd .r1_baseline && timeout 600 python scripts/sweep.py 512x256 1024x256 4096x256 1024x64 1024x80 4096x80 --json ../$O/sweep_r1base_$TAG.jsonl > ../$O/sweep_r1base_$TAG.log 2>&1)
  timeout 600 python scripts/sweep.py 512x256 1024x256 4096x256 1024x64 1024x80 4096x80 --json $O/sweep_now_$TAG.jsonl > $O/sweep_now_$TAG.log 2>&1
  python - <<PY
import json
a={(d["E"],d["N"]):d for d in map(json.loads,open("$O/sweep_r1base_$TAG.jsonl"))}
b={(d["E"],d["N"]):d for d in map(json.loads,open("$O/sweep_now_$TAG.jsonl"))}
for k in a:
    if k in b: print("%5dx%-4d r1 %7.2f us  now %7.2f us  (%+.1f %%)   forces r1 %7.2f now %7.2f" % (k[0],k[1],a[k]["us_steady"],b[k]["us_steady"],100*(b[k]["us_steady"]/a[k]["us_steady"]-1),a[k]["us_forces"],b[k]["us_forces"]))
PY
fi
