#!/bin/bash
# A/B of .variants/*.so (ctypes binding) against the round-1 tree on the same box:  bash scripts/gpu_variants.sh TAG "shapes" "variant names"
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
TAG=$1; SHAPES=$2
R1=$(echo "$SHAPES" | tr ' ' '\n' | cut -d: -f1 | sort -u | tr '\n' ' ')
rm -f $O/var_${TAG}_*.jsonl
(cd .r1_baseline && timeout 600 python scripts/sweep.py $R1 --json ../$O/var_${TAG}_r1.jsonl > ../$O/var_${TAG}_r1.log 2>&1)
for v in $3; do
  SWARM_B200_LIB=$PWD/.variants/libswarm_$v.so timeout 600 python scripts/sweep.py $SHAPES --binding ctypes --json $O/var_${TAG}_$v.jsonl > $O/var_${TAG}_$v.log 2>&1 || tail -3 $O/var_${TAG}_$v.log
done
(cd .r1_baseline && timeout 600 python scripts/sweep.py $R1 --json ../$O/var_${TAG}_r1b.jsonl > ../$O/var_${TAG}_r1b.log 2>&1)
python - <<PY
import json, glob, os
def load(f): return {d.get("shape", "%dx%d" % (d["E"], d["N"])): d for d in map(json.loads, open(f))}
r1 = load("$O/var_${TAG}_r1.jsonl"); r1b = load("$O/var_${TAG}_r1b.jsonl")
names = "$3".split()
vs = {v: load("$O/var_${TAG}_%s.jsonl" % v) for v in names if os.path.exists("$O/var_${TAG}_%s.jsonl" % v)}
shapes = "$SHAPES".split()
print("%-22s %9s %9s " % ("shape", "r1", "r1(again)") + " ".join("%12s" % v for v in vs))
for s in shapes:
    b = r1.get(s.split(":")[0]); b2 = r1b.get(s.split(":")[0])
    line = "%-22s %9.2f %9.2f " % (s, b["us_steady"] if b else 0, b2["us_steady"] if b2 else 0)
    for v in vs:
        d = vs[v].get(s)
        line += " %7.2f(%+4.1f)" % (d["us_steady"], 100 * (d["us_steady"] / b["us_steady"] - 1)) if d and b else "      -     "
    print(line)
PY
