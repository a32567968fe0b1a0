#!/bin/bash
# round-2 ncu evidence: launch list of the bench command + one --set full capture per hot kernel / shape (each after the
# same program has exited 0 without ncu).  Everything lands in gpurun_out/r2/; scripts/ncu_r2_summaries.sh turns it into profiles/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
BENCH="python bench.py --steps 8 --warmup 3 --no-graph --no-cpu-baseline --no-paac --no-secondary"
timeout 600 $BENCH > $O/bench_nograph.json 2>&1; echo "bench(no graph) rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02.csv $BENCH > $O/ncu_launches_r02.log 2>&1; echo "launch list rc=$?"
cap() {  # cap NAME KERNEL_REGEX E N
  timeout 300 python scripts/run_steps.py $3 $4 > $O/plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip 6 -c 1 -f -o $O/prof_r02_$1 \
      python scripts/run_steps.py $3 $4 > $O/ncu_r02_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap k_step_c4 k_step 4096 256
cap k_raster_follow_c4 k_raster_follow 4096 256
cap k_step_512x256 k_step 512 256
cap k_step_n80 k_step 1024 80
cap k_step_c2 k_step 1024 64
cap k_forces_c4 k_forces 4096 256
# summarise ON the box (the .ncu-rep files together exceed what gpurun copies back), keep the two headline captures
bash scripts/ncu_r2_summaries.sh > $O/summaries.log 2>&1; tail -3 $O/summaries.log
mkdir -p $O/profiles_out; cp profiles/r02_* $O/profiles_out/ 2>/dev/null
for n in k_step_c4 k_raster_follow_c4 k_step_512x256 k_step_n80 k_step_c2; do
  python scripts/ncu_lines.py $O/prof_r02_$n.ncu-rep 25 > $O/profiles_out/r02_lines_$n.txt 2>&1
done
rm -f $O/prof_r02_k_forces_c4.ncu-rep $O/prof_r02_k_raster_follow_c4.ncu-rep $O/prof_r02_k_step_n80.ncu-rep $O/prof_r02_k_step_c2.ncu-rep
ls -la $O | grep -i "r02\|profiles_out"
