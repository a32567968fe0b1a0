#!/bin/bash
# round 2, call 1: baseline numbers of the regimes VERDICT r01 names (small E at N=256, N=80, C2) + sanitizer evidence
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 600 python scripts/sweep.py 512x256 1024x256 2048x256 4096x256 1024x64 1024x80 4096x80 32x80 4096x64 --boundary --json $O/sweep0.jsonl > $O/sweep0.log 2>&1
echo "sweep rc=$?"; cat $O/sweep0.log | tail -12
for s in "1024 80" "1024 64" "512 256"; do
  set -- $s
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 6 -c 1 -f \
      -o $O/prof_k_step_${1}x${2} python scripts/run_steps.py $1 $2 > $O/ncu_${1}x${2}.log 2>&1
  echo "ncu $s rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_raster_follow --launch-skip 6 -c 1 -f \
      -o $O/prof_k_raster_follow_512x256 python scripts/run_steps.py 512 256 > $O/ncu_follow_512x256.log 2>&1
echo "ncu follow rc=$?"
SAN=/usr/local/cuda/bin/compute-sanitizer
K='follower or autoreset or graph_replay'
timeout 900 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > $O/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 $O/sanitizer_memcheck.log
timeout 900 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > $O/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 $O/sanitizer_racecheck.log
ls -la $O
