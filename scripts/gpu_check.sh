#!/bin/bash
# Run on the GPU box (via gpurun): GPU parity tests, smoke, bench (C4 + C2), ncu launch list and
# one full ncu capture per hot kernel.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke_rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c4.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_c4.log
timeout 600 python bench.py --workload c2 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_c2.log
if [ "${1:-}" != "noprof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --no-paac --no-secondary > gpurun_out/ncu_launches.log 2>&1
for k in k_step k_raster_follow k_forces k_rasterize; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 -c 1 -f \
      -o gpurun_out/prof_$k python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graph --no-paac --no-secondary > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
fi
ls -la gpurun_out
