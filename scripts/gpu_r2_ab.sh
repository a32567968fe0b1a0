#!/bin/bash
# one call: the whole GPU suite on the in-tree library, then a same-box A/B of .variants/*.so
#   bash scripts/gpu_r2_ab.sh TAG "shapes" "variant names"
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
TAG=$1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > $O/pytest_$TAG.log 2>&1; echo "pytest_rc=$?" >> $O/pytest_$TAG.log
tail -8 $O/pytest_$TAG.log
bash scripts/gpu_variants.sh $TAG "$2" "$3"
