#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02ag}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sdfr_forward_kernel" -c 12 -f -o gpurun_out/${TAG}_prof \
   python bench.py --steps 2 --warmup 1 --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_bench.log 2>&1; echo "exit $?"
