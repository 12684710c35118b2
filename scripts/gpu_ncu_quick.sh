#!/bin/bash
# quick ncu capture of selected kernels of the bench step; usage: gpu_ncu_quick.sh TAG 'regex'
set -u
mkdir -p gpurun_out
TAG=${1:-q}; RX=${2:-sdfr_}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 4 -c 6 -f -o gpurun_out/${TAG}_prof \
   python bench.py --steps 2 --warmup 1 --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1; echo "exit $?"
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}_prof.ncu-rep
