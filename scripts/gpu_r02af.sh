#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02af}
OPS_N=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdfr_conv3 -c 6 -f -o gpurun_out/${TAG}_conv \
   python scripts/gpu_decoder_ops.py ${TAG}_ncu > gpurun_out/${TAG}_ncu.log 2>&1; echo "exit $?"; tail -3 gpurun_out/${TAG}_ncu.log
