#!/bin/bash
# ncu --set full of the decoder tail / resize kernels at C2 shapes (one launch each)
set -u
mkdir -p gpurun_out
TAG=${1:-tail}
OPS_N=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tail|upsample" -c 12 -f -o gpurun_out/${TAG}_tail \
   python scripts/gpu_decoder_ops.py ${TAG}_ncu > gpurun_out/${TAG}_ncu.log 2>&1; echo "exit $?"; tail -3 gpurun_out/${TAG}_ncu.log
