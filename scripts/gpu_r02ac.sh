#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02ac}
echo "== configs c1"; SKIP_C3=1 SKIP_C5=1 timeout 600 python scripts/gpu_configs.py ${TAG} > gpurun_out/${TAG}_configs.log 2>&1; echo "exit $?"; tail -60 gpurun_out/${TAG}_configs.log
echo "== micro"; timeout 900 python scripts/gpu_micro.py ${TAG} > gpurun_out/${TAG}_micro.log 2>&1; echo "exit $?"; head -30 gpurun_out/${TAG}_micro.log
