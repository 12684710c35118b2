#!/bin/bash
# C1 latency decomposition + C4 sweep with the categories as parallel graph branches (A/B on one GPU)
set -u
mkdir -p gpurun_out
TAG=${1:-r02ab}
echo "== c1 latency"; timeout 600 python scripts/gpu_c1_latency.py ${TAG} > gpurun_out/${TAG}_c1_latency.log 2>&1; echo "exit $?"; tail -150 gpurun_out/${TAG}_c1_latency.log
for N in 512 4096; do for SER in 1 0; do
echo "== sweep N=$N serial=$SER"; SWEEP_N=$N SWEEP_SERIAL=$SER timeout 600 python scripts/gpu_sweep.py ${TAG}_N${N}_ser${SER} 2>&1 | tail -1
done; done
