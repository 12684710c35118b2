#!/bin/bash
# One multi-GPU box session (gpurun --gpus N): bench.py, the C4 sweep and C5 at N ranks.
set -u
N=${1:-8}; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${TAG}_gpus_n${N}.txt 2>&1
echo "== bench N=$N"; timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],"e2e_grids",d["e2e_grids"].get("value"))
print("loop",d["loop"]["ms_per_iteration"],d["loop"]["hyp_iter_per_s"],"pose_only",d["loop"]["pose_only"]["ms_per_iteration"])
PY
echo "== sweep (C4) N=$N"; timeout 600 $TR --master-port 29533 scripts/gpu_sweep.py ${TAG} > gpurun_out/${TAG}_sweep_n${N}.log 2>&1; echo "exit $?"; tail -3 gpurun_out/${TAG}_sweep_n${N}.log | cut -c1-600
echo "== C5 N=$N"; SKIP_C1=1 SKIP_C3=1 timeout 600 $TR --master-port 29544 scripts/gpu_configs.py ${TAG} > gpurun_out/${TAG}_configs_n${N}.log 2>&1; echo "exit $?"; tail -25 gpurun_out/${TAG}_configs_n${N}.log | cut -c1-300
