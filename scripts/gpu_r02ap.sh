#!/bin/bash
set -u
N=${1:-8}; TAG=${2:-r02zz}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "exit $?"; tail -2 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print("value",d["value"],"ms",d["ms_per_step"],"launches",d["gpu_launches"],"exchange",d["exchange"]["steps_per_all_gather"],"e2e",d["e2e"]["value"],d["e2e"]["two_steps_in_flight"].get("value"),"loop",d["loop"]["hyp_iter_per_s"])
PY
