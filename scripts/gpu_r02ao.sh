#!/bin/bash
set -u
N=${1:-2}; TAG=${2:-r02ao}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for KG in 5 1; do
echo "== bench N=$N gather-every $KG"; timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --gather-every $KG --no-ref-ext > gpurun_out/${TAG}_bench_n${N}_g${KG}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "exit $?"; tail -3 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n${N}_g${KG}.json"))
print("value",d["value"],"ms",d["ms_per_step"],"launches",d["gpu_launches"],"exchange",d["exchange"],"e2e",d["e2e"]["value"])
PY
done
if [ "$N" == "2" ]; then
echo "== 1-GPU sweep"; CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/gpu_sweep.py ${TAG} 2>&1 | tail -1
fi
