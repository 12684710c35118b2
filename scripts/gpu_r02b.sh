#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02b}
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "exit $?"; tail -40 gpurun_out/${TAG}_pytest.log | cut -c1-300
echo "== bench ours"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"])
print(json.dumps(d["roofline"]["kernels"],indent=0))
print(d["roofline"]["work"]); print(d.get("loop")); print(d.get("reference_cuda_ext"))
PY
tail -5 gpurun_out/${TAG}_bench.err
echo "== micro"; timeout 600 python scripts/gpu_micro.py ${TAG} > gpurun_out/${TAG}_micro.log 2>&1; echo "exit $?"; head -30 gpurun_out/${TAG}_micro.log
