#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02d}
echo "== debug c3"; timeout 300 python scripts/gpu_debug_c3.py 2>&1 | tail -30
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "exit $?"; tail -12 gpurun_out/${TAG}_pytest.log | cut -c1-300; grep -n "AssertionError\|^E  " gpurun_out/${TAG}_pytest.log | cut -c1-250 | head
echo "== bench ours"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"])
for n,v in d["roofline"]["kernels"].items(): print(n, round(v["ms"]*1e3,1),"us")
print(d.get("loop"))
PY
tail -5 gpurun_out/${TAG}_bench.err
