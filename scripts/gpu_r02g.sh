#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02g}
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "exit $?"; tail -8 gpurun_out/${TAG}_pytest.log | cut -c1-300; grep -n "AssertionError\|^E  " gpurun_out/${TAG}_pytest.log | cut -c1-250 | head
echo "== bench ours"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"])
for n,v in d["roofline"]["kernels"].items(): print(n, round(v["ms"]*1e3,1),"us")
print(d.get("loop"))
PY
tail -5 gpurun_out/${TAG}_bench.err
echo "== micro"; timeout 600 python scripts/gpu_micro.py ${TAG} > gpurun_out/${TAG}_micro.log 2>&1; echo "exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_micro.json"))
for k,v in d.items():
    if "median_us" in v: print(k, round(v["median_us"],1))
    else:
        for kk,vv in v.items():
            if "median_us" in vv: print(" ",k,kk, round(vv["median_us"],1))
            else: print(" ",k,kk,{a:round(b["median_us"],1) for a,b in vv.items()})
PY
