#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02am}
echo "== bench"; timeout 900 python bench.py --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['step'], d['gpu_launches'], d['e2e']['ms_per_step'], d['loop']['ms_per_iteration'])"
