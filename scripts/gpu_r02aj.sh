#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02aj}
echo "== pytest bench parity"; timeout 1200 python -m pytest tests/test_bench_parity.py -q -m gpu -x 2>&1 | tail -15
echo "== bench"; timeout 900 python bench.py --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d["step"], d["roofline"]["kernels"]["fused_normalized_incl_memsets"], d['roofline']['kernels']['fused_incl_memsets']['ms'], d['gpu_launches'])"
echo "== bench nograph"; timeout 900 python bench.py --no-ref-ext --no-cpu-baseline --no-step-graph > gpurun_out/${TAG}_bench_nograph.json 2>gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_nograph.json')); print(d['value'], d['ms_per_step'])"
