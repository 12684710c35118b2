#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02an}
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/${TAG}_pytest.log
python scripts/ab/ab_c2.py; python scripts/ab/ab_c2.py
echo "== bench"; timeout 900 python bench.py --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['loop']['ms_per_iteration'], d['loop']['pose_only']['ms_per_iteration'])"
