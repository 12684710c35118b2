#!/bin/bash
# Quick GPU iteration: parity tests, bench (ours), optional micro, ncu launch list + full capture of
# the hot kernels.  usage: gpu_iter.sh TAG [micro] [noncu]
set -u
mkdir -p gpurun_out
TAG=${1:-it}
shift || true
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/${TAG}_pytest.log
echo "== bench ours"; timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [[ " $* " == *" micro "* ]]; then
echo "== micro"; timeout 900 python scripts/gpu_micro.py ${TAG} > gpurun_out/${TAG}_micro.log 2>&1; echo "exit $?"; tail -60 gpurun_out/${TAG}_micro.log
fi
if [[ " $* " != *" noncu "* ]]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sdfr_|emset" -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch_bench.log 2>&1; echo "exit $?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdfr_ -s 8 -c 4 -f -o gpurun_out/${TAG}_prof \
   python bench.py --steps 2 --warmup 1 --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_bench.log 2>&1; echo "exit $?"
fi
ls -la gpurun_out | tail -8
