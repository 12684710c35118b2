#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1; nproc >> gpurun_out/${TAG}_gpu.txt
echo "== bench parity tests"; timeout 1200 python -m pytest tests/test_bench_parity.py -q -m gpu -x > gpurun_out/${TAG}_pytest_parity.log 2>&1; echo "exit $?"; tail -25 gpurun_out/${TAG}_pytest_parity.log
echo "== gather peaks"; timeout 300 scripts/micro/gather_peaks > gpurun_out/${TAG}_gather_peaks.json 2>gpurun_out/${TAG}_gather_peaks.err; echo "exit $?"; cat gpurun_out/${TAG}_gather_peaks.json
echo "== C1 launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sdfr_|emset" --csv --log-file gpurun_out/${TAG}_c1_launches.csv python scripts/gpu_c1_probe.py 6 > gpurun_out/${TAG}_c1_probe.log 2>&1; echo "exit $?"; tail -30 gpurun_out/${TAG}_c1_launches.csv
echo "== bench ours"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "exit $?"; cut -c1-1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
