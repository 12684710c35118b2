#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (ours + reference arm), micro measurements, ncu
# launch list and a full ncu capture of the hot kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
if [ "${SKIP_REF:-0}" != "1" ]; then
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
fi
echo "== bench ours"; timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== micro"; timeout 900 python scripts/gpu_micro.py ${TAG} > gpurun_out/${TAG}_micro.log 2>&1; echo "exit $?"; tail -100 gpurun_out/${TAG}_micro.log
echo "== configs 3 and 5"; timeout 300 python scripts/gpu_configs.py ${TAG} > gpurun_out/${TAG}_configs.log 2>&1; echo "exit $?"; tail -5 gpurun_out/${TAG}_configs.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sdfr_|emset" -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch_bench.log 2>&1; echo "exit $?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdfr_ -s 8 -c 4 -f -o gpurun_out/${TAG}_prof \
   python bench.py --steps 2 --warmup 1 --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_bench.log 2>&1; echo "exit $?"
ls -la gpurun_out | tail -20
