#!/bin/bash
# re-entry check of HEAD on a fresh box + C1 latency evidence (ncu launch list and full capture of one frame)
set -u
mkdir -p gpurun_out
TAG=${1:-r02aa}
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/${TAG}_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
echo "== bench ours"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "exit $?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== C1 ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sdfr_|emset" -c 60 --csv --log-file gpurun_out/${TAG}_c1_launches.csv \
   python scripts/gpu_c1_probe.py 8 > gpurun_out/${TAG}_c1_probe.log 2>&1; echo "exit $?"
echo "== C1 ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sdfr_forward|sdfr_backward" -s 8 -c 4 -f -o gpurun_out/${TAG}_c1_prof \
   python scripts/gpu_c1_probe.py 8 > gpurun_out/${TAG}_c1_full.log 2>&1; echo "exit $?"
ls -la gpurun_out | tail -8
