#!/bin/bash
# compute-sanitizer: memcheck over the whole GPU suite (the reference drop-in tests deselected: they JIT-load the
# reference package), memcheck + racecheck on smoke() (the benchmarked kernel instantiation).
set -u
mkdir -p gpurun_out
TAG=${1:-san}
CS="compute-sanitizer --error-exitcode 9 --print-limit 5"
echo "== memcheck: GPU suite"; timeout 1500 $CS --tool memcheck python -m pytest tests -q -m gpu --deselect tests/test_dropin_reference.py > gpurun_out/${TAG}_memcheck_suite.log 2>&1; echo "exit $?"; tail -4 gpurun_out/${TAG}_memcheck_suite.log
echo "== memcheck: smoke"; timeout 600 $CS --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/${TAG}_memcheck_smoke.log
echo "== racecheck: smoke"; timeout 600 $CS --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck_smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/${TAG}_racecheck_smoke.log
