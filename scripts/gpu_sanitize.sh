#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (memcheck, then racecheck on the render
# kernels' shared-memory work lists).  Slow (10-50x): sizes are the tests' smallest.
set -u
mkdir -p gpurun_out
TAG=${1:-san}
CS="compute-sanitizer --error-exitcode 9 --print-limit 5"
echo "== memcheck: smoke"; timeout 900 $CS --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/${TAG}_memcheck_smoke.log
echo "== memcheck: views / constraint / bounds / tail / pipeline reuse"
timeout 2400 $CS --tool memcheck python -m pytest -q -m gpu -x \
  tests/test_views.py tests/test_pipeline_gpu.py \
  "tests/test_decoder_tail.py::test_tail_forward_with_bounds_equals_tail_then_scan" \
  "tests/test_hypothesis_step.py::test_step_kernel_matches_oracle" \
  "tests/test_hypothesis_step.py::test_fused_optimizer_matches_torch_optimizer" \
  > gpurun_out/${TAG}_memcheck_tests.log 2>&1; echo "exit $?"; tail -6 gpurun_out/${TAG}_memcheck_tests.log
echo "== racecheck: smoke (shared-memory work list, moment accumulators)"; timeout 900 $CS --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck_smoke.log 2>&1; echo "exit $?"; tail -4 gpurun_out/${TAG}_racecheck_smoke.log
