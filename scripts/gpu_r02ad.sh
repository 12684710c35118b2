#!/bin/bash
# FFMA2 convolution A/B: decoder parity tests on the default build, per-kernel times of both builds, the loop
set -u
mkdir -p gpurun_out
TAG=${1:-r02ad}
echo "== decoder tests"; timeout 900 python -m pytest tests/test_decoder_trunk.py tests/test_decoder_tail.py tests/test_hypothesis_step.py -q -m gpu -x 2>&1 | tail -3
echo "== ops default (FFMA2)"; timeout 600 python scripts/gpu_decoder_ops.py ${TAG}_ffma2 2>&1 | grep conv
echo "== ops scalar"; SDFR_LIB_PATH=$PWD/build/variants/conv_scalar.so timeout 600 python scripts/gpu_decoder_ops.py ${TAG}_scalar 2>&1 | grep conv
echo "== ops default again"; timeout 600 python scripts/gpu_decoder_ops.py ${TAG}_ffma2b 2>&1 | grep conv
echo "== bench"; timeout 900 python bench.py --no-ref-ext --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['loop']['ms_per_iteration'], d['e2e']['ms_per_step'])"
