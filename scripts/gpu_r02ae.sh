#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02ae}
echo "== ops default (FFMA2)"; timeout 600 python scripts/gpu_decoder_ops.py ${TAG}_default 2>&1 | grep conv
for v in build/variants/*.so; do n=$(basename $v .so); echo "== ops $n"; SDFR_LIB_PATH=$PWD/$v timeout 600 python scripts/gpu_decoder_ops.py ${TAG}_$n 2>&1 | grep conv; done
