#!/bin/bash
# same-box A/B: the default build and every variant under build/variants/ on the C2 kernels
# (build a variant with: SDFR_NVCC_EXTRA="-DFLAG" SDFR_BUILD_OUT=$PWD/build/variants/name.so python -m sdfest_b200.build --force)
TAG=${1:-ab}
mkdir -p gpurun_out
{
python scripts/ab/ab_c2.py
for so in build/variants/*.so; do SDFR_LIB_PATH=$PWD/$so python scripts/ab/ab_c2.py; done
python scripts/ab/ab_c2.py
} 2>&1 | grep '^{' | tee gpurun_out/${TAG}_variants.jsonl
