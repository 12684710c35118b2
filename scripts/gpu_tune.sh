#!/bin/bash
# run scripts/gpu_micro.py against every prebuilt variant under build/variants/
mkdir -p gpurun_out
for so in build/variants/*.so; do
  tag=tune_$(basename $so .so)
  echo "== $so"
  SDFR_LIB_PATH=$PWD/$so timeout 600 python scripts/gpu_micro.py $tag > gpurun_out/${tag}.log 2>&1; echo "exit $?"
done
