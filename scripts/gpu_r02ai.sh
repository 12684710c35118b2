#!/bin/bash
set -u
mkdir -p gpurun_out
bash scripts/ab/ab_variants.sh ${1:-r02ai}
