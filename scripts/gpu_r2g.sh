#!/bin/bash
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
bash scripts/gpu_visit_ab.sh $TAG "$@"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gi_dense -s 1 -c 1 -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full.log 2>&1
