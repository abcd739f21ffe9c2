#!/bin/bash
# r5b: one --set full capture of the headline kernel (config A, 16 384 QPs) with the shipped library
TAG=r5b; OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gi_dense -s 1 -c 1 -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log | cut -c1-300
