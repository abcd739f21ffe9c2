#!/bin/bash
# r5h: one --set full capture of the large-n kernel on config C (MultiIK n = 387, cold), 4096 QPs
TAG=r5h; OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_large_kernel -s 1 -c 1 -f -o $OUT/${TAG}_prof_C python bench.py --config C --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log | cut -c1-300
