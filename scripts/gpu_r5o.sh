#!/bin/bash
# r5o: ncu capture of the WARM large-n kernel on config C (MultiIK n = 387, warm start from a neighbour's active set)
TAG=r5o; OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_large_kernel -s 3 -c 1 -f -o $OUT/${TAG}_prof_Cwarm python bench.py --config C --warm --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
