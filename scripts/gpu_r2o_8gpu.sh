#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python scripts/multi_gpu_e2e.py --gpus 4,8 --steps 4 --json $OUT/r2o_multi_e2e.json > $OUT/r2o_multi_e2e.txt 2>&1; tail -6 $OUT/r2o_multi_e2e.txt
