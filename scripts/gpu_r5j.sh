#!/bin/bash
TAG=r5j; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "ring" > $OUT/${TAG}_pytest_ring.txt 2>&1; tail -5 $OUT/${TAG}_pytest_ring.txt
timeout 400 python bench.py --config C --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_C_cold.json 2> $OUT/${TAG}_C.err; cut -c1-250 $OUT/${TAG}_bench_C_cold.json
