#!/bin/bash
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
bash scripts/gpu_visit_ab.sh $TAG "$@"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_sequence.py -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
