#!/bin/bash
# round-2 visit A: parity of the restructured dense kernel + A/B against the round-1 kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/quick_parity.py > $OUT/r2a_quick.txt 2>&1; tail -50 $OUT/r2a_quick.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee $OUT/r2a_pytest.txt
bash scripts/gpu_visit_ab.sh r2a base main
