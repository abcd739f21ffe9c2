#!/bin/bash
# r5d: split-lane Cholesky columns (JRLQP_CHOL_LPR) A/B against the build without them, + cold parity tests
TAG=r5d; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python scripts/ab_variants.py --config A --batch 65536 --rounds 3 main lpr0 2>&1 | tail -8 | tee $OUT/${TAG}_ab_A.txt
timeout 600 python scripts/ab_variants.py --config B --batch 524288 --rounds 2 main lpr0 2>&1 | tail -8 | tee $OUT/${TAG}_ab_B.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sequence.py -m gpu -q -x > $OUT/${TAG}_pytest_parity.txt 2>&1; tail -3 $OUT/${TAG}_pytest_parity.txt
