#!/bin/bash
# r5c: QR4 (four lanes per column in the warm-start QR) parity + sequence throughput; then the --set full capture of the headline kernel
TAG=r5c; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sequence.py tests/test_gpu_parity.py -m gpu -q -x -k "warm or sequence or experimental" > $OUT/${TAG}_pytest_warm.txt 2>&1; tail -4 $OUT/${TAG}_pytest_warm.txt
timeout 200 python bench.py --sequence --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_seq.json 2> $OUT/${TAG}_seq.err; cut -c1-330 $OUT/${TAG}_seq.json
bash scripts/gpu_r5b.sh
