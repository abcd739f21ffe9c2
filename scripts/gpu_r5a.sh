#!/bin/bash
# r5a: factor cache of warm sequences (tests + A/B), e2e chunk sweep
TAG=r5a; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_sequence.py -m gpu -q -x > $OUT/${TAG}_pytest_sequence.txt 2>&1; tail -5 $OUT/${TAG}_pytest_sequence.txt
JRLQP_SEQ_FCACHE=0 timeout 200 python bench.py --sequence --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_seq_fcache0.json 2> $OUT/${TAG}_seq0.err; cut -c1-400 $OUT/${TAG}_seq_fcache0.json
JRLQP_SEQ_FCACHE=1 timeout 200 python bench.py --sequence --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_seq_fcache1.json 2> $OUT/${TAG}_seq1.err; cut -c1-400 $OUT/${TAG}_seq_fcache1.json
timeout 300 python scripts/e2e_chunk_sweep.py > $OUT/${TAG}_e2e_chunk_sweep.txt 2>&1; cat $OUT/${TAG}_e2e_chunk_sweep.txt
