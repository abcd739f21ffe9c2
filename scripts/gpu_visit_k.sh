#!/bin/bash
# r01k visit: phase timing (A, D), BlockGI end-to-end bench (config E), quick parity.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $OUT/r01k_pytest.log; cat $OUT/r01k_pytest.log
timeout 300 python scripts/phase_timing.py --config A --batch 32768 > $OUT/r01k_phase_A.txt 2>&1; tail -40 $OUT/r01k_phase_A.txt
timeout 300 python scripts/phase_timing.py --config D --batch 4096 > $OUT/r01k_phase_D.txt 2>&1; tail -70 $OUT/r01k_phase_D.txt
timeout 600 python scripts/bench_blockgi.py --batch 16384 --base 256 --steps 2 --warmup 1 > $OUT/r01k_blockgi_E_tri.json 2> $OUT/r01k_blockgi_E_tri.err; tail -1 $OUT/r01k_blockgi_E_tri.json; tail -3 $OUT/r01k_blockgi_E_tri.err
