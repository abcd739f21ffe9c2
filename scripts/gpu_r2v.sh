#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python bench.py --sequence --steps 3 --warmup 3 > $OUT/r2v_bench_sequence.json 2> $OUT/r2v_seq.err; tail -1 $OUT/r2v_bench_sequence.json | cut -c1-900; tail -2 $OUT/r2v_seq.err
timeout 600 python bench.py --config E --steps 3 --warmup 3 > $OUT/r2v_bench_E.json 2> $OUT/r2v_E.err; tail -1 $OUT/r2v_bench_E.json | cut -c1-600; tail -2 $OUT/r2v_E.err
timeout 400 python scripts/shape_sweep.py --json $OUT/r2v_shape_sweep.json > $OUT/r2v_shape_sweep.txt 2>&1; cat $OUT/r2v_shape_sweep.txt
