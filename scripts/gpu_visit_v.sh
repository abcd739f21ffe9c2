#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/r01v_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r01v_bench_A.json 2> $OUT/r01v_bench_A.err; tail -1 $OUT/r01v_bench_A.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('A', round(d['value']), 'e2e', d['e2e'], 'cpu', d['cpu_baseline']['value'])"
