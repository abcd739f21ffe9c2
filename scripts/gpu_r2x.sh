#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_blockgi.py -x -q 2>&1 | tail -5 | tee $OUT/r2x_pytest_blockgi.txt
for t in 128 256 64; do JRLQP_BLOCKGI_THREADS=$t timeout 400 python scripts/bench_blockgi.py --batch 32768 --steps 2 --warmup 2 --cpu-sample 512 --dense-sample 512 > $OUT/r2x_blockgi_t$t.json 2> $OUT/r2x_err_$t.txt; python - <<PY
import json
try:
    d=json.loads(open('$OUT/r2x_blockgi_t$t.json').read().strip().split('\n')[-1])
    print('threads $t: %.1f k QP/s  %s  cpu %.0f'%(d['value']/1e3, d['verified'], d['cpu_baseline']['value']))
except Exception as e:
    print('threads $t failed', e); print(open('$OUT/r2x_err_$t.txt').read()[-600:])
PY
done
