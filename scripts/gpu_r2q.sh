#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_structured.py -x -q 2>&1 | tail -8 | tee $OUT/r2q_pytest_structured.txt
for k in 1 2 3; do timeout 300 python scripts/bench_structured.py --kernel $k --cpu-sample 2048 > $OUT/r2q_structured_tri_k$k.json 2> $OUT/r2q_err_$k.txt; python - <<PY
import json
try:
    d=json.loads(open('$OUT/r2q_structured_tri_k$k.json').read().strip().split('\n')[-1])
    print('kernel $k: LLT %.2f M/s  hbm frac %.3f  parity %s'%(d['value']/1e6, d['roofline']['frac'], d['verified']))
except Exception as e:
    print('kernel $k failed', e); print(open('$OUT/r2q_err_$k.txt').read()[-800:])
PY
done
