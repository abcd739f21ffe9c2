#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_blockgi.py tests/test_orthonormal_sequence.py tests/test_gpu_structured.py -x -q 2>&1 | tail -5 | tee $OUT/r2z_pytest.txt
for t in 128 256; do JRLQP_BLOCKGI_THREADS=$t timeout 400 python scripts/bench_blockgi.py --batch 32768 --steps 2 --warmup 2 --cpu-sample 512 --dense-sample 512 > $OUT/r2z_blockgi_t$t.json 2> $OUT/r2z_err_$t.txt; python - <<PY
import json
try:
    d=json.loads(open('$OUT/r2z_blockgi_t$t.json').read().strip().split('\n')[-1])
    print('threads $t: %.1f k QP/s  %s'%(d['value']/1e3, d['verified']))
except Exception as e:
    print('threads $t failed', e); print(open('$OUT/r2z_err_$t.txt').read()[-600:])
PY
done
timeout 300 python scripts/bench_structured.py --cpu-sample 2048 > $OUT/r2z_structured_tri.json 2>/dev/null; python -c "
import json
d=json.loads(open('$OUT/r2z_structured_tri.json').read().strip().split(chr(10))[-1]); print('LLT %.2f M/s, solves %s'%(d['value']/1e6, d['solves']))"
