#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for t in 32 64 128 256; do
  export JRLQP_BLOCKGI_THREADS=$t
  timeout 600 python scripts/bench_blockgi.py --batch 8192 --base 256 --steps 2 --warmup 1 --cpu-sample 512 --dense-sample 256 > $OUT/r02a_blockgi_E_t$t.json 2> $OUT/r02a_blockgi_E_t$t.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r02a_blockgi_E_t$t.json").read().strip().splitlines()[-1]); print("threads $t", round(d["value"]), d["config"]["kernel"]["ctas_per_sm"], d["verified"])
except Exception as e: print("threads $t failed", e); print(open("$OUT/r02a_blockgi_E_t$t.err").read()[-500:])
PY
done
