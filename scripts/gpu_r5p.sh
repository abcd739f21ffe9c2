#!/bin/bash
# r5p: J = J Q of the warm large-n kernel on shared-memory slabs: parity tests, config C / C2 warm with the slabs on and off
TAG=r5p; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x > $OUT/${TAG}_pytest_large.txt 2>&1; tail -3 $OUT/${TAG}_pytest_large.txt
for c in C C2; do for sl in 1 0; do
  JRLQP_LARGE_SLAB=$sl timeout 300 python bench.py --config $c --warm --batch 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${c}_warm_slab$sl.json 2> $OUT/${TAG}_${c}_warm_slab$sl.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_${c}_warm_slab$sl.json").read().strip().splitlines()[-1])
    print("$c warm slab=$sl", round(d["value"]), "QP/s", d["verified"]["all_success"], d["verified"].get("oracle_bit_exact_sample"))
except Exception as e:
    print("$c slab=$sl FAILED", e)
PY
done; done
