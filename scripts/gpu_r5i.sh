#!/bin/bash
# r5i: TMA column ring of the large-n kernel: parity tests (n > 128), config C / C2 cold and warm with the ring on and off
TAG=r5i; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x > $OUT/${TAG}_pytest_large.txt 2>&1; tail -3 $OUT/${TAG}_pytest_large.txt
for c in C C2; do for r in 1 0; do for w in "" "--warm"; do
  JRLQP_LARGE_RING=$r timeout 300 python bench.py --config $c $w --batch 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${c}${w}_ring$r.json 2> $OUT/${TAG}_${c}${w}_ring$r.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_${c}${w}_ring$r.json").read().strip().splitlines()[-1])
    print("$c $w ring=$r", round(d["value"]), "QP/s", d["config"]["kernel"]["smem_bytes_per_qp"], d["config"]["kernel"]["qps_per_sm"], d["verified"]["all_success"], d["verified"].get("oracle_bit_exact_sample"))
except Exception as e:
    print("$c $w ring=$r FAILED", e)
PY
done; done; done
