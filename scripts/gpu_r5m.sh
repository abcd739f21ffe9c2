#!/bin/bash
# r5m: shipped build (ring 3 stages for cold n >= 256, evict_last on the shared transposed C): large parity tests, config C / C2 cold and warm
TAG=r5m; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x > $OUT/${TAG}_pytest_large.txt 2>&1; tail -3 $OUT/${TAG}_pytest_large.txt
for c in C C2; do for w in "" "--warm"; do
  timeout 300 python bench.py --config $c $w --batch 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${c}${w}.json 2> $OUT/${TAG}_${c}${w}.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_${c}${w}.json").read().strip().splitlines()[-1])
    print("$c $w", round(d["value"]), "QP/s smem", d["config"]["kernel"]["smem_bytes_per_qp"], d["verified"]["all_success"], d["verified"].get("oracle_bit_exact_sample"))
except Exception as e:
    print("$c $w FAILED", e)
PY
done; done
