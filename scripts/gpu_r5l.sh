#!/bin/bash
# r5l: ring with 4 stages, L2 evict_last policy on the scan's loads of the transposed C, evict_first on the ring copies (config C cold)
TAG=r5l; OUT=gpurun_out; mkdir -p $OUT
for v in main rs4 evl evlf; do
  if [ $v = main ]; then L=jrl-qp_b200/_build/libjrlqp_b200.so; else L=jrl-qp_b200/_build/libjrlqp_b200_$v.so; fi
  JRLQP_B200_LIB=$PWD/$L timeout 300 python bench.py --config C --batch 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_C_$v.json 2> $OUT/${TAG}_C_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_C_$v.json").read().strip().splitlines()[-1])
    print("C $v", round(d["value"]), "QP/s smem", d["config"]["kernel"]["smem_bytes_per_qp"], d["verified"]["all_success"], d["verified"].get("oracle_bit_exact_sample"))
except Exception as e:
    print("C $v FAILED", e)
PY
done
for v in main evl; do
  if [ $v = main ]; then L=jrl-qp_b200/_build/libjrlqp_b200.so; else L=jrl-qp_b200/_build/libjrlqp_b200_$v.so; fi
  JRLQP_B200_LIB=$PWD/$L timeout 300 python bench.py --config D --batch 16384 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('D $v', round(d['value']))"
done
