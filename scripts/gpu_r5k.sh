#!/bin/bash
# r5k: L2 prefetch down the columns in d = J^T n+ of the large kernel (A/B on config C cold), then an ncu capture of the kernel with the TMA ring
TAG=r5k; OUT=gpurun_out; mkdir -p $OUT
for v in main dpf0 dpf256; do
  if [ $v = main ]; then L=jrl-qp_b200/_build/libjrlqp_b200.so; else L=jrl-qp_b200/_build/libjrlqp_b200_$v.so; fi
  for c in C C2; do
  JRLQP_B200_LIB=$PWD/$L timeout 300 python bench.py --config $c --batch 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_${c}_$v.json 2> $OUT/${TAG}_${c}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_${c}_$v.json").read().strip().splitlines()[-1])
    print("$c $v", round(d["value"]), "QP/s", d["verified"]["all_success"], d["verified"].get("oracle_bit_exact_sample"))
except Exception as e:
    print("$c $v FAILED", e)
PY
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_large_kernel -s 1 -c 1 -f -o $OUT/${TAG}_prof_C python bench.py --config C --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-200
