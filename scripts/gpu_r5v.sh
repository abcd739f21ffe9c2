#!/bin/bash
# r5v: d = J^T n+ of the large kernel with 4 instead of 2 iterations in flight (config C cold, 32 k QPs)
TAG=r5v; OUT=gpurun_out; mkdir -p $OUT
for v in main du4; do
  if [ $v = main ]; then L=jrl-qp_b200/_build/libjrlqp_b200.so; else L=jrl-qp_b200/_build/libjrlqp_b200_$v.so; fi
  JRLQP_B200_LIB=$PWD/$L timeout 200 python bench.py --config C --batch 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C cold $v', round(d['value']), d['verified']['all_success'], d['verified'].get('oracle_bit_exact_sample'))" | tee -a $OUT/${TAG}_ab.txt
done
