#!/bin/bash
# r5t: warm_B of the large warm kernel (L entries one group ahead, proven quotients): parity tests + config C / C2 warm, new vs old
TAG=r5t; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "warm or ring" > $OUT/${TAG}_pytest_large_warm.txt 2>&1; tail -2 $OUT/${TAG}_pytest_large_warm.txt
for v in main wb0; do
  if [ $v = main ]; then L=jrl-qp_b200/_build/libjrlqp_b200.so; else L=jrl-qp_b200/_build/libjrlqp_b200_$v.so; fi
  for c in C C2; do
  JRLQP_B200_LIB=$PWD/$L timeout 300 python bench.py --config $c --warm --batch 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c warm $v', round(d['value']), d['verified']['all_success'], d['verified'].get('oracle_bit_exact_sample'))" | tee -a $OUT/${TAG}_ab.txt
  done
done
