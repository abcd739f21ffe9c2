#!/bin/bash
# r5f: register cap of the two-warp WARM kernel: sequences of n = 50 and n = 40, uncapped (main: 221 registers) vs 160 registers
TAG=r5f; OUT=gpurun_out; mkdir -p $OUT
for n in 50 40; do
for v in main w2m6; do
  if [ $v = main ]; then L=jrl-qp_b200/_build/libjrlqp_b200.so; else L=jrl-qp_b200/_build/libjrlqp_b200_$v.so; fi
  JRLQP_B200_LIB=$PWD/$L timeout 300 python bench.py --sequence --seq-n $n --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_seq${n}_$v.json 2> $OUT/${TAG}_seq${n}_$v.err
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_seq${n}_$v.json"))
print("n=$n $v", round(d["value"]), "QP-steps/s warm;", round(d["cold"]["qp_steps_per_s"]), "cold;", d["warm"]["iterations_per_sequence"], d["max_abs_dx_warm_vs_cold_last_step"])
PY
done; done
timeout 300 python -m pytest tests/test_gpu_sequence.py -m gpu -q -x 2>&1 | tail -2
