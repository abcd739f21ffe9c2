#!/bin/bash
# r5e: register cap of the one-warp WARM kernel (n <= 32): sequences of n = 20 with the uncapped (main), 168- and 128-register builds
TAG=r5e; OUT=gpurun_out; mkdir -p $OUT
for v in main wm12 wm16; do
  if [ $v = main ]; then L=jrl-qp_b200/_build/libjrlqp_b200.so; else L=jrl-qp_b200/_build/libjrlqp_b200_$v.so; fi
  JRLQP_B200_LIB=$PWD/$L timeout 300 python bench.py --sequence --seq-n 20 --batch 65536 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/${TAG}_seq20_$v.json 2> $OUT/${TAG}_seq20_$v.err
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_seq20_$v.json"))
print("$v", round(d["value"]), "QP-steps/s warm;", round(d["cold"]["qp_steps_per_s"]), "cold;", d["config"]["kernel"], d["warm"]["iterations_per_sequence"], d["max_abs_dx_warm_vs_cold_last_step"])
PY
done
