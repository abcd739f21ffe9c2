#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for mode in zerocopy lower; do
  unset JRLQP_H2D_FULL_G JRLQP_G_ZEROCOPY
  if [ $mode = zerocopy ]; then export JRLQP_G_ZEROCOPY=1; fi
  for cfg in "A 131072" "B 1048576" "D 16384"; do
    set -- $cfg
    timeout 400 python bench.py --config $1 --batch $2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r01u_e2e_${mode}_$1.json 2> $OUT/r01u_e2e_${mode}_$1.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/r01u_e2e_${mode}_$1.json").read().strip().splitlines()[-1]); print("$mode $1", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["verified"]["oracle_bit_exact_sample"])
except Exception as e: print("$mode $1 failed", e); print(open("$OUT/r01u_e2e_${mode}_$1.err").read()[-600:])
PY
  done
done
