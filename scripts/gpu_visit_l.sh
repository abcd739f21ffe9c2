#!/bin/bash
# r01l visit: CH sweep (values in flight in the constraint scan) + pipelined back substitution; ncu full capture of A.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $OUT/r01l_pytest.log; cat $OUT/r01l_pytest.log
B=jrl-qp_b200/_build
for lib in base ch28 ch36 ch52; do
  if [ $lib = base ]; then unset JRLQP_B200_LIB; else export JRLQP_B200_LIB=$PWD/$B/libjrlqp_b200_$lib.so; fi
  for cfg in "A 131072" "B 1048576" "D 16384"; do
    set -- $cfg
    timeout 300 python bench.py --config $1 --batch $2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r01l_${lib}_$1.json 2> $OUT/r01l_${lib}_$1.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/r01l_${lib}_$1.json").read().strip().splitlines()[-1]); print("$lib $1", round(d["value"]), d["verified"], d["config"]["kernel"]["regs_per_thread"])
except Exception as e: print("$lib $1 failed", e)
PY
  done
done
unset JRLQP_B200_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_dense -s 1 -c 1 -f -o $OUT/r01l_prof \
  python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-e2e > $OUT/r01l_ncu_full.log 2>&1
ls -la $OUT | tail -20
