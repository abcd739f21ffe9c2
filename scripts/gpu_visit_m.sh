#!/bin/bash
# r01m visit: transposed-C constraint scan vs in-place scan.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $OUT/r01m_pytest.log; cat $OUT/r01m_pytest.log
for mode in t i; do
  if [ $mode = i ]; then FLAG=--scan-inplace; else FLAG=; fi
  for cfg in "A 131072" "B 1048576" "D 16384"; do
    set -- $cfg
    for st in -1 0; do
      if [ $1 != B ] && [ $st = 0 ]; then continue; fi
      timeout 300 python bench.py --config $1 --batch $2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --stage-c $st $FLAG > $OUT/r01m_${mode}_$1_$st.json 2> $OUT/r01m_${mode}_$1_$st.err
      python - <<PY
import json
try:
    d=json.loads(open("$OUT/r01m_${mode}_$1_$st.json").read().strip().splitlines()[-1]); print("$mode $1 stage $st", round(d["value"]), d["verified"]["all_success"], d["verified"]["oracle_bit_exact_sample"], d["config"]["kernel"])
except Exception as e: print("$mode $1 failed", e); print(open("$OUT/r01m_${mode}_$1_$st.err").read()[-800:])
PY
    done
  done
done
timeout 300 python scripts/phase_timing.py --config A --batch 32768 > $OUT/r01m_phase_A.txt 2>&1; head -16 $OUT/r01m_phase_A.txt
timeout 300 python scripts/phase_timing.py --config D --batch 4096 > $OUT/r01m_phase_D.txt 2>&1; head -16 $OUT/r01m_phase_D.txt
