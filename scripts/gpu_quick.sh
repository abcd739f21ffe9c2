#!/bin/bash
# Quick GPU visit: parity tests + device-resident bench lines for configs A/B/D (no CPU baseline, no e2e).
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "A 131072" "B 1048576" "D 16384"; do
  set -- $cfg
  timeout 300 python bench.py --config $1 --batch $2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['verified'], d['config']['kernel']['regs_per_thread'], d['config']['kernel']['qps_per_sm'])
except Exception as e: print('$1 failed', e)"
done
