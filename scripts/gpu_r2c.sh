#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2c}
timeout 300 python scripts/quick_parity.py > $OUT/${TAG}_quick.txt 2>&1; grep -c "bit-exact True" $OUT/${TAG}_quick.txt; grep "mismatching instances [1-9]\|bit-exact False\|Error\|error" $OUT/${TAG}_quick.txt | head
bash scripts/gpu_visit_ab.sh $TAG base main
for c in A B D; do timeout 200 python scripts/phase_timing.py --no-build --config $c --batch 16384 > $OUT/${TAG}_phase_$c.txt 2>&1; done
cat $OUT/${TAG}_phase_A.txt
