#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $OUT/r01zc_smoke.txt
for t in tri down up; do
  timeout 600 python scripts/bench_structured.py --type $t > $OUT/r01zc_structured_$t.json 2> $OUT/r01zc_structured_$t.err; tail -1 $OUT/r01zc_structured_$t.json | cut -c1-300
done
timeout 900 python scripts/bench_blockgi.py --batch 65536 --base 256 --steps 2 --warmup 1 > $OUT/r01zc_blockgi_E_tri.json 2> $OUT/r01zc_blockgi_E_tri.err; tail -1 $OUT/r01zc_blockgi_E_tri.json | cut -c1-300
timeout 900 python scripts/bench_blockgi.py --type up --batch 32768 --base 256 --steps 2 --warmup 1 > $OUT/r01zc_blockgi_E_up.json 2> $OUT/r01zc_blockgi_E_up.err; tail -1 $OUT/r01zc_blockgi_E_up.json | cut -c1-300
