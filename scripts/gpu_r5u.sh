#!/bin/bash
# r5u: last check of the tree: the whole GPU suite, the default bench line, config C / C2 warm at the BASELINE batch
TAG=r5u; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -2 $OUT/${TAG}_pytest.log
timeout 300 python bench.py > $OUT/${TAG}_bench_A_default.json 2> $OUT/${TAG}_bench_A.err; tail -1 $OUT/${TAG}_bench_A_default.json | cut -c1-160
for c in C C2; do
  timeout 300 python bench.py --config $c --warm --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_${c}_warm.json 2> $OUT/${TAG}_bench_${c}_warm.err; tail -1 $OUT/${TAG}_bench_${c}_warm.json | cut -c1-160
done
