#!/bin/bash
# Config C (MultiIK fixtures) on the global-workspace kernel: cold and warm bench lines.
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
for cfg in "C 16384" "C2 65536"; do
  set -- $cfg
  timeout 900 python bench.py --config $1 --batch $2 --steps 3 --warmup 3 > $OUT/${TAG}_bench_$1_cold.json 2> $OUT/${TAG}_bench_$1_cold.err; tail -c 1500 $OUT/${TAG}_bench_$1_cold.json; tail -3 $OUT/${TAG}_bench_$1_cold.err
  timeout 900 python bench.py --config $1 --batch $2 --steps 3 --warmup 3 --warm --no-cpu-baseline > $OUT/${TAG}_bench_$1_warm.json 2> $OUT/${TAG}_bench_$1_warm.err; tail -c 1500 $OUT/${TAG}_bench_$1_warm.json; tail -3 $OUT/${TAG}_bench_$1_warm.err
done
