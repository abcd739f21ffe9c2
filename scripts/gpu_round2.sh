#!/bin/bash
# Full validation visit: parity tests, bench lines (A with CPU baseline and e2e, B, D), reference arm, ncu launch list,
# one --set full capture of the headline kernel, phase timing. Usage (under gpurun): bash scripts/gpu_round2.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_A.json 2> $OUT/${TAG}_bench_A.err; tail -1 $OUT/${TAG}_bench_A.json | cut -c1-400
timeout 300 python bench.py --config B --batch 1048576 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_B.json 2> $OUT/${TAG}_bench_B.err; tail -1 $OUT/${TAG}_bench_B.json | cut -c1-200
timeout 300 python bench.py --config D --batch 16384 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_D.json 2> $OUT/${TAG}_bench_D.err; tail -1 $OUT/${TAG}_bench_D.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_bench_A_reference_arm.json 2> $OUT/${TAG}_bench_ref.err; tail -1 $OUT/${TAG}_bench_A_reference_arm.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --batch 32768 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gi_dense -s 1 -c 1 -f -o $OUT/${TAG}_prof \
  python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full.log 2>&1
timeout 300 python scripts/phase_timing.py --config A --batch 32768 > $OUT/${TAG}_phase_A.txt 2>&1
timeout 300 python scripts/phase_timing.py --config D --batch 4096 > $OUT/${TAG}_phase_D.txt 2>&1
ls -la $OUT | tail -15
