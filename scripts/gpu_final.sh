#!/bin/bash
# Final validation visit of a round (one GPU): parity tests, every bench configuration of BASELINE.json under the driver
# contract, the reference arm, ncu launch list, phase timing. Usage (under gpurun): bash scripts/gpu_final.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,driver_version --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_A.json 2> $OUT/${TAG}_bench_A.err; tail -1 $OUT/${TAG}_bench_A.json | cut -c1-300
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_bench_A_reference_arm.json 2> $OUT/${TAG}_bench_ref.err; tail -1 $OUT/${TAG}_bench_A_reference_arm.json | cut -c1-200
timeout 400 python bench.py --config B --steps 5 --warmup 3 > $OUT/${TAG}_bench_B.json 2> $OUT/${TAG}_bench_B.err; tail -1 $OUT/${TAG}_bench_B.json | cut -c1-200
timeout 600 python bench.py --config D --steps 3 --warmup 3 > $OUT/${TAG}_bench_D.json 2> $OUT/${TAG}_bench_D.err; tail -1 $OUT/${TAG}_bench_D.json | cut -c1-200
for c in C C2; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > $OUT/${TAG}_bench_${c}_cold.json 2> $OUT/${TAG}_bench_${c}_cold.err; tail -1 $OUT/${TAG}_bench_${c}_cold.json | cut -c1-200
  timeout 600 python bench.py --config $c --warm --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_${c}_warm.json 2> $OUT/${TAG}_bench_${c}_warm.err; tail -1 $OUT/${TAG}_bench_${c}_warm.json | cut -c1-200
done
timeout 600 python bench.py --config E --steps 3 --warmup 3 > $OUT/${TAG}_bench_E.json 2> $OUT/${TAG}_bench_E.err; tail -1 $OUT/${TAG}_bench_E.json | cut -c1-200
timeout 400 python bench.py --sequence --steps 3 --warmup 3 > $OUT/${TAG}_bench_sequence.json 2> $OUT/${TAG}_bench_seq.err; tail -1 $OUT/${TAG}_bench_sequence.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --batch 32768 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 300 python scripts/phase_timing.py --no-build --config A --batch 32768 > $OUT/${TAG}_phase_A.txt 2>&1
timeout 300 python scripts/phase_timing.py --no-build --config D --batch 4096 > $OUT/${TAG}_phase_D.txt 2>&1
ls -la $OUT | grep ${TAG} | wc -l
