#!/bin/bash
# Usage (under gpurun --gpus N): bash scripts/gpu_scale.sh <tag> <N>
TAG=$1; N=$2
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_A_n$N.json 2> $OUT/${TAG}_bench_A_n$N.err
tail -1 $OUT/${TAG}_bench_A_n$N.json | cut -c1-1500
tail -3 $OUT/${TAG}_bench_A_n$N.err
