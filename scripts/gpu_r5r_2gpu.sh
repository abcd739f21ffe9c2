#!/bin/bash
# r5r (2 GPUs): the multi-GPU library entry on two devices + the bench under torchrun
TAG=r5r; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/${TAG}_pytest_multi_2gpu.txt 2>&1; tail -2 $OUT/${TAG}_pytest_multi_2gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_A_n2.json 2> $OUT/${TAG}_bench_A_n2.err; tail -1 $OUT/${TAG}_bench_A_n2.json | cut -c1-400
