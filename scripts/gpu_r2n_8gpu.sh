#!/bin/bash
# 8-GPU visit: host-link ceiling + multi-GPU entry scaling (one process), then the driver-style torchrun bench at N=8
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r2n_topo.txt 2>&1
timeout 600 python scripts/multi_gpu_e2e.py --gpus 1,2,4,8 --json $OUT/r2n_multi_e2e.json > $OUT/r2n_multi_e2e.txt 2>&1; tail -8 $OUT/r2n_multi_e2e.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2n_bench_A_n8.json 2> $OUT/r2n_bench_A_n8.err; tail -1 $OUT/r2n_bench_A_n8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=8 value', round(d['value']), 'e2e', json.dumps(d['e2e'])[:900])"
tail -3 $OUT/r2n_bench_A_n8.err
