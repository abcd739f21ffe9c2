#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python scripts/ab_variants.py --config A --batch 65536 --rounds 4 v0 v1 v2 v3 v4 v5 2>&1 | tail -8 | tee $OUT/r01n_ab_A.txt
timeout 600 python scripts/ab_variants.py --config D --batch 8192 --rounds 3 v0 v1 v2 v3 v4 v5 2>&1 | tail -8 | tee $OUT/r01n_ab_D.txt
timeout 600 python scripts/ab_variants.py --config B --batch 524288 --rounds 3 v0 v1 v2 v3 v4 v5 2>&1 | tail -8 | tee $OUT/r01n_ab_B.txt
