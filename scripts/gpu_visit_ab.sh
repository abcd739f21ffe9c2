#!/bin/bash
# Usage: bash scripts/gpu_visit_ab.sh <tag> <names...>
OUT=gpurun_out
mkdir -p $OUT
TAG=$1; shift
timeout 600 python scripts/ab_variants.py --config A --batch 65536 --rounds 3 "$@" 2>&1 | tail -12 | tee $OUT/${TAG}_ab_A.txt
timeout 600 python scripts/ab_variants.py --config D --batch 8192 --rounds 2 "$@" 2>&1 | tail -12 | tee $OUT/${TAG}_ab_D.txt
timeout 600 python scripts/ab_variants.py --config B --batch 524288 --rounds 2 "$@" 2>&1 | tail -12 | tee $OUT/${TAG}_ab_B.txt
