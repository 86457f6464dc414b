#!/bin/bash
set -u
TAG=${1:-oz3}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
CHEQ_GEMM_LOG=$OUT/${TAG}_gemm_log.csv CHEQ_OZAKI=7 CHEQ_OZAKI_MIN_K=512 CHEQ_OZAKI_R=5 timeout 300 python bench.py --steps 1 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"
head -c 200 $OUT/${TAG}_bench.json; echo
