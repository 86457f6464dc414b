#!/bin/bash
# Ozaki path as the default: smoke, whole GPU test-suite, bench, ncu capture of oz_gemm_kernel
set -u
TAG=${1:-oz4}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
echo "== smoke"; date
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_smoke.log
echo "== pytest -m gpu"; date
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log; grep -E "^\[S20k|FAILED|Error" $OUT/${TAG}_pytest.log | head -20
echo "== bench"; date
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; head -c 400 $OUT/${TAG}_bench.json; echo
echo "== ncu full: oz_gemm_kernel"; date
timeout 300 python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/${TAG}_plain.json 2>/dev/null && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -c 16 -o $OUT/${TAG}_ncu_ozgemm \
    python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1; echo "rc=$?"
date
