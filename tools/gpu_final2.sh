#!/bin/bash
# Final 1-GPU session of the round: smoke, the whole GPU test-suite, LU-fallback timing, CUPTI launch list, bench
set -u
TAG=${1:-r2y}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
echo "== smoke"; date
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== pytest -m gpu"; date
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log; grep -E "FAILED|Error" $OUT/${TAG}_pytest.log | head -10
echo "== forced LU fallback timing"; date
timeout 300 python tools/lu_timing.py > $OUT/${TAG}_lu_timing.json 2> $OUT/${TAG}_lu_timing.err; echo "rc=$?"; cat $OUT/${TAG}_lu_timing.json; tail -2 $OUT/${TAG}_lu_timing.err
echo "== CUPTI launch list of one solve"; date
timeout 300 python tools/cupti_launches.py > $OUT/${TAG}_cupti_launches.txt 2> $OUT/${TAG}_cupti.err; echo "rc=$?"; head -14 $OUT/${TAG}_cupti_launches.txt; tail -2 $OUT/${TAG}_cupti.err
echo "== bench"; date
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; head -c 300 $OUT/${TAG}_bench.json; echo
date
