#!/bin/bash
# GPU session 2 (1 GPU): new parity tests, bench regression check, MD / B3k workloads and the MD launch profile.
set -u
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest (headline + dist world=1)"; date
timeout 1500 python -m pytest tests/test_gpu_headline.py tests/test_dist_gpu.py -m gpu -q -s --timeout 1200 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
echo "== bench s20k"; date
timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo rc=$?
echo "== bench md"; date
timeout 600 python bench.py --workload md --steps 2 --warmup 3 > $OUT/${TAG}_bench_md.json 2> $OUT/${TAG}_bench_md.err; echo rc=$?
echo "== bench b3k"; date
timeout 300 python bench.py --workload b3k --steps 10 --warmup 3 > $OUT/${TAG}_bench_b3k.json 2> $OUT/${TAG}_bench_b3k.err; echo rc=$?
echo "== ncu launch list: md 1024 frames"; date
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches_md.csv \
    python bench.py --workload md --frames 1024 --steps 1 --warmup 1 > $OUT/${TAG}_ncu_md.log 2>&1; echo rc=$?
echo "== matrix-free timing: S20k forced iterative"; date
timeout 900 python tools/matfree_run.py 6667 > $OUT/${TAG}_matfree_s20k.txt 2>&1; echo rc=$?
date
