#!/bin/bash
# One GPU-box session: parity tests, bench (A/B: stale-factor iterations on / off), ncu launch list + full captures.
# Usage (from the repo root, through gpurun):  bash tools/gpu_run.sh <tag> [pytest-args]
set -u
TAG=${1:-run}
shift || true
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest" ; date
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 1500 "$@" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
echo "== bench (default)" ; date
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
echo "== bench (CHEQ_KRYLOV=0: refactor every iteration)" ; date
CHEQ_KRYLOV=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-secondary > $OUT/${TAG}_bench_nokrylov.json 2> $OUT/${TAG}_bench_nokrylov.err
echo "bench rc=$?"
echo "== ncu launch list" ; date
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
echo "ncu launch list rc=$?"
echo "== ncu full: tile_matvec kernels" ; date
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_matvec_kernel -s 40 -c 6 \
    -o $OUT/${TAG}_ncu_matvec python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/${TAG}_ncu_matvec.log 2>&1
echo "ncu matvec rc=$?"
date
