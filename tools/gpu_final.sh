#!/bin/bash
# Final 1-GPU session of the round: smoke, the whole GPU test-suite, both bench arms as the driver runs them, the
# secondary workloads, a fresh ncu launch list.
set -u
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== smoke"; date
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 1500 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
echo "== bench --impl reference"; date
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "rc=$?"; head -c 300 $OUT/${TAG}_bench_reference.json; echo
echo "== bench (driver flags)"; date
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; head -c 300 $OUT/${TAG}_bench.json; echo
echo "== bench md / b3k"; date
timeout 600 python bench.py --workload md --steps 2 --warmup 3 > $OUT/${TAG}_bench_md.json 2> $OUT/${TAG}_bench_md.err; echo "rc=$?"
timeout 300 python bench.py --workload b3k --steps 10 --warmup 3 > $OUT/${TAG}_bench_b3k.json 2> $OUT/${TAG}_bench_b3k.err; echo "rc=$?"
echo "== ncu launch list"; date
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full: tile_matvec (tile-major) + rect_matvec"; date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_matvec_kernel|rect_matvec_t_kernel" -s 30 -c 4 \
    -o $OUT/${TAG}_ncu_matvec python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/${TAG}_ncu_matvec.log 2>&1; echo "rc=$?"
timeout 60 tools/tc05_probe > $OUT/${TAG}_tc05_probe.json 2>&1; cat $OUT/${TAG}_tc05_probe.json
date
