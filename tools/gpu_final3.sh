#!/bin/bash
# row-maximum kernel with more loads in flight (probe + bench), then the ncu launch list of the first 1500 launches of a
# steady-state solve (B200_PROFILING.md recipe)
set -u
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
echo "== probe"; date
timeout 200 tools/ozaki_probe > $OUT/${TAG}_probe.json 2>&1; echo "rc=$?"; grep -E "ms_total|failed" $OUT/${TAG}_probe.json | cut -c1-60,230-
echo "== bench"; date
timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; rc=$?; echo "rc=$rc"
python - <<PY
import json
for l in open("$OUT/${TAG}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"]["dq"], d["parity"]["ok"])
PY
echo "== ncu launch list (launches 3850..5350 = the start of the second solve)"; date
[ $rc -eq 0 ] && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3850 -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1; echo "rc=$?"
date
