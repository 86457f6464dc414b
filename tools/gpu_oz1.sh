#!/bin/bash
# Ozaki (tcgen05 int8 slices) GEMM: stand-alone probe, then S20k through bench.py with CHEQ_OZAKI=7
set -u
TAG=${1:-oz1}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
echo "== probe quick"; date
timeout 120 tools/ozaki_probe quick > $OUT/${TAG}_probe_quick.json 2>&1; rc=$?; echo "rc=$rc"; cat $OUT/${TAG}_probe_quick.json
if [ $rc -eq 0 ]; then
  echo "== probe full"; date
  timeout 300 tools/ozaki_probe > $OUT/${TAG}_probe.json 2>&1; echo "rc=$?"; cat $OUT/${TAG}_probe.json
  echo "== bench S20k CHEQ_OZAKI=7"; date
  CHEQ_OZAKI=7 timeout 400 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench_oz7.json 2> $OUT/${TAG}_bench_oz7.err; echo "rc=$?"
  head -c 600 $OUT/${TAG}_bench_oz7.json; echo; tail -3 $OUT/${TAG}_bench_oz7.err
  python - <<PY
import json
for l in open("$OUT/${TAG}_bench_oz7.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d.get("value"), "parity", d.get("parity")); print(d["config"].get("phases_ms"))
PY
fi
echo "== bench S20k CHEQ_GEMM_VARIANT=1 (DMMA 64x128 two CTAs per SM)"; date
CHEQ_GEMM_VARIANT=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench_v1.json 2> $OUT/${TAG}_bench_v1.err; echo "rc=$?"
head -c 300 $OUT/${TAG}_bench_v1.json; echo
date
