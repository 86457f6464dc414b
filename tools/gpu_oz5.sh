#!/bin/bash
set -u
TAG=${1:-oz5}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
echo "== pytest (ozaki gemm + headline S20k)"; date
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -q -s --timeout 600 -k "ozaki or s20k_full or gemm_nt" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log; grep -E "^\[S20k|FAILED|Error" $OUT/${TAG}_pytest.log | head -20
for oz in 0 7; do
  echo "== b3k CHEQ_OZAKI=$oz"; CHEQ_OZAKI=$oz timeout 300 python bench.py --workload b3k --steps 10 --warmup 3 > $OUT/${TAG}_b3k_oz$oz.json 2> $OUT/${TAG}_b3k_oz$oz.err; echo "rc=$?"; head -c 160 $OUT/${TAG}_b3k_oz$oz.json; echo
done
echo "== md 2048 frames"; timeout 300 python bench.py --workload md --frames 2048 --steps 2 --warmup 2 > $OUT/${TAG}_md2048.json 2> $OUT/${TAG}_md2048.err; echo "rc=$?"; head -c 160 $OUT/${TAG}_md2048.json; echo
echo "== s20k"; timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"
python - <<PY
import json
for l in open("$OUT/${TAG}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "e2e", d["e2e"], "parity", d["parity"]["dq"])
PY
date
