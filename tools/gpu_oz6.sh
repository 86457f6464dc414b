#!/bin/bash
# column-list mode of the tcgen05 path: probe, panel schedule on one GPU (the multi-GPU code path), parity tests
set -u
TAG=${1:-oz6}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
echo "== probe quick"; date
timeout 120 tools/ozaki_probe quick > $OUT/${TAG}_probe_quick.json 2>&1; rc=$?; echo "rc=$rc"; grep -E "column list|failed" $OUT/${TAG}_probe_quick.json
[ $rc -eq 0 ] || exit 1
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err; echo "rc=$? $name"
  python - <<PY
import json
for l in open("$OUT/${TAG}_bench_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d.get("parity") or {}; c = d["config"]
        print("  value %.4f  dq %.2e  it %s/%s  factor_once %.1f  scf %.1f  krylov_ms %.1f" % (
            d["value"], p.get("dq", -1), p.get("iterations"), p.get("iterations_ref"), c["phases_ms"]["ms_factor_once"], c["phases_ms"]["ms_scf"], c["krylov_ms"]))
PY
}
run panels4_dmma CHEQ_FACTOR_MODE=1 CHEQ_OZAKI=0
run panels4 CHEQ_FACTOR_MODE=1
run panels8 CHEQ_FACTOR_MODE=1 CHEQ_TILES_PER_PANEL=8
run panels16 CHEQ_FACTOR_MODE=1 CHEQ_TILES_PER_PANEL=16
echo "== pytest dist world=1 + panels parity"; date
timeout 900 python -m pytest tests/test_dist_gpu.py tests/test_gpu_headline.py -m gpu -q -s --timeout 600 -k "distributed or w8k" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log; grep -E "FAILED|Error|PASS" $OUT/${TAG}_pytest.log | head
date
