#!/bin/bash
# Ozaki GEMM tuning: probe timings, then S20k with several thresholds / slice counts for R
set -u
TAG=${1:-oz2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
echo "== probe full"; date
timeout 300 tools/ozaki_probe > $OUT/${TAG}_probe.json 2>&1; echo "rc=$?"; grep -E "ms_total|failed" $OUT/${TAG}_probe.json | cut -c1-60,230-
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err; echo "rc=$? $name"
  python - <<PY
import json
for l in open("$OUT/${TAG}_bench_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d.get("parity") or {}; r = d["roofline"]; c = d["config"]
        print("  value %.4f  dq %.2e  it %s/%s  gemm_ms %.1f  potrf %.1f  backsolve %.1f  factor_once %.1f  scf %.1f  krylov_ms %.1f  steps %s" % (
            d["value"], p.get("dq", -1), p.get("iterations"), p.get("iterations_ref"), r["kernel_ms_per_step"], r["potrf_ms_per_step"],
            r["backsolve_ms_per_step"], c["phases_ms"]["ms_factor_once"], c["phases_ms"]["ms_scf"], c["krylov_ms"], sum(c["gmres_steps"])))
PY
}
run k1024 CHEQ_OZAKI=7
run k512 CHEQ_OZAKI=7 CHEQ_OZAKI_MIN_K=512
run k256 CHEQ_OZAKI=7 CHEQ_OZAKI_MIN_K=256
run k512_r5 CHEQ_OZAKI=7 CHEQ_OZAKI_MIN_K=512 CHEQ_OZAKI_R=5
run k512_r4 CHEQ_OZAKI=7 CHEQ_OZAKI_MIN_K=512 CHEQ_OZAKI_R=4
run s8_k512 CHEQ_OZAKI=8 CHEQ_OZAKI_MIN_K=512
date
