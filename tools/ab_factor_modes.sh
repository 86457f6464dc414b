#!/bin/bash
# A/B of the factorisation schedules on one GPU: recursive (default) vs right-looking panels with look-ahead.
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab_mode0.json 2> gpurun_out/ab_mode0.err
for tpp in 2 4 8; do
  CHEQ_FACTOR_MODE=1 CHEQ_TILES_PER_PANEL=$tpp python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab_mode1_tpp$tpp.json 2> gpurun_out/ab_mode1_tpp$tpp.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_mode*.json")):
    try:
        d = json.load(open(f))
        r = d["roofline"]
        print(f, "value", round(d["value"], 4), "scf_ms", round(d["config"]["phases_ms"]["ms_scf"], 1), "once_ms",
              round(d["config"]["phases_ms"]["ms_factor_once"], 1), "gemm_ms", round(r["kernel_ms_per_step"], 1),
              "TF", round(r["achieved"], 1), "potrf", round(r["potrf_ms_per_step"], 1), "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "FAILED", e)
PY
