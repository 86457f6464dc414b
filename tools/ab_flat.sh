#!/bin/bash
# one GPU: tests, then the blocked schedule with and without tile-level look-ahead inside the panel
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CHEQ_FACTOR_MODE=1 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/flat_mode1.json 2> gpurun_out/flat_mode1.err
CHEQ_PANEL_FLAT=0 CHEQ_FACTOR_MODE=1 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/rec_mode1.json 2> gpurun_out/rec_mode1.err
python - <<PY
import json
for f in ("gpurun_out/flat_mode1.json", "gpurun_out/rec_mode1.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, round(d["value"], 4), "scf", round(d["config"]["phases_ms"]["ms_scf"], 1), "once", round(d["config"]["phases_ms"]["ms_factor_once"], 1),
              "potrf", round(r["potrf_ms_per_step"], 1), "md", round(d["secondary"]["md"]["frames_per_s"]), "b3k", round(d["secondary"]["b3k"]["ms_per_solve"], 1))
    except Exception as e:
        print(f, "FAIL", e)
PY
