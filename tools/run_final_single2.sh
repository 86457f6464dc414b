#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_auto.json 2> gpurun_out/bench_r1_auto.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r1_auto.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", r["frac"], "achieved", r["achieved"], "share", r["share_of_step"], "launches", d["gpu_launches"])
print("md", d["secondary"]["md"]["frames_per_s"], "b3k", d["secondary"]["b3k"]["ms_per_solve"])
PY
