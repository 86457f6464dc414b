#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_final_ref.json 2> gpurun_out/bench_r1_final_ref.err; echo "ref rc=$?"
wc -l gpurun_out/bench_r1_final.json gpurun_out/bench_r1_final_ref.json
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r1_final.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline frac", r["frac"], "achieved", r["achieved"], "peak", r["peak"], "share", r["share_of_step"],
      "cpu", d.get("cpu_baseline", {}).get("value"), "launches", d["gpu_launches"], "clocks", d["clocks"])
print("md", d["secondary"]["md"]["frames_per_s"], "b3k", d["secondary"]["b3k"]["ms_per_solve"])
d = json.loads(open("gpurun_out/bench_r1_final_ref.json").read().strip().splitlines()[-1])
print("ref", d["value"], d["cpu_baseline"]["cores"])
PY
