#!/bin/bash
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_dist${N}_final.json 2> gpurun_out/bench_dist${N}_final.err
echo "bench rc=$?"
python - <<PY
import json
txt = open("gpurun_out/bench_dist8_final.json").read().strip().splitlines()
d = json.loads(txt[-1])
print("lines", len(txt), "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "scf_ms", round(d["config"]["phases_ms"]["ms_scf"], 1), "md", round(d["secondary"]["md"]["frames_per_s"]))
PY
