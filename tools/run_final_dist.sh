#!/bin/bash
# final defaults: parity worker + bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CHEQ_DIST_TEST_BIG=1000 timeout 300 $TR --master-port 29511 tests/dist_worker.py > gpurun_out/dist${N}_worker_final.log 2>&1
echo "worker rc=$?"; tail -2 gpurun_out/dist${N}_worker_final.log
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_dist${N}_final.json 2> gpurun_out/bench_dist${N}_final.err
echo "bench rc=$?"
python - <<PY
import json
f = "gpurun_out/bench_dist${N}_final.json"
txt = open(f).read().strip().splitlines()
print("stdout lines:", len(txt))
d = json.loads(txt[-1])
print("value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "scf_ms", round(d["config"]["phases_ms"]["ms_scf"], 1), "md", round(d["secondary"]["md"]["frames_per_s"]))
PY
