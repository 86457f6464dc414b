#!/bin/bash
# N-GPU parity + S20k bench with the peer-push data plane (default) and with NCCL broadcasts (CHEQ_DIST_P2P=0)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CHEQ_DIST_TEST_BIG=1500 timeout 300 $TR --master-port 29511 tests/dist_worker.py > gpurun_out/dist${N}_worker_p2p.log 2>&1
echo "worker(p2p) rc=$?"; tail -4 gpurun_out/dist${N}_worker_p2p.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_p2p.json 2> gpurun_out/bench_dist${N}_p2p.err
echo "bench(p2p) rc=$?"; tail -c 400 gpurun_out/bench_dist${N}_p2p.err
CHEQ_DIST_P2P=0 timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_nccl.json 2> gpurun_out/bench_dist${N}_nccl.err
echo "bench(nccl) rc=$?"
python - <<PY
import json
for f in ("gpurun_out/bench_dist${N}_p2p.json", "gpurun_out/bench_dist${N}_nccl.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "scf_ms", round(d["config"]["phases_ms"]["ms_scf"], 1),
              "once", round(d["config"]["phases_ms"]["ms_factor_once"], 1), "md", round(d["secondary"]["md"]["frames_per_s"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
