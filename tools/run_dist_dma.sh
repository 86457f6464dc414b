#!/bin/bash
# N-GPU: parity + S20k bench with the copy-engine push (CHEQ_DIST_P2P=3) and the automatic choice
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CHEQ_DIST_P2P=3 CHEQ_DIST_TEST_BIG=1000 timeout 300 $TR --master-port 29511 tests/dist_worker.py > gpurun_out/dist${N}_worker_dma.log 2>&1
echo "worker(dma) rc=$?"; tail -3 gpurun_out/dist${N}_worker_dma.log
CHEQ_DIST_P2P=3 timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_dma.json 2> gpurun_out/bench_dist${N}_dma.err
echo "bench(dma) rc=$?"
timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_auto.json 2> gpurun_out/bench_dist${N}_auto.err
echo "bench(auto) rc=$?"
python - <<PY
import json
for tag in ("dma", "auto"):
    f = "gpurun_out/bench_dist${N}_%s.json" % tag
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(tag, "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "scf_ms", round(d["config"]["phases_ms"]["ms_scf"], 1),
              "once", round(d["config"]["phases_ms"]["ms_factor_once"], 1), "md", round(d["secondary"]["md"]["frames_per_s"]))
    except Exception as e:
        print(tag, "FAILED", e)
PY
