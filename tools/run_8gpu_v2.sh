#!/bin/bash
# 8-GPU session v2: parity with the peer-push data plane, S20k bench (p2p tpp 4 / 2, NCCL), S100k with p2p.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CHEQ_DIST_TEST_BIG=1000 timeout 300 $TR --master-port 29511 tests/dist_worker.py > gpurun_out/dist${N}_worker_p2p.log 2>&1
echo "worker(p2p) rc=$?"; tail -3 gpurun_out/dist${N}_worker_p2p.log
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_p2p.json 2> gpurun_out/bench_dist${N}_p2p.err
echo "bench(p2p) rc=$?"
CHEQ_TILES_PER_PANEL=2 timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_p2p_tpp2.json 2> gpurun_out/bench_dist${N}_p2p_tpp2.err
echo "bench(p2p,tpp2) rc=$?"
CHEQ_DIST_P2P=0 timeout 200 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_nccl.json 2> gpurun_out/bench_dist${N}_nccl.err
echo "bench(nccl) rc=$?"
timeout 300 $TR --master-port 29515 tools/dist_s100k.py --rows 4 > gpurun_out/s100k_dist${N}_p2p.json 2> gpurun_out/s100k_dist${N}_p2p.err
echo "s100k rc=$?"; grep "^{" gpurun_out/s100k_dist${N}_p2p.json | cut -c1-1500
python - <<PY
import json
for tag in ("p2p", "p2p_tpp2", "nccl"):
    f = "gpurun_out/bench_dist${N}_%s.json" % tag
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(tag, "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "scf_ms", round(d["config"]["phases_ms"]["ms_scf"], 1),
              "once", round(d["config"]["phases_ms"]["ms_factor_once"], 1), "md", round(d["secondary"]["md"]["frames_per_s"]))
    except Exception as e:
        print(tag, "FAILED", e)
PY
