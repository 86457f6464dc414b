#!/bin/bash
# 8-GPU session v3: copy-engine push (CHEQ_DIST_P2P=3): parity, S20k bench (tpp 4 / 2), S100k.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export CHEQ_DIST_P2P=3
CHEQ_DIST_TEST_BIG=1000 timeout 300 $TR --master-port 29511 tests/dist_worker.py > gpurun_out/dist${N}_worker_dma.log 2>&1
echo "worker(dma) rc=$?"; tail -3 gpurun_out/dist${N}_worker_dma.log
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_dma.json 2> gpurun_out/bench_dist${N}_dma.err
echo "bench(dma) rc=$?"
CHEQ_TILES_PER_PANEL=2 timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_dma_tpp2.json 2> gpurun_out/bench_dist${N}_dma_tpp2.err
echo "bench(dma,tpp2) rc=$?"
timeout 300 $TR --master-port 29515 tools/dist_s100k.py --rows 4 > gpurun_out/s100k_dist${N}_dma.json 2> gpurun_out/s100k_dist${N}_dma.err
echo "s100k rc=$?"; grep "^{" gpurun_out/s100k_dist${N}_dma.json | cut -c1-1500
python - <<PY
import json
for tag in ("dma", "dma_tpp2"):
    f = "gpurun_out/bench_dist${N}_%s.json" % tag
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(tag, "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "scf_ms", round(d["config"]["phases_ms"]["ms_scf"], 1),
              "once", round(d["config"]["phases_ms"]["ms_factor_once"], 1), "md", round(d["secondary"]["md"]["frames_per_s"]))
    except Exception as e:
        print(tag, "FAILED", e)
PY
