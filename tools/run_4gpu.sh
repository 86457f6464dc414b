#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_nccl_flat.json 2> gpurun_out/bench_dist${N}_nccl_flat.err
echo "rc=$?"
CHEQ_PANEL_FLAT=0 timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_nccl_rec.json 2> gpurun_out/bench_dist${N}_nccl_rec.err
echo "rc=$?"
CHEQ_DIST_P2P=3 timeout 200 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}_dma_flat.json 2> gpurun_out/bench_dist${N}_dma_flat.err
echo "rc=$?"
python - <<PY
import json
for tag in ("nccl_flat", "nccl_rec", "dma_flat"):
    f = "gpurun_out/bench_dist${N}_%s.json" % tag
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(tag, "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "scf_ms", round(d["config"]["phases_ms"]["ms_scf"], 1),
              "once", round(d["config"]["phases_ms"]["ms_factor_once"], 1), "md", round(d["secondary"]["md"]["frames_per_s"]))
    except Exception as e:
        print(tag, "FAILED", e)
PY
