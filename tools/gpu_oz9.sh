#!/bin/bash
# 8 GPUs: S20k with the tcgen05 path in the trailing updates (NCCL data plane beyond 4 GPUs)
set -u
TAG=${1:-oz9}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
N=${GPUS:-8}
echo "== bench s20k on $N GPUs"; date
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 4 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench_dist${N}.json 2> $OUT/${TAG}_bench_dist${N}.err
echo "rc=$?"; python - <<PY
import json
for l in open("$OUT/${TAG}_bench_dist${N}.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d.get("parity") or {}; c = d["config"]
        print("  value %.4f  dq %.2e  it %s/%s  factor_once %.1f  scf %.1f  krylov_ms %.1f dist_check %s" % (
            d["value"], p.get("dq", -1), p.get("iterations"), p.get("iterations_ref"), c["phases_ms"]["ms_factor_once"], c["phases_ms"]["ms_scf"], c["krylov_ms"], d.get("dist_check")))
PY
tail -2 $OUT/${TAG}_bench_dist${N}.err
date
