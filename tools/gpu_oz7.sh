#!/bin/bash
# 2 GPUs: dist parity with the tcgen05 path in the trailing updates, S20k with 4- and 8-tile panels
set -u
TAG=${1:-oz7}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1 CHEQ_ORACLE_CACHE=tools/dev_cache
N=${GPUS:-2}
echo "== dist worker on $N ranks"; date
CHEQ_DIST_TEST_BIG=1500 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29540 tests/dist_worker.py > $OUT/${TAG}_dist${N}_parity.txt 2>&1
echo "rc=$?"; tail -4 $OUT/${TAG}_dist${N}_parity.txt | cut -c1-200
i=0
for TPP in 4 8; do
  i=$((i+1))
  echo "== bench s20k on $N GPUs, tiles per panel $TPP"; date
  CHEQ_TILES_PER_PANEL=$TPP timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2955$i bench.py --gpus $N --steps 4 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench_dist${N}_tpp$TPP.json 2> $OUT/${TAG}_bench_dist${N}_tpp$TPP.err
  echo "rc=$?"; python - <<PY
import json
for l in open("$OUT/${TAG}_bench_dist${N}_tpp$TPP.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d.get("parity") or {}; c = d["config"]
        print("  value %.4f  dq %.2e  it %s/%s  factor_once %.1f  scf %.1f  krylov_ms %.1f dist_check %s" % (
            d["value"], p.get("dq", -1), p.get("iterations"), p.get("iterations_ref"), c["phases_ms"]["ms_factor_once"], c["phases_ms"]["ms_scf"], c["krylov_ms"], d.get("dist_check")))
PY
done
date
