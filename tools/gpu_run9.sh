#!/bin/bash
# GPU session 9 (8 GPUs): S100k with tile-major Krylov operands + per-step GMRES timestamps.
set -u
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=${GPUS:-8}
echo "== bench s100k on $N GPUs (debug trace on)"; date
CHEQ_KRYLOV_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --workload s100k --steps 2 --warmup 1 > $OUT/${TAG}_bench_s100k_${N}gpu.json 2> $OUT/${TAG}_bench_s100k_${N}gpu.err
echo "rc=$?"; head -c 200 $OUT/${TAG}_bench_s100k_${N}gpu.json; echo
date
