#!/bin/bash
# GPU session 6 (8 GPUs): S20k strong-scaling line at 8 ranks + per-step GMRES timing on the group (debug trace).
set -u
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=${GPUS:-8}
echo "== bench s20k on $N GPUs"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 --no-secondary > $OUT/${TAG}_bench_dist${N}.json 2> $OUT/${TAG}_bench_dist${N}.err
echo "rc=$?"; head -c 300 $OUT/${TAG}_bench_dist${N}.json; echo
echo "== same with the GMRES debug trace (2 steps)"; date
CHEQ_KRYLOV_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 1 --warmup 1 --no-secondary > $OUT/${TAG}_bench_dist${N}_debug.json 2> $OUT/${TAG}_bench_dist${N}_debug.err
echo "rc=$?"
date
