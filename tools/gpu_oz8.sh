#!/bin/bash
# S100k on N GPUs with the tcgen05 path in the trailing updates
set -u
TAG=${1:-oz8}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=${GPUS:-2}
TPP=${TPP:-8}
echo "== bench s100k on $N GPUs, tiles per panel $TPP"; date
CHEQ_TILES_PER_PANEL=$TPP timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --workload s100k --steps 1 --warmup 1 > $OUT/${TAG}_bench_s100k_${N}gpu_tpp$TPP.json 2> $OUT/${TAG}_bench_s100k_${N}gpu_tpp$TPP.err
echo "rc=$?"; head -c 1500 $OUT/${TAG}_bench_s100k_${N}gpu_tpp$TPP.json; echo; tail -3 $OUT/${TAG}_bench_s100k_${N}gpu_tpp$TPP.err
date
