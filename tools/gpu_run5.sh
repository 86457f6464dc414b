#!/bin/bash
# GPU session 5 (8 GPUs): S100k through bench.py (dense direct path + stale-factor iterations on the group) and the
# matrix-free path on a 200,001-atom cluster that no dense path can hold.
set -u
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=${GPUS:-8}
echo "== bench s100k on $N GPUs"; date
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload s100k --steps 2 --warmup 1 > $OUT/${TAG}_bench_s100k_${N}gpu.json 2> $OUT/${TAG}_bench_s100k_${N}gpu.err
echo "rc=$?"; tail -c 400 $OUT/${TAG}_bench_s100k_${N}gpu.err; head -c 300 $OUT/${TAG}_bench_s100k_${N}gpu.json; echo
echo "== matrix-free, 200,001 atoms on $N GPUs"; date
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29522 tools/matfree_run.py 66667 --no-dense > $OUT/${TAG}_matfree_200k_${N}gpu.txt 2>&1
echo "rc=$?"; tail -2 $OUT/${TAG}_matfree_200k_${N}gpu.txt | cut -c1-600
nvidia-smi --query-gpu=index,memory.used,memory.total --format=csv > $OUT/${TAG}_smi.txt 2>&1
date
