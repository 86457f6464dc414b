#!/bin/bash
# GPU session 3 (2 GPUs): distributed parity worker (stale-factor + matrix-free on the group) and the 2-GPU bench line.
set -u
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== dist worker, 2 ranks"; date
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29517 tests/dist_worker.py > $OUT/${TAG}_dist2_parity.txt 2>&1
echo "rc=$?"; tail -12 $OUT/${TAG}_dist2_parity.txt
echo "== bench --gpus 2"; date
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --no-secondary > $OUT/${TAG}_bench_dist2.json 2> $OUT/${TAG}_bench_dist2.err
echo "rc=$?"; tail -c 600 $OUT/${TAG}_bench_dist2.err
echo "== matrix-free S20k on 2 ranks"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29519 tools/matfree_run.py 6667 --no-dense > $OUT/${TAG}_matfree_s20k_2gpu.txt 2>&1
echo "rc=$?"; tail -3 $OUT/${TAG}_matfree_s20k_2gpu.txt
date
