#!/bin/bash
# GPU session 7 (2 GPUs): peer-memory all-reduce + rectangular W pass: dist parity, 2-GPU bench, 1-GPU regression.
set -u
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== dist worker, 2 ranks"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29517 tests/dist_worker.py > $OUT/${TAG}_dist2_parity.txt 2>&1
echo "rc=$?"; tail -8 $OUT/${TAG}_dist2_parity.txt | cut -c1-200
echo "== bench --gpus 2"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --no-secondary > $OUT/${TAG}_bench_dist2.json 2> $OUT/${TAG}_bench_dist2.err
echo "rc=$?"; head -c 200 $OUT/${TAG}_bench_dist2.json; echo
timeout 900 python -m pytest tests/test_gpu_headline.py -m gpu -q -s --timeout 600 -k "w8k or policy" > $OUT/${TAG}_pytest.log 2>&1
echo "rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "rc=$?"; head -c 200 $OUT/${TAG}_bench.json; echo
date
