#!/bin/bash
# GPU session 8 (8 GPUs): S20k at 8 ranks with the peer-memory all-reduce vs ncclAllReduce, then S100k again.
set -u
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=${GPUS:-8}
echo "== dist worker on $N ranks (peer-memory all-reduce forced on)"; date
CHEQ_KRYLOV_P2P=1 CHEQ_DIST_TEST_BIG=1500 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29540 tests/dist_worker.py > $OUT/${TAG}_dist${N}_parity.txt 2>&1
echo "rc=$?"; tail -4 $OUT/${TAG}_dist${N}_parity.txt | cut -c1-200
for P2P in 1 0; do
  echo "== bench s20k on $N GPUs, CHEQ_KRYLOV_P2P=$P2P"; date
  CHEQ_KRYLOV_P2P=$P2P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2954$P2P bench.py --gpus $N --steps 5 --warmup 3 --no-secondary > $OUT/${TAG}_bench_dist${N}_p2p$P2P.json 2> $OUT/${TAG}_bench_dist${N}_p2p$P2P.err
  echo "rc=$?"; head -c 160 $OUT/${TAG}_bench_dist${N}_p2p$P2P.json; echo
done
echo "== bench s100k on $N GPUs"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --workload s100k --steps 2 --warmup 1 > $OUT/${TAG}_bench_s100k_${N}gpu.json 2> $OUT/${TAG}_bench_s100k_${N}gpu.err
echo "rc=$?"; head -c 200 $OUT/${TAG}_bench_s100k_${N}gpu.json; echo
echo "== tcgen05 probe (GPU 0)"; date
timeout 60 tools/tc05_probe > $OUT/${TAG}_tc05_probe.json 2>&1; echo "rc=$?"; cat $OUT/${TAG}_tc05_probe.json
date
