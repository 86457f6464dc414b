#!/bin/bash
# One 8-GPU session: parity worker, S20k strong-scaling bench, S100k distributed solve.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CHEQ_DIST_TEST_BIG=1500 timeout 300 $TR --master-port 29511 tests/dist_worker.py > gpurun_out/dist${N}_worker.log 2>&1
echo "worker rc=$?"; tail -12 gpurun_out/dist${N}_worker.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_dist${N}.json 2> gpurun_out/bench_dist${N}.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_dist${N}.err; head -c 1500 gpurun_out/bench_dist${N}.json; echo
timeout 420 $TR --master-port 29513 tools/dist_s100k.py > gpurun_out/s100k_dist${N}.json 2> gpurun_out/s100k_dist${N}.err
echo "s100k rc=$?"; tail -c 600 gpurun_out/s100k_dist${N}.err; cat gpurun_out/s100k_dist${N}.json
