#!/bin/bash
# GPU session 4 (1 GPU): validate fp32 preconditioner storage + trusted estimate, profile the matrix-free path.
set -u
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest subset"; date
timeout 1500 python -m pytest tests/test_gpu_headline.py -m gpu -q -s --timeout 1200 -k "s20k or w8k or policy or matrix_free" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
echo "== bench s20k"; date
timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo rc=$?
echo "== bench s20k fp64 R"; date
CHEQ_KRYLOV_F32=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-secondary > $OUT/${TAG}_bench_f64R.json 2> $OUT/${TAG}_bench_f64R.err; echo rc=$?
echo "== matrix-free S20k (timing)"; date
timeout 600 python tools/matfree_run.py 6667 --no-dense > $OUT/${TAG}_matfree_s20k.txt 2>&1; echo rc=$?; tail -2 $OUT/${TAG}_matfree_s20k.txt
echo "== ncu launch list: matrix-free W8k"; date
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 3000 --csv --log-file $OUT/${TAG}_launches_matfree.csv \
    python tools/matfree_run.py 2700 --no-dense > $OUT/${TAG}_ncu_mf.log 2>&1; echo rc=$?
echo "== ncu full: matfree_matvec_kernel"; date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:matfree_matvec_kernel -s 20 -c 2 \
    -o $OUT/${TAG}_ncu_matfree python tools/matfree_run.py 6667 --no-dense > $OUT/${TAG}_ncu_mf2.log 2>&1; echo rc=$?
date
