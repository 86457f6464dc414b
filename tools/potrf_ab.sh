#!/bin/bash
# quick A/B of the diagonal-tile kernels: time per solve phase + parity tests under each variant
for m in 0 1 2; do
  CHEQ_POTRF_CW=$m python tools/gpu_probe.py 1000 2>&1 | tail -1 | sed "s/.*| gemm/cw=$m: gemm/"
done
for m in 1 2; do
  CHEQ_POTRF_CW=$m timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cholesky or rappe or indefinite or pivoted or batch or clusters" 2>&1 | tail -1
done
