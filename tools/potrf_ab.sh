#!/bin/bash
# quick A/B of the diagonal-tile kernels: correctness (Cholesky solve residual, a few parity tests) + time per launch
python tools/potrf_probe.py 2>&1 | grep -E "residual"
CHEQ_POTRF_CW=0 python tools/gpu_probe.py 1000 2>&1 | tail -1 | sed 's/.*| gemm/v1: gemm/'
python tools/gpu_probe.py 1000 2>&1 | tail -1 | sed 's/.*| gemm/cw: gemm/'
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cholesky or rappe or indefinite or pivoted or batch or clusters" 2>&1 | tail -2
