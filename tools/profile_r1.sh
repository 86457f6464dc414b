#!/bin/bash
# ncu evidence for the round: launch list of one S20k solve (second repetition) + full captures of the chained
# back substitution and of the list-mode GEMM of the blocked schedule.  Each ncu pass follows a plain run.
mkdir -p gpurun_out
python tools/gpu_probe.py 6667 > gpurun_out/plain_s20k.log 2>&1 || { tail -5 gpurun_out/plain_s20k.log; exit 1; }
tail -2 gpurun_out/plain_s20k.log
L=$(grep -o "launches [0-9]*" gpurun_out/plain_s20k.log | tail -1 | awk '{print $2}')
echo "launches per solve: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $L -c $L --csv \
    --log-file gpurun_out/launches_s20k_v3.csv python tools/gpu_probe.py 6667 > gpurun_out/ncu_list_v3.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_s20k_v3.csv > gpurun_out/launches_s20k_v3.txt; cat gpurun_out/launches_s20k_v3.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:backsolve_chain -s 20 -c 2 \
    -o gpurun_out/prof_chain python tools/gpu_probe.py 6667 > gpurun_out/ncu_chain.log 2>&1; tail -2 gpurun_out/ncu_chain.log
CHEQ_FACTOR_MODE=1 python tools/gpu_probe.py 6667 > gpurun_out/plain_s20k_blocked.log 2>&1 && tail -1 gpurun_out/plain_s20k_blocked.log
CHEQ_FACTOR_MODE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 2500 -c 3 \
    -o gpurun_out/prof_gemm_list python tools/gpu_probe.py 6667 > gpurun_out/ncu_gemm_list.log 2>&1; tail -2 gpurun_out/ncu_gemm_list.log
