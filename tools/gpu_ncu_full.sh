#!/bin/bash
mkdir -p gpurun_out
python tools/prof_kernels.py 3 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kd_token_loss|gemm_tc' -s 3 -c 6 -f -o gpurun_out/prof_r1b python tools/prof_kernels.py 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
