#!/bin/bash
# one `ncu --set full` capture of the headline kernels run stand-alone at config-2 shapes (plain run first)
mkdir -p gpurun_out
python tools/prof_kernels.py 3 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kd_token_loss|gemm_tc|mha_.*_mma|clip_adamw|grad_sqnorm|attn_post_reg' -s 21 -c 21 -f -o gpurun_out/prof_r1c python tools/prof_kernels.py 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/prof_plain.log
