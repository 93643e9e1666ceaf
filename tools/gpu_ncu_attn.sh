#!/bin/bash
# ncu --set full of the attention / MHA kernels inside a real (eager) KD step; plain run first.
mkdir -p gpurun_out
CMD="python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'attn_step|attn_post|mha_' -s 19 -c 24 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"
tail -2 gpurun_out/attn_plain.log
