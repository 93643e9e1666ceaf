#!/bin/bash
# ncu --set full of the persistent forward-recurrence kernel (one launch of the stand-alone decoder forward); plain run first
mkdir -p gpurun_out
python tools/recur_trace.py > gpurun_out/recur_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:recur_fwd -s 2 -c 1 -f -o gpurun_out/prof_recur python tools/recur_trace.py > gpurun_out/ncu_recur.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_recur.log; grep -E "grid|epi|mma|producer" gpurun_out/recur_plain.log
