#!/bin/bash
mkdir -p gpurun_out
python tools/prof_kernels.py 3 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kd_token_loss' -s 1 -c 2 -f -o gpurun_out/prof_kd python tools/prof_kernels.py 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu kd exit $?"
python bench.py --profile --no-graph --steps 2 --warmup 2 > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_warm2.csv python bench.py --profile --no-graph --steps 2 --warmup 2 > gpurun_out/ncu_launch.log 2>&1
echo "ncu list exit $?"
