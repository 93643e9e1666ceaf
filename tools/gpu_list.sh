#!/bin/bash
# per-launch device times of two eagerly issued KD steps (ncu launch list; the same command runs plainly first)
mkdir -p gpurun_out
python bench.py --profile --no-graph --steps 2 --warmup 2 > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv python bench.py --profile --no-graph --steps 2 --warmup 2 > gpurun_out/ncu_launch.log 2>&1
echo "ncu list exit $?"; tail -1 gpurun_out/bench_plain.log | cut -c1-200
