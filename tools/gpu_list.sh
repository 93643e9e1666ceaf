#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "greedy" tests/test_gpu_parity.py > gpurun_out/t_new.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/t_new.log
python bench.py --profile --no-graph --steps 2 --warmup 2 > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv python bench.py --profile --no-graph --steps 2 --warmup 2 > gpurun_out/ncu_launch.log 2>&1
echo "ncu list exit $?"
