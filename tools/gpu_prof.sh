#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 8 gpurun_out/$name.log; }
run t_kernels 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q
run t_parity 900 python -m pytest tests/test_gpu_parity.py -m gpu -q
run bench_eager 300 python bench.py --no-graph --steps 10 --warmup 3 --no-cpu-baseline
run bench 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run bench_plain 300 python bench.py --profile --no-graph --steps 2 --warmup 2
if grep -q profile_run gpurun_out/bench_plain.log; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_warm.csv python bench.py --profile --no-graph --steps 2 --warmup 2 > gpurun_out/ncu_launch.log 2>&1
  echo "ncu exit $?"; wc -l gpurun_out/launches_warm.csv
fi
