#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo -n "exit $? $name: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$name.log | head -1; echo; }
B="python bench.py --steps 100 --warmup 20 --no-cpu-baseline"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "tests exit $?"; grep -v "Warning\|run_backward\|^$" gpurun_out/t_gpu.log | tail -n 4 | cut -c1-300
run ab_new 300 $B
grep -o '"kernels": {.*}}' gpurun_out/ab_new.log | cut -c1-600
python tools/timeline.py 2>&1 | grep -v Warning | tail -3
exit 0
