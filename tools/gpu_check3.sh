#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep -v "Warning\|run_backward\|^$" gpurun_out/$name.log | tail -n 2 | cut -c1-250; }
run t_gpu 1200 python -m pytest tests -m gpu -q
run bench 400 python bench.py --steps 100 --warmup 20 --no-cpu-baseline
echo -n "bench: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench.log | head -1
python tools/timeline.py 2>&1 | grep -v Warning | tail -5
exit 0
