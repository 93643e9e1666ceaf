#!/bin/bash
# development loop: every GPU test + one bench run (no ncu), optional A/B env in $BENCH_AB
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo -n "exit $? $name: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$name.log | head -1; echo; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "tests exit $?"; grep -v "Warning\|run_backward\|^$" gpurun_out/t_gpu.log | tail -n 4 | cut -c1-300
B="python bench.py --steps 100 --warmup 20 --no-cpu-baseline"
run bench 300 $B
if [ -n "$BENCH_AB" ]; then run bench_ab 300 env $BENCH_AB $B; run bench2 300 $B; fi
python tools/timeline.py 2>&1 | grep -v Warning | tail -3
exit 0
