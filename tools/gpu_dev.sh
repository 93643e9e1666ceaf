#!/bin/bash
# development loop: every GPU test + one full bench run (no ncu), optional A/B env in $BENCH_AB
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo -n "exit $? $name: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$name.log | head -1; echo; }
timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_ARGS} > gpurun_out/t_gpu.log 2>&1; echo "tests exit $?"; grep -v "Warning\|run_backward\|^$" gpurun_out/t_gpu.log | tail -n 6 | cut -c1-400
grep -E "^(FAILED|ERROR)" gpurun_out/t_gpu.log | cut -c1-300
B="python bench.py --steps 20 --warmup 5"
run bench 600 $B
tail -c 1500 gpurun_out/bench.err
if [ -n "$BENCH_AB" ]; then run bench_ab 300 env $BENCH_AB $B --quick; run bench2 300 $B --quick; fi
exit 0
