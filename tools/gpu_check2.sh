#!/bin/bash
# all GPU tests + one bench run (value + e2e), each under its own timeout
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep -v "Warning\|run_backward\|^$" gpurun_out/$name.log | tail -n 3 | cut -c1-330; }
run t_gpu 1200 python -m pytest tests -m gpu -q
run bench 400 python bench.py --steps 100 --warmup 20 ${BENCH_ARGS:---no-cpu-baseline}
if [ -n "$BENCH_AB" ]; then run bench_b 400 env $BENCH_AB python bench.py --steps 100 --warmup 20 --no-cpu-baseline; fi
for f in bench bench_b; do [ -f gpurun_out/$f.log ] && { echo -n "$f: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1; }; done
