#!/bin/bash
# tests + bench on one GPU; every group under its own timeout
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep -v "Warning\|run_backward\|^$" gpurun_out/$name.log | tail -n 4 | cut -c1-400; }
run t_kernels 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x
run t_parity 900 python -m pytest tests/test_gpu_parity.py -m gpu -q
run bench 400 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS:---no-cpu-baseline}
if [ -n "$BENCH_AB" ]; then B2C_PDL=0 run bench_nopdl 400 env B2C_PDL=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline; fi
