#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo -n "exit $? $name: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$name.log | head -1; }
B="python bench.py --steps 100 --warmup 20 --no-cpu-baseline"
run ab_all 300 $B
run ab_nobearly 300 env B2C_GEMM_BEARLY=0 $B
run ab_nocellpre 300 env B2C_CELL_PRE=0 $B
run ab_neither 300 env B2C_GEMM_BEARLY=0 B2C_CELL_PRE=0 $B
run ab_all2 300 $B
run t_optim 300 python -m pytest tests/test_gpu_optim.py -m gpu -q
tail -2 gpurun_out/t_optim.log
