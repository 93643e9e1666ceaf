#!/bin/bash
# tests, then A/B of the attention-step variants and of the optimizer (each a plain bench run)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep -v "Warning\|run_backward\|^$" gpurun_out/$name.log | tail -n 3 | cut -c1-330; }
run t_optim 300 python -m pytest tests/test_gpu_optim.py -m gpu -q -x
run t_kernels 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x
run t_parity 900 python -m pytest tests/test_gpu_parity.py -m gpu -q
B="python bench.py --steps 100 --warmup 20 --no-cpu-baseline"
run b_ka4_pb13 300 env B2C_ATT_FWD_KA=4 B2C_ATT_BWD_PB=13 $B
run b_ka7_pb13 300 env B2C_ATT_FWD_KA=7 B2C_ATT_BWD_PB=13 $B
run b_ka4_pb8 300 env B2C_ATT_FWD_KA=4 B2C_ATT_BWD_PB=8 $B
run b_ka4_pb25 300 env B2C_ATT_FWD_KA=4 B2C_ATT_BWD_PB=25 $B
run b_torchopt 300 env B2C_ATT_FWD_KA=4 B2C_ATT_BWD_PB=13 $B --torch-optimizer
for f in b_ka4_pb13 b_ka7_pb13 b_ka4_pb8 b_ka4_pb25 b_torchopt; do echo -n "$f: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1; done
