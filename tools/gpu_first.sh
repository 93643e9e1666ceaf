#!/bin/bash
# first GPU pass: every group in its own process under a timeout so one bad kernel cannot take the rest down
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 25 gpurun_out/$name.log; }
run t_gemm_ffma 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -rA -k "fp32_ffma"
run t_gemm_tc 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -rA -k "tcgen05"
run t_kernels 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -rA -k "not gemm"
run t_parity 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA
run smoke 300 python __graft_entry__.py --smoke
run bench 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
