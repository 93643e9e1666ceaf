#!/bin/bash
# round-end measurement pass on ONE GPU (no ncu here): tests, smoke, both bench arms, decode bench, kernel timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "tests exit $?"; grep -v "Warning\|run_backward\|^$" gpurun_out/t_gpu.log | tail -n 3 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 300 python bench.py --workload decode > gpurun_out/bench_decode.json 2> gpurun_out/bench_decode.err; echo "decode exit $?"; cut -c1-300 gpurun_out/bench_decode.json
timeout 300 python bench.py --torch-optimizer --no-cpu-baseline > gpurun_out/bench_torchopt.json 2>/dev/null; echo -n "torch optimizer A/B: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_torchopt.json
python tools/timeline.py 2>&1 | grep -v Warning | tail -6
exit 0
