#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo -n "exit $? $name: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$name.log | head -1; }
B="python bench.py --steps 100 --warmup 20 --no-cpu-baseline"
run ab_p1_wg2 300 $B
run ab_p1_wg1 300 env B2C_WG=1 $B
run ab_p0_wg2 300 env B2C_CHAIN_PRIORITY=0 $B
run ab_p0_wg1 300 env B2C_CHAIN_PRIORITY=0 B2C_WG=1 $B
run ab_p1_wg2b 300 $B
python tools/timeline.py 2>&1 | grep -v Warning | tail -6
run t_parity 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x
tail -2 gpurun_out/t_parity.log
