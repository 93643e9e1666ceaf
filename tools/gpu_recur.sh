#!/bin/bash
# persistent-recurrence development loop: the parity tests that exercise it, the phase trace, and an A/B bench against the per-step kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "config1 or config2 or large_variant or graphed or dropout" > gpurun_out/t_recur.log 2>&1; echo "tests exit $?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/t_recur.log | cut -c1-300
timeout 300 python tools/recur_trace.py 2>&1 | grep -v Warning | tail -12
for v in 1 0; do
  env B2C_PERSISTENT=$v timeout 300 python bench.py --steps 50 --warmup 10 --quick > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err
  echo -n "B2C_PERSISTENT=$v: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_$v.log | head -1
done
env B2C_PERSISTENT=1 timeout 300 python tools/timeline.py 2>&1 | grep -v Warning | grep -E "span|recur" 
exit 0
