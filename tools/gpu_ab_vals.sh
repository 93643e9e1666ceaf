#!/bin/bash
# bench A/B over several values of one environment switch: tools/gpu_ab_vals.sh VAR v1 v2 ...
mkdir -p gpurun_out
VAR=$1; shift
for rep in 1 2; do for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 50 --warmup 10 --quick > gpurun_out/ab_${VAR}_$v.log 2> gpurun_out/ab_${VAR}_$v.err
  echo -n "$VAR=$v: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_${VAR}_$v.log | head -1; grep -h "Error\|error" gpurun_out/ab_${VAR}_$v.err | tail -3
done; done
exit 0
