#!/bin/bash
# A/B of one environment switch: quick bench (eval-mode headline) + timeline for both settings
mkdir -p gpurun_out
for v in 1 0; do
  env ${AB_VAR:-B2C_PERSISTENT}=$v timeout 300 python bench.py --steps 50 --warmup 10 --quick > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err
  echo -n "${AB_VAR:-B2C_PERSISTENT}=$v: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_$v.log | head -1
  env ${AB_VAR:-B2C_PERSISTENT}=$v timeout 300 python tools/timeline.py 2>&1 | grep -v Warning | tail -32 > gpurun_out/tl_$v.txt
  cp gpurun_out/timeline.csv gpurun_out/timeline_$v.csv
done
exit 0
