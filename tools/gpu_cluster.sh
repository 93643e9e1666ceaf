#!/bin/bash
# cluster-recurrence development loop: A/B against the per-step kernels, then the graph-replay bench both ways
mkdir -p gpurun_out
timeout 900 python tools/cluster_ab.py ${AB_SIZES:-16 35 512} 2>&1 | grep -v Warning | tail -40
if [ -n "$BENCH" ]; then
for v in 1 0; do
  env B2C_CLUSTER=$v timeout 300 python bench.py --steps 50 --warmup 10 --quick > gpurun_out/abc_$v.log 2> gpurun_out/abc_$v.err
  echo -n "B2C_CLUSTER=$v: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/abc_$v.log | head -1; tail -c 600 gpurun_out/abc_$v.err
done
fi
exit 0
