#!/bin/bash
# A/B of two builds of the library (B2C_LIB): quick bench each, twice (box noise)
mkdir -p gpurun_out
for rep in 1 2; do for lib in libb2c.so libb2c_deep.so; do
  env B2C_LIB=$PWD/imagecaptioner_b200/lib/$lib timeout 300 python bench.py --steps 50 --warmup 10 --quick > gpurun_out/ab_$lib.$rep.log 2> gpurun_out/ab_$lib.$rep.err
  echo -n "$lib rep $rep: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_$lib.$rep.log | head -1
done; done
timeout 300 python tools/timeline.py 2>&1 | grep -v Warning | tail -32 > gpurun_out/tl_new.txt; cp gpurun_out/timeline.csv gpurun_out/timeline_new.csv
exit 0
